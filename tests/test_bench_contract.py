"""CPU: the reference arm of bench.py runs anywhere (it is the oracle port on host cores) and
prints the contract's JSON line; the GPU arm is exercised by the driver on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '-L', '18',
                          '--steps', '2', '--warmup', '1', '--cpu-seconds', '0.2'],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    # stdout carries exactly the result line: anything a library prints there goes to stderr
    assert len(res.stdout.splitlines()) == 1, res.stdout[:500]
    out = json.loads(res.stdout)
    assert out['impl'] == 'reference' and out['metric'] == 'shell_matmult_per_s'
    assert out['higher_is_better'] is True and out['scaling'] == 'weak' and out['vs_baseline'] is None
    assert out['value'] > 0 and out['steps'] == 2 and out['warmup'] == 1
    assert out['cpu_baseline']['kind'] == 'port' and out['cpu_baseline']['cores'] >= 1
    assert out['e2e'] == {'value': out['value'], 'unit': out['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert out['gpu_launches'] == 0 and 'workload' in out['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ''


def test_both_arms_describe_the_same_workload():
    """The driver compares the two arms' `config` blocks key by key: they come from one function."""
    sys.path.insert(0, ROOT)
    import bench
    from dynamite_b200.hamiltonians import build_hamiltonian

    class Args:
        H = 'MBL'
        no_precompute_diagonal = False
    cfg = bench.workload_config(Args, 31, 2, build_hamiltonian('MBL', 31))
    assert cfg['L'] == 31 and cfg['rows'] == 1 << 31 and cfg['rows_per_gpu'] == 1 << 30
    assert cfg['unique_masks'] == 31 and cfg['nterms'] == 121 and cfg['precompute_diagonal'] is True
    ref = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '-L', '18',
                          '--steps', '1', '--warmup', '1', '--cpu-seconds', '0.2'],
                         capture_output=True, text=True, timeout=300)
    assert sorted(json.loads(ref.stdout)['config']) == sorted(cfg)


def test_analytic_host_input_and_parity_check():
    """bench.py's GPU arm multiplies an analytic host vector and checks sampled row blocks of the
    result against the oracle; here the 'result' is the oracle's own full product."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    import oracle
    from dynamite_b200.hamiltonians import build_hamiltonian
    L = 21
    n = 1 << L
    x = np.empty(n, dtype=np.complex128)
    bench.fill_host_x(x, 0, n)
    assert np.array_equal(x[5:105], bench.host_x(5, 100, n))
    assert np.array_equal(x[bench.XP + 7:bench.XP + 9], bench.host_x(bench.XP + 7, 2, n))
    half = np.empty(n // 2, dtype=np.complex128)           # rank 1 of 2
    bench.fill_host_x(half, n // 2, n)
    assert np.array_equal(half, x[n // 2:])
    H = build_hamiltonian('MBL', L)
    omsc, osub = bench.oracle_problem(H, L)
    y, _ = oracle.matmult_fast(omsc, osub, x, nthreads=8)
    assert bench.parity_check(H, L, y[n // 2:], n // 2, n, rows=1 << 12) < 1e-14
    y[n // 2 + 100] += 1e-6
    assert bench.parity_check(H, L, y[n // 2:], n // 2, n, rows=1 << 12) > 1e-9


def test_pipelined_e2e_is_kept_only_when_it_passes_the_oracle_check():
    """bench.pipelined_e2e poisons the rows parity_check samples before the timed batch call, so a call that
    does not write the result cannot inherit the single call's; here the 'GPU' is the oracle."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    import oracle
    from dynamite_b200.hamiltonians import build_hamiltonian
    L = 18
    n = 1 << L
    H = build_hamiltonian('MBL', L)
    omsc, osub = bench.oracle_problem(H, L)
    x = np.empty(n, dtype=np.complex128)
    bench.fill_host_x(x, 0, n)
    y = np.zeros(n, dtype=np.complex128)

    class Good:
        calls = []

        def mult_host_batch(self, xs, ys):
            self.calls.append(len(xs))
            for xi, yi in zip(xs, ys):
                yi[:] = oracle.matmult_fast(omsc, osub, xi, nthreads=4)[0]

    class Lazy:
        def mult_host_batch(self, xs, ys):
            pass

    good = bench.pipelined_e2e(Good(), x, y, 4, 1.0, H, L, 0, n)
    assert Good.calls == [2, 4] and good['steps'] == 4 and good['value'] > 0
    assert good['parity_rel_err_vs_oracle'] < 1e-14
    lazy = bench.pipelined_e2e(Lazy(), x, y, 4, 1.0, H, L, 0, n)     # y still holds the right answer ... minus the poison
    assert lazy['parity_rel_err_vs_oracle'] == float('inf')
