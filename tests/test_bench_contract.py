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
