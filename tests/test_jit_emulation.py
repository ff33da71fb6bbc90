"""CPU (no GPU): what the GENERATED pass kernels compute, checked against the oracle.

dnm_jit_set_host_emulation(1) makes dnm_jit_dryrun emit the same kernel bodies it hands to NVRTC, with
a prelude that maps the CUDA vocabulary to C++: every CUDA thread of a block is an OS thread,
__syncthreads is a pthread barrier, cp.async copies are memcpy, the TMA tile load / reduce-add of a
tensor box become the box's nested copy / add loops, an mbarrier is a phase counter.  The source is
compiled with g++ and the passes of one MatMult are run on random vectors; the result must agree
with oracle.matmult (the restatement of the reference's MatMult_CPU) to rounding.  The plans are small
(2^13..2^16 rows per rank) but go through the same planner and generator as the L = 30..33 ones:
windows, FAR masks, TMA boxes, lean coefficient forms, pipelined rings and folded remote masks (the
peer shards are ordinary arrays here).

The GPU execution of the very same source is covered by tests/test_gpu_matmult.py and
tests/multi_gpu_worker.py; this file is what guards the generator where no GPU is present."""
import ctypes as C
import re
import subprocess

import numpy as np
import pytest

import oracle
from dynamite_b200 import _capi, msc_tools
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.subspaces import Full, Parity

from test_jit_generator import dryrun


class NotGenerated(Exception):
    pass


class Emulated:
    """One rank's generated passes, compiled for the host."""
    built = 0

    def __init__(self, tmp_path, name, L, tag, **kw):
        Emulated.built += 1
        tag = f'{tag}_{Emulated.built}'                       # dlopen caches by path: one file per build
        lib = _capi.lib()
        assert lib.dnm_jit_set_host_emulation(1) == 0
        try:
            self.info = dryrun(name, L, **kw)
        finally:
            lib.dnm_jit_set_host_emulation(0)
        assert self.info['cubin'] == 0                       # NVRTC is not involved in this mode
        if not self.info['kernels'] == self.info['passes'] >= 1:
            # (a pass that keeps the table-driven kernel cannot be emulated: that kernel is not generated source)
            raise NotGenerated({k: v for k, v in self.info.items() if k != 'src'})
        src = tmp_path / f'emu_{tag}.cpp'
        so = tmp_path / f'emu_{tag}.so'
        src.write_text(self.info['src'])
        res = subprocess.run(['g++', '-O1', '-shared', '-fPIC', '-pthread', '-o', str(so), str(src)],
                             capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr[-3000:]
        self.lib = C.CDLL(str(so))
        self.lib.dnm_emu_mult.restype = C.c_int
        self.lib.dnm_emu_mult.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int]
        assert self.lib.dnm_emu_npasses() == self.info['passes']

    def mult(self, shards, diag, pipelined_grid=3):
        """shards[h] = the input shard of rank (this ^ h); returns this rank's rows of the product."""
        n = shards[0].size
        y = np.full(n, np.nan + 1j * np.nan, dtype=np.complex128)      # pass 0 must write every row
        ptrs = (C.c_void_p * len(shards))(*[s.ctypes.data for s in shards])
        rc = self.lib.dnm_emu_mult(ptrs, len(shards), y.ctypes.data, diag.ctypes.data if diag is not None else None,
                                   n, pipelined_grid)
        assert rc == 0
        return y


def oracle_subspace(sub, L):
    if isinstance(sub, Parity):
        return oracle.Subspace({'type': 'parity', 'L': L, 'space': sub.space})
    return oracle.Subspace({'type': 'full', 'L': L})


def reference_product(name, L, sub, x):
    """(y, diag): the oracle's product and the cached diagonal the planner assumes (mask-0 terms)."""
    H = build_hamiltonian(name, L) if isinstance(name, str) else name
    H.reduce_msc()
    masks, offs = msc_tools.mask_offsets(H.msc)
    osub = oracle_subspace(sub, L)
    omsc = oracle.Msc(masks, offs, H.msc['signs'], H.msc['coeffs'])
    y = oracle.matmult(omsc, osub, osub, x)
    diag = None
    if masks[0] == 0:
        n0 = int(offs[1])
        dmsc = oracle.Msc(masks[:1], offs[:2], H.msc['signs'][:n0], H.msc['coeffs'][:n0])
        d = oracle.matmult(dmsc, osub, osub, np.ones(x.size, dtype=np.complex128))
        assert np.all(d.imag == 0)
        diag = np.ascontiguousarray(d.real)
    return y, diag


def random_state(n, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)


def check(tmp_path, name, L, sub=None, nranks=1, seed=0, **kw):
    sub = sub if sub is not None else Full(L=L)
    dim = sub.get_dimension()
    x = random_state(dim, seed)
    y_ref, diag = reference_product(name, L, sub, x)
    nloc = dim // nranks
    scale = np.abs(y_ref).max()
    infos = []
    for rank in range(nranks):
        label = name if isinstance(name, str) else 'op'
        emu = Emulated(tmp_path, name, L, f'{label}_{L}_{nranks}_{rank}', sub=sub, nranks=nranks, rank=rank, **kw)
        shards = [np.ascontiguousarray(x[(rank ^ h) * nloc:((rank ^ h) + 1) * nloc]) for h in range(nranks)]
        d = None if diag is None else np.ascontiguousarray(diag[rank * nloc:(rank + 1) * nloc])
        y = emu.mult(shards, d)
        err = np.abs(y - y_ref[rank * nloc:(rank + 1) * nloc]).max() / scale
        assert err < 1e-14, (name, L, rank, err)       # f64 sums in a different order: rounding only
        infos.append(emu.info)
    return infos


@pytest.mark.parametrize('name', ['MBL', 'heisenberg', 'long_range', 'ising', 'XX'])
def test_generated_passes_match_oracle(tmp_path, name):
    infos = check(tmp_path, name, 14, tile_bits=9, far_bits=2)
    assert infos[0]['passes'] >= 2


@pytest.mark.parametrize('name', ['MBL', 'long_range'])
@pytest.mark.parametrize('tune', [0, 1, 2, 3])
def test_autotuner_shapes(tmp_path, name, tune):
    """The plan shapes the first-use autotuner chooses from (T = 11, 12, 13: 256..1024 threads per tile)."""
    check(tmp_path, name, 18, tune=tune)


def test_rank_4_boxes_and_three_passes(tmp_path):
    info = check(tmp_path, 'heisenberg', 20, tune=3)[0]
    assert info['passes'] == 3 and 'const int c[4]' in info['src']


def test_far_masks_and_tma_boxes(tmp_path):
    """The default L=30 shape in miniature: a writing pass on the contiguous tile and an accumulating
    pass whose window is a tensor box, the masks that leave the window served as FAR loads."""
    info = check(tmp_path, 'MBL', 15, tile_bits=9, far_bits=3)[0]
    assert ' FAR' in info['src'] and 'reduce_p' in info['src'] and 'load_p' in info['src']
    info = check(tmp_path, 'MBL', 15, tile_bits=10, far_bits=0)[0]
    assert ' FAR' not in info['src']


def test_cp_async_staging_and_plain_accumulate(tmp_path, monkeypatch):
    monkeypatch.setenv('DNM_JIT_TMA', '0')
    info = check(tmp_path, 'heisenberg', 14, tile_bits=9, far_bits=2)[0]
    assert 'load_p' not in info['src'] and 'cpa16(&tile' in info['src']
    monkeypatch.delenv('DNM_JIT_TMA')
    monkeypatch.setenv('DNM_JIT_TMA_REDUCE', '0')
    info = check(tmp_path, 'heisenberg', 14, tile_bits=9, far_bits=2)[0]
    assert 'load_p' in info['src'] and 'reduce_p' not in info['src']


def test_pipelined_ring(tmp_path):
    """Persistent CTAs: 3 CTAs walk 32 tiles, every ring slot and both barrier phases are reused."""
    info = check(tmp_path, 'MBL', 14, tile_bits=9, far_bits=2, pipeline=1)[0]
    assert info['pipelined'] >= 1 and 'mbar_wait(&full[b], phase)' in info['src']


def test_parity_subspace(tmp_path):
    check(tmp_path, 'heisenberg', 15, sub=Parity('even', L=15), tile_bits=9, far_bits=2)
    check(tmp_path, 'MBL', 15, sub=Parity('odd', L=15), tile_bits=9, far_bits=2)


@pytest.mark.parametrize('name,nranks', [('MBL', 2), ('long_range', 4), ('heisenberg', 4), ('XX', 2)])
def test_folded_remote_masks_read_the_peer_shards(tmp_path, name, nranks):
    """Sharded MatMult in fold mode: every rank's passes read the partner shards through Peers."""
    L = 13 + nranks.bit_length() - 1
    infos = check(tmp_path, name, L, nranks=nranks, tile_bits=9, far_bits=2)
    for info in infos:
        assert info['remote'] >= 1 and 'xs.p[' in info['src']


def test_emulation_switch_does_not_leak(tmp_path):
    """After the switch is cleared the generator emits the CUDA source again, byte for byte."""
    a = dryrun('MBL', 24)
    Emulated(tmp_path, 'MBL', 14, 'leak', tile_bits=9, far_bits=2)
    b = dryrun('MBL', 24)
    assert a['src'] == b['src'] and b['cubin'] > 0
    assert not re.search(r'emu_|pthread', b['src'])


@pytest.mark.parametrize('knob', ['DNM_REMOTE_STAGE=1', 'DNM_FOLD_SPLIT=1', 'DNM_JIT_FAR_FIRST=1', 'DNM_JIT_HINTS=0',
                                  'DNM_JIT_ROWS=4'])
def test_opt_in_variants_compute_the_same_product(tmp_path, monkeypatch, knob):
    """The experiment knobs of INTEGRATION.md section 6 change the schedule, never the result."""
    key, val = knob.split('=')
    monkeypatch.setenv(key, val)
    check(tmp_path, 'long_range', 15, nranks=4, tile_bits=9, far_bits=2)
    check(tmp_path, 'MBL', 15, tile_bits=9, far_bits=3)


def test_c5_in_miniature_long_range_on_8_ranks(tmp_path):
    """BASELINE config C5 (long_range, 8 GPUs) scaled down to 2^13 rows per rank: 6 cross-rank masks, 7 partners."""
    infos = check(tmp_path, 'long_range', 16, nranks=8, tile_bits=9, far_bits=2)
    assert all(i['remote'] >= 6 for i in infos)


def random_pauli_operator(L, seed, nstrings=14):
    """A random Hermitian sum of Pauli strings on up to three sites each (real coefficients)."""
    from dynamite_b200.operators import sigmax, sigmay, sigmaz
    rng = np.random.default_rng(seed)
    H = None
    for _ in range(nstrings):
        sites = rng.choice(L, size=int(rng.integers(1, 4)), replace=False)
        term = None
        for site in sites:
            p = (sigmax, sigmay, sigmaz)[int(rng.integers(0, 3))](int(site))
            term = p if term is None else term * p
        term = float(rng.uniform(-1.5, 1.5)) * term
        H = term if H is None else H + term
    H.L = L
    return H


@pytest.mark.parametrize('seed', range(8))
def test_random_pauli_operators(tmp_path, seed):
    """Operators the fixed models do not reach: flips on arbitrary site triples, imaginary (sigma_y) coefficients,
    masks with one, two and three sign patterns -- on one rank and folded over two."""
    L = 14
    rng = np.random.default_rng(100 + seed)
    H = random_pauli_operator(L, seed)
    kw = dict(tile_bits=int(rng.choice([9, 10])), far_bits=int(rng.integers(0, 4)))
    infos = check(tmp_path, H, L, seed=seed, **kw)
    assert infos[0]['kernels'] == infos[0]['passes']
    check(tmp_path, random_pauli_operator(L, seed), L, nranks=2, seed=seed, **kw)


def random_pair_operator(L, seed, ngroups=12):
    """Masks that carry TWO sign patterns each, in every combination the generator has a case for: XX + YY
    with equal and unequal weights, XY +/- YX (imaginary coefficients), a field plus a conditional field."""
    from dynamite_b200.operators import sigmax, sigmay, sigmaz
    rng = np.random.default_rng(seed)
    H = None
    used = set()
    while len(used) < ngroups:
        i, j, k = (int(v) for v in rng.choice(L, size=3, replace=False))
        a, b = (float(v) for v in rng.uniform(-1.0, 1.0, size=2))
        kind = int(rng.integers(0, 5))
        mask = (1 << i) | (1 << j) if kind < 3 else (1 << i)
        if mask in used:                       # a third sign pattern on one mask would leave the lean form
            continue
        used.add(mask)
        if kind == 0:
            g = a * sigmax(i) * sigmax(j) + a * sigmay(i) * sigmay(j)          # half of the rows vanish
        elif kind == 1:
            g = a * sigmax(i) * sigmax(j) + b * sigmay(i) * sigmay(j)
        elif kind == 2:
            g = a * sigmax(i) * sigmay(j) - a * sigmay(i) * sigmax(j)
        elif kind == 3:
            g = a * sigmax(i) + b * sigmax(i) * sigmaz(k)
        else:
            g = a * sigmay(i) * sigmaz(j) + b * sigmay(i) * sigmaz(k) + 0.3 * sigmaz(j) * sigmaz(k)
        H = g if H is None else H + g
    H.L = L
    return H


@pytest.mark.parametrize('seed', range(6))
def test_random_two_pattern_operators(tmp_path, seed):
    L = 14
    rng = np.random.default_rng(200 + seed)
    kw = dict(tile_bits=int(rng.choice([9, 10])), far_bits=int(rng.integers(0, 4)))
    try:
        check(tmp_path, random_pair_operator(L, seed), L, seed=seed, **kw)
        check(tmp_path, random_pair_operator(L, seed), L, nranks=2, seed=seed, **kw)
    except NotGenerated:
        pytest.skip('groups collided into a mask with more than two sign patterns: a pass keeps the generic kernel')


def test_contiguous_accumulating_pass(tmp_path):
    """Found by fuzzing this file's checks: when the window of an ACCUMULATING pass is the contiguous tile there is
    no tensor box for the reduce-add epilogue; the generator used to emit a 0-d tensor reduce that NVRTC rejects
    (the plan then silently lost its generated kernels).  Such a pass keeps the read-modify-write epilogue."""
    H = random_pauli_operator(16, 1022, nstrings=20)
    cuda = dryrun(H, 16, tile_bits=13, far_bits=0)
    assert cuda['cubin'] > 0 and cuda['kernels'] == cuda['passes'] == 2
    assert 'accumulate' in cuda['src'] and 'cp.reduce.async.bulk.tensor' not in cuda['src']
    check(tmp_path, random_pauli_operator(16, 1022, nstrings=20), 16, seed=1, tile_bits=13, far_bits=0)
    piped = dryrun(random_pauli_operator(16, 1022, nstrings=20), 16, tile_bits=13, far_bits=0, pipeline=1)
    assert piped['cubin'] > 0           # (the persistent variant leaves that pass to the classic kernel)


def test_the_default_L30_shape_with_its_full_L2_window(tmp_path):
    """(T, run bits, far) = (11, 4, 10), the shape the autotuner keeps for the L=30 MBL benchmark: at L=22 the
    writing pass has the same 20 groups, half of them FAR loads over ten index positions."""
    info = check(tmp_path, 'MBL', 22, tune=0)[0]
    assert 'far_bits=10' in info['src'] and info['src'].count(' FAR') >= 10
