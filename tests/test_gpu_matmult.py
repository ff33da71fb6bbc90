"""GPU parity: the CUDA MatMult path through the C ABI against the golden fixtures
(reference msc_to_numpy) and the oracle.  Tolerance: 1e-12 relative (north star)."""
import numpy as np
import pytest

import oracle
from helpers import (case_terms, device_mult, golden_cases, product_mat, product_subspace,
                     rand_state, rel_err)

pytestmark = pytest.mark.gpu
CASES = golden_cases()
TOL = 1e-12


@pytest.mark.parametrize('tag', sorted(CASES))
def test_matmult_vs_reference_cases(gpu, tag):
    c = CASES[tag]
    terms = case_terms(c)
    for diag in (False, True):
        if diag and not (c['left'] == c['right'] and terms[0][0] == 0):
            continue
        mat = product_mat(terms, c['left'], c['right'], c['xparity'], precompute_diag=diag)
        assert mat.getSize() == c['A'].shape
        y = device_mult(mat, c['x'])
        assert rel_err(y, c['y']) < TOL, (tag, diag, mat.get_info('kernel'))
        assert abs(mat.norm() - float(c['norm_inf'])) < 1e-12 * max(1.0, float(c['norm_inf']))
        mat.destroy()


@pytest.mark.parametrize('tag', ['heisenberg_L6_full', 'long_range_L7_full', 'SYK_L5_par0', 'SYK_L5_par1',
                                 'heisenberg_L8_par0', 'MBL_L8_full', 'ising_L6_full', 'heisenberg_L7_xparity_minus'])
@pytest.mark.parametrize('L', [10, 13])
@pytest.mark.parametrize('tile_bits', [8, 9, 10, 11, 12, 13])
def test_tiled_kernel_vs_oracle(gpu, tag, L, tile_bits):
    """Lift small operators onto longer chains so the window planner needs several
    passes; compare the tiled kernel, the general kernel and the oracle."""
    c = CASES[tag]
    terms = case_terms(c)
    spec = dict(c['left'])
    if c['xparity']:
        pytest.skip('xparity operators are tied to their L')
    spec['L'] = L
    sub = oracle.Subspace(spec)
    msc = oracle.Msc.from_terms(terms)
    x = rand_state(sub.dim, 99)
    want = oracle.matmult(msc, sub, sub, x)
    for diag in (False, True):
        mat = product_mat(terms, spec, spec, False, precompute_diag=diag)
        mat.set_option('tile_bits', tile_bits)
        mat.set_option('kernel', 2)
        y = device_mult(mat, x)
        assert mat.get_info('kernel') == 2
        assert rel_err(y, want) < TOL, (tag, L, tile_bits, diag, mat.get_info('passes'))
        if tile_bits >= 10:
            mat.set_option('tile_rows', 16)      # 16 rows per thread
            assert rel_err(device_mult(mat, x), want) < TOL, (tag, L, tile_bits, diag, 'rows=16')
            mat.set_option('tile_rows', 0)
        for far in (1, 3):                       # masks that leave the window served through the L2 window
            mat.set_option('far_bits', far)
            assert rel_err(device_mult(mat, x), want) < TOL, (tag, L, tile_bits, diag, 'far_bits', far)
        mat.set_option('far_bits', 0)
        mat.set_option('kernel', 1)
        y1 = device_mult(mat, x)
        assert mat.get_info('kernel') == 1
        assert rel_err(y1, want) < TOL
        mat.destroy()


@pytest.mark.parametrize('name,L', [('MBL', 20), ('heisenberg', 19), ('long_range', 18), ('ising', 18), ('XX', 19)])
def test_far_masks_l2_window(gpu, name, L):
    """FAR masks: a pass serves masks that leave its shared-memory window from global memory (the
    L2 window = the lowest tile-number bits).  Every (tile size, far_bits) combination changes which
    masks are far and the tile order; all are checked against the oracle's fast path."""
    from dynamite_b200.hamiltonians import build_hamiltonian
    H = build_hamiltonian(name, L)
    H.reduce_msc()
    terms = [(int(m), int(s), complex(c)) for m, s, c in zip(H.msc['masks'], H.msc['signs'], H.msc['coeffs'])]
    spec = {'type': 'full', 'L': L}
    sub = oracle.Subspace(spec)
    x = rand_state(sub.dim, 5)
    want, _ = oracle.matmult_fast(oracle.Msc.from_terms(terms), sub, x, nthreads=8)
    for diag in (True, False):
        mat = product_mat(terms, spec, spec, False, precompute_diag=diag)
        mat.set_option('kernel', 2)
        for tile_bits in (9, 10, 11, 12, 13):
            mat.set_option('tile_bits', tile_bits)
            for far in (0, 2, 5, 8, 12):
                mat.set_option('far_bits', far)
                assert rel_err(device_mult(mat, x), want) < TOL, (name, tile_bits, diag, far)
        mat.destroy()


def test_far_masks_parity_subspace(gpu):
    from dynamite_b200.hamiltonians import build_hamiltonian
    L = 18
    H = build_hamiltonian('heisenberg', L)
    H.reduce_msc()
    terms = [(int(m), int(s), complex(c)) for m, s, c in zip(H.msc['masks'], H.msc['signs'], H.msc['coeffs'])]
    for space in (0, 1):
        spec = {'type': 'parity', 'L': L, 'space': space}
        sub = oracle.Subspace(spec)
        x = rand_state(sub.dim, 7)
        want, _ = oracle.matmult_fast(oracle.Msc.from_terms(terms), sub, x, nthreads=8)
        mat = product_mat(terms, spec, spec, False, precompute_diag=True)
        mat.set_option('kernel', 2)
        for tile_bits, far in ((10, 4), (11, 6), (12, 3)):
            mat.set_option('tile_bits', tile_bits)
            mat.set_option('far_bits', far)
            assert rel_err(device_mult(mat, x), want) < TOL, (space, tile_bits, far)
        mat.destroy()


@pytest.mark.parametrize('name,L,sub', [('MBL', 20, 'full'), ('heisenberg', 19, 'full'), ('long_range', 18, 'full'),
                                        ('ising', 18, 'full'), ('XX', 19, 'full'), ('heisenberg', 19, 'parity0'),
                                        ('long_range', 18, 'parity1'), ('XX', 18, 'parity1')])
def test_generated_kernels_vs_oracle(gpu, name, L, sub):
    """Operator-specialised (NVRTC) pass kernels, forced on at small sizes: the one-tile-per-CTA shape
    and the persistent TMA + mbarrier + reduce-add shape, for several tile sizes, L2 windows and run
    lengths, against the oracle's fast path; 'jit_passes' proves the generated kernels ran."""
    import os
    from dynamite_b200.hamiltonians import build_hamiltonian
    H = build_hamiltonian(name, L)
    H.reduce_msc()
    terms = [(int(m), int(s), complex(c)) for m, s, c in zip(H.msc['masks'], H.msc['signs'], H.msc['coeffs'])]
    spec = {'type': 'full', 'L': L} if sub == 'full' else {'type': 'parity', 'L': L, 'space': int(sub[-1])}
    osub = oracle.Subspace(spec)
    x = rand_state(osub.dim, 11)
    want, _ = oracle.matmult_fast(oracle.Msc.from_terms(terms), osub, x, nthreads=8)
    for diag in (True, False):
        if not diag and name in ('MBL', 'heisenberg', 'long_range', 'ising'):
            continue    # the many-term diagonal group is not a lean pass: generic kernel (tested elsewhere)
        mat = product_mat(terms, spec, spec, False, precompute_diag=diag)
        mat.set_option('kernel', 2)
        mat.set_option('jit', 1)
        for run_bits in ('3', '4'):
            os.environ['DNM_TILE_RUN_BITS'] = run_bits
            try:
                for tile_bits, far in ((9, 3), (10, 0), (11, 4), (11, 9), (12, 5), (13, 2)):
                    mat.set_option('tile_bits', tile_bits)
                    mat.set_option('far_bits', far)
                    for pipeline in (0, 1):
                        mat.set_option('pipeline', pipeline)
                        y = device_mult(mat, x)
                        assert mat.get_info('jit_passes') >= 1, (name, tile_bits, far, pipeline)
                        assert rel_err(y, want) < TOL, (name, sub, diag, run_bits, tile_bits, far, pipeline)
            finally:
                del os.environ['DNM_TILE_RUN_BITS']
        # the default plan with generated kernels on
        mat.set_option('tile_bits', 0)
        mat.set_option('far_bits', -1)
        mat.set_option('pipeline', 0)
        assert rel_err(device_mult(mat, x), want) < TOL, (name, sub, diag, 'default plan')
        assert mat.get_info('jit_passes') >= 1
        mat.destroy()


@pytest.mark.parametrize('name', ['MBL', 'long_range'])
def test_autotuned_plan(gpu, name):
    """First-use autotuner (forced on at a small size): times its plan shapes on the caller's vectors,
    keeps one, and the product is still the oracle's."""
    from dynamite_b200.hamiltonians import build_hamiltonian
    L = 20
    H = build_hamiltonian(name, L)
    H.reduce_msc()
    terms = [(int(m), int(s), complex(c)) for m, s, c in zip(H.msc['masks'], H.msc['signs'], H.msc['coeffs'])]
    spec = {'type': 'full', 'L': L}
    osub = oracle.Subspace(spec)
    x = rand_state(osub.dim, 3)
    want, _ = oracle.matmult_fast(oracle.Msc.from_terms(terms), osub, x, nthreads=8)
    mat = product_mat(terms, spec, spec, False, precompute_diag=True)
    mat.set_option('kernel', 2)
    mat.set_option('jit', 1)
    mat.set_option('autotune', 1)
    assert rel_err(device_mult(mat, x), want) < TOL
    assert mat.get_info('tuned_shape') >= 0
    assert rel_err(device_mult(mat, x), want) < TOL      # the kept plan, second use
    mat.destroy()


def test_tiled_xparity(gpu):
    for tag in ('heisenberg_L7_xparity_plus', 'heisenberg_L7_xparity_minus', 'ising_L6_xparity_plus'):
        c = CASES[tag]
        mat = product_mat(case_terms(c), c['left'], c['right'], True)
        mat.set_option('kernel', 2)
        y = device_mult(mat, c['x'])
        assert rel_err(y, c['y']) < TOL
        mat.destroy()


def test_wide_masks_fall_back_to_direct_gather(gpu):
    # a mask with more set bits than any window can hold, plus ordinary ones
    L = 14
    allx = (1 << L) - 1
    terms = sorted([(0, 0b101, 0.7), (0b11, 0, 0.5), (allx, 0, 0.25), (allx, allx, -0.25 * (-1) ** (L // 2)),
                    (1 << (L - 1) | 1, 0, 0.3)])
    spec = {'type': 'full', 'L': L}
    sub = oracle.Subspace(spec)
    x = rand_state(sub.dim, 4)
    want = oracle.matmult(oracle.Msc.from_terms(terms), sub, sub, x)
    mat = product_mat(terms, spec, spec)
    mat.set_option('tile_bits', 9)
    mat.set_option('kernel', 2)
    y = device_mult(mat, x)
    assert rel_err(y, want) < TOL
    mat.destroy()


@pytest.mark.parametrize('left,right', [
    ({'type': 'spinconserve', 'L': 14, 'k': 7}, {'type': 'spinconserve', 'L': 14, 'k': 7}),
    ({'type': 'spinconserve', 'L': 14, 'k': 5}, {'type': 'full', 'L': 14}),
    ({'type': 'full', 'L': 14}, {'type': 'spinconserve', 'L': 14, 'k': 5}),
    ({'type': 'parity', 'L': 14, 'space': 1}, {'type': 'spinconserve', 'L': 14, 'k': 7}),
    ({'type': 'parity', 'L': 14, 'space': 0}, {'type': 'parity', 'L': 14, 'space': 1}),
])
def test_general_pairs_vs_oracle(gpu, left, right):
    terms = case_terms(CASES['long_range_L7_full'])   # has parity-changing and conserving terms
    msc = oracle.Msc.from_terms(terms)
    ol, orr = oracle.Subspace(left), oracle.Subspace(right)
    x = rand_state(orr.dim, 21)
    want = oracle.matmult(msc, ol, orr, x)
    mat = product_mat(terms, left, right)
    y = device_mult(mat, x)
    assert rel_err(y, want) < TOL
    assert abs(mat.norm() - oracle.norm_inf(msc, ol, orr)) < 1e-12 * mat.norm()
    mat.destroy()


def test_explicit_and_auto_subspaces(gpu):
    terms = case_terms(CASES['heisenberg_L8_sc4'])
    msc = oracle.Msc.from_terms(terms)
    L = 12
    states = oracle.brute_states({'type': 'spinconserve', 'L': L, 'k': 6})
    R = np.random.RandomState(5)
    shuffled = states.copy()
    R.shuffle(shuffled)
    for lst in (states, shuffled):
        spec = {'type': 'explicit', 'L': L, 'states': lst.tolist()}
        sub = oracle.Subspace(spec)
        x = rand_state(sub.dim, 8)
        want = oracle.matmult(msc, sub, sub, x)
        for diag in (False, True):
            mat = product_mat(terms, spec, spec, precompute_diag=diag)
            assert rel_err(device_mult(mat, x), want) < TOL
            mat.destroy()


def test_device_index_maps_bit_exact(gpu):
    import ctypes as C
    from dynamite_b200._capi import ip
    specs = [{'type': 'full', 'L': 20}, {'type': 'parity', 'L': 21, 'space': 0}, {'type': 'parity', 'L': 21, 'space': 1},
             {'type': 'spinconserve', 'L': 26, 'k': 13}, {'type': 'spinconserve', 'L': 40, 'k': 4},
             {'type': 'spinconserve', 'L': 12, 'k': 0}, {'type': 'spinconserve', 'L': 12, 'k': 12}]
    R = np.random.RandomState(17)
    st = np.sort(R.choice(1 << 16, size=5000, replace=False))
    sh = st.copy()
    R.shuffle(sh)
    specs += [{'type': 'explicit', 'L': 16, 'states': st.tolist()}, {'type': 'explicit', 'L': 16, 'states': sh.tolist()}]
    lib = gpu.lib()
    for spec in specs:
        orc = oracle.Subspace(spec)
        cdata = product_subspace(spec)._to_c()['data']   # keep the arrays behind the descriptor alive
        desc = cdata.desc
        dim = orc.dim
        idx = np.unique(np.concatenate([R.randint(0, dim, 20000), [0, dim - 1]])).astype(np.int64)
        states = np.empty_like(idx)
        gpu.check(lib.dnm_subspace_i2s_device(C.byref(desc), idx.size, ip(idx), ip(states)))
        assert np.array_equal(states, orc.i2s(idx)), spec
        probe = np.concatenate([states, R.randint(0, 1 << spec['L'], 20000, dtype=np.int64)])
        got = np.empty_like(probe)
        gpu.check(lib.dnm_subspace_s2i_device(C.byref(desc), probe.size, ip(probe), ip(got)))
        assert np.array_equal(got, orc.s2i(probe)), spec


def test_check_conserves_device(gpu):
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.subspaces import Full, Parity, SpinConserve, XParity
    H = build_hamiltonian('heisenberg', 10)
    assert H.conserves(SpinConserve(10, 5))
    assert H.conserves(Full(L=10), SpinConserve(10, 5))
    assert not H.conserves(SpinConserve(10, 4), SpinConserve(10, 5))
    assert H.conserves(Parity('odd', L=10))
    assert H.conserves(XParity(SpinConserve(10, 5)))
    lr = build_hamiltonian('long_range', 10)
    assert not lr.conserves(Parity('even', L=10))
    assert lr.conserves(Full(L=10))
    syk = build_hamiltonian('SYK', 5)
    assert syk.conserves(Parity('even', L=5)) and not syk.conserves(SpinConserve(5, 2))


def test_error_paths(gpu):
    from dynamite_b200._capi import BackendError
    from dynamite_b200.petsc import Vec
    spec = {'type': 'full', 'L': 6}
    with pytest.raises(BackendError, match='non-Hermitian'):
        product_mat([(1, 0, 1j)], spec, spec)
    with pytest.raises(ValueError, match='sorted'):
        product_mat([(2, 0, 1.0), (1, 0, 1.0)], spec, spec)
    mat = product_mat([(1, 0, 1.0)], spec, spec)
    x, y = Vec(64), Vec(32)
    with pytest.raises(BackendError, match='length'):
        mat.mult(x, y)
    with pytest.raises(BackendError, match='in place'):
        mat.mult(x, x)
    mat.destroy()


def test_spinconserve_rank_delta_random_masks(gpu):
    """Same-sector SpinConserve uses row + rank_delta instead of a full re-rank: stress it with
    number-conserving hops over arbitrary distances (2- and 4-site flips) and long sign strings."""
    R = np.random.RandomState(23)
    for L, k in ((16, 8), (18, 5), (20, 3), (33, 2)):
        terms = {}
        for _ in range(40):
            nflip = R.choice([2, 4])
            sites = R.choice(L, size=nflip, replace=False)
            mask = int(sum(1 << int(s) for s in sites))
            sign = int(R.randint(0, 1 << min(L, 30))) & ~mask      # sign string away from the flipped sites
            terms[(mask, sign)] = float(R.uniform(-1, 1))           # popcount(mask & sign) = 0 -> real coefficient
        terms[(0, 0b1011)] = 0.37
        tlist = sorted((m, s, c) for (m, s), c in terms.items())
        spec = {'type': 'spinconserve', 'L': L, 'k': k}
        sub = oracle.Subspace(spec)
        x = rand_state(sub.dim, 77)
        want = oracle.matmult(oracle.Msc.from_terms(tlist), sub, sub, x)
        for diag in (False, True):
            mat = product_mat(tlist, spec, spec, precompute_diag=diag)
            assert rel_err(device_mult(mat, x), want) < TOL, (L, k, diag)
            mat.destroy()


def test_wide_states_beyond_32_bits(gpu):
    """64-bit states end to end (reference tests/integration/test_matrices.py:183-232 uses L=33):
    SpinConserve and Explicit subspaces of a 40-spin chain, column by column against the oracle."""
    from dynamite_b200.hamiltonians import build_hamiltonian
    L = 40
    H = build_hamiltonian('heisenberg', L)
    H.reduce_msc()
    terms = [(int(m), int(s), complex(c)) for m, s, c in H.msc]
    msc = oracle.Msc.from_terms(terms)
    spec = {'type': 'spinconserve', 'L': L, 'k': 2}
    sub = oracle.Subspace(spec)
    assert sub.dim == 780
    states = sub.i2s(np.arange(sub.dim))
    R = np.random.RandomState(2)
    shuffled = states.copy()
    R.shuffle(shuffled)
    for sp in (spec, {'type': 'explicit', 'L': L, 'states': states.tolist()},
               {'type': 'explicit', 'L': L, 'states': shuffled.tolist()}):
        osub = oracle.Subspace(sp)
        mat = product_mat(terms, sp, sp)
        for col in (0, 1, 389, 779):
            e = np.zeros(osub.dim, complex)
            e[col] = 1
            assert np.allclose(device_mult(mat, e), oracle.matmult(msc, osub, osub, e), atol=1e-13)
        x = rand_state(osub.dim, 3)
        assert rel_err(device_mult(mat, x), oracle.matmult(msc, osub, osub, x)) < TOL
        mat.destroy()
