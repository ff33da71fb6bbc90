"""CPU: pin the oracle against the reference's own KATs and the fixtures generated
from the reference's msc_to_numpy (tests/golden/generate_golden.py)."""
import numpy as np
import pytest

import oracle
from helpers import case_terms, golden_cases, kats, rand_state, rel_err

CASES = golden_cases()


def osub(spec):
    return oracle.Subspace(spec)


def test_parity_kats():
    k = kats()
    for space in (0, 1):
        want = np.array(k['parity_L4'][str(space)])
        s = osub({'type': 'parity', 'L': 4, 'space': space})
        assert s.dim == 8
        assert np.array_equal(s.i2s(np.arange(8)), want)
        assert np.array_equal(s.s2i(want), np.arange(8))
    s = osub({'type': 'parity', 'L': 5, 'space': 0})
    for idx, st in k['parity_L5_even_single']:
        assert s.i2s(idx)[0] == st and s.s2i(st)[0] == idx
    for space, st in k['parity_L5_invalid']:
        assert osub({'type': 'parity', 'L': 5, 'space': space}).s2i(st)[0] == -1


def test_spinconserve_kats():
    k = kats()
    for L, kk, dim in k['spinconserve_dims']:
        assert osub({'type': 'spinconserve', 'L': L, 'k': kk}).dim == dim
    idx, st = k['spinconserve_L6_k3_single']
    s = osub({'type': 'spinconserve', 'L': 6, 'k': 3})
    assert s.i2s(idx)[0] == st and s.s2i(st)[0] == idx
    for L, kk, st in k['spinconserve_invalid']:
        assert osub({'type': 'spinconserve', 'L': L, 'k': kk}).s2i(st)[0] == -1
    for kk, want in k['spinconserve_L4'].items():
        s = osub({'type': 'spinconserve', 'L': 4, 'k': int(kk)})
        want = np.array(want)
        assert np.array_equal(s.i2s(np.arange(want.size)), want)
        assert np.array_equal(s.s2i(want), np.arange(want.size))
        assert s.s2i(0b0111)[0] == -1


@pytest.mark.parametrize('spec', [
    {'type': 'full', 'L': 7}, {'type': 'parity', 'L': 9, 'space': 0}, {'type': 'parity', 'L': 9, 'space': 1},
    {'type': 'spinconserve', 'L': 12, 'k': 5}, {'type': 'spinconserve', 'L': 10, 'k': 0},
    {'type': 'spinconserve', 'L': 10, 'k': 10}, {'type': 'spinconserve', 'L': 11, 'k': 1},
])
def test_maps_match_brute_force_definition(spec):
    states = oracle.brute_states(spec)
    s = osub(spec)
    assert s.dim == states.size
    assert np.array_equal(s.i2s(np.arange(states.size)), states)
    allst = np.arange(1 << spec['L'])
    want = np.full(allst.size, -1)
    want[states] = np.arange(states.size)
    assert np.array_equal(s.s2i(allst), want)
    # NextState steps through the same sequence (bsubspace_impl.h:230-245)
    st = int(states[0])
    for i in range(1, min(states.size, 500)):
        st = s.next_state(st, i)
        assert st == states[i]


def test_explicit_sorted_and_unsorted():
    R = np.random.RandomState(3)
    states = np.sort(R.choice(1 << 10, size=200, replace=False))
    shuffled = states.copy()
    R.shuffle(shuffled)
    for lst in (states, shuffled):
        s = osub({'type': 'explicit', 'L': 10, 'states': lst.tolist()})
        assert s.dim == 200
        assert np.array_equal(s.i2s(np.arange(200)), lst)
        assert np.array_equal(s.s2i(lst), np.arange(200))
        missing = np.setdiff1d(np.arange(1 << 10), states)
        assert np.all(s.s2i(missing) == -1)


def test_msc_definition_kat():
    k = kats()
    terms = [(m, s, complex(*c)) for m, s, c in k['msc_full_terms']]
    want = np.array([[complex(*v) for v in row] for row in k['msc_full_dense']])
    # numpy restatement of the definition
    got = oracle.msc_to_dense(terms, np.arange(8), np.arange(8))
    assert np.array_equal(got, want)
    # hand-typed golden rows of tests/unit/test_msc_tools.py:142-172
    assert want[0, 1] == -0.5j and want[0, 4] == -2 and want[1, 0] == 0.5j and want[7, 6] == -0.5j
    # and the C matmult reproduces it column by column
    msc = oracle.Msc.from_terms(sorted(terms))
    sub = osub({'type': 'full', 'L': 3})
    for j in range(8):
        e = np.zeros(8, complex)
        e[j] = 1
        assert np.allclose(oracle.matmult(msc, sub, sub, e), want[:, j], atol=0)


@pytest.mark.parametrize('tag', sorted(CASES))
def test_oracle_vs_reference_cases(tag):
    c = CASES[tag]
    terms = case_terms(c)
    left, right = osub(c['left']), osub(c['right'])
    msc = oracle.Msc.from_terms(terms)
    xp = c['xparity']
    y = oracle.matmult(msc, left, right, c['x'], xparity=xp)
    assert rel_err(y, c['y']) < 1e-13
    assert abs(oracle.norm_inf(msc, left, right, xparity=xp) - float(c['norm_inf'])) < 1e-12 * max(1, float(c['norm_inf']))
    if 'diag' in c and msc.masks[0] == 0 and c['left'] == c['right']:
        d = oracle.precompute_diag(msc, right, xparity=xp)
        assert np.allclose(d, c['diag'], atol=1e-13)
        y2 = oracle.matmult(msc, left, right, c['x'], xparity=xp, diag=d)
        assert rel_err(y2, c['y']) < 1e-13
    # dense restatement agrees with the reference's dense matrix
    if not xp:
        A = oracle.msc_to_dense(terms, oracle.brute_states(c['left']), oracle.brute_states(c['right']))
        assert np.allclose(A, c['A'], atol=1e-15)


@pytest.mark.parametrize('tag', ['heisenberg_L6_full', 'SYK_L5_par0', 'SYK_L5_par1', 'heisenberg_L8_par0', 'MBL_L8_full'])
def test_fast_path_matches_general(tag):
    # the fast path needs dim >= 2048: lift the small operators to a longer chain
    c = CASES[tag]
    terms = case_terms(c)
    L = 14
    spec = dict(c['left'])
    spec['L'] = L
    sub = osub(spec)
    msc = oracle.Msc.from_terms(terms)
    x = rand_state(sub.dim, 5)
    want = oracle.matmult(msc, sub, sub, x)
    for nthreads in (1, 3):
        got, used = oracle.matmult_fast(msc, sub, x, nthreads=nthreads)
        assert used == nthreads
        assert rel_err(got, want) < 1e-13
    if msc.masks[0] == 0:
        d = oracle.precompute_diag(msc, sub)
        got, _ = oracle.matmult_fast(msc, sub, x, diag=d, nthreads=2)
        assert rel_err(got, want) < 1e-13


def test_check_conserves():
    c = CASES['heisenberg_L8_sc4']
    msc = oracle.Msc.from_terms(case_terms(c))
    sc4, sc3 = osub({'type': 'spinconserve', 'L': 8, 'k': 4}), osub({'type': 'spinconserve', 'L': 8, 'k': 3})
    full = osub({'type': 'full', 'L': 8})
    assert oracle.check_conserves(msc, sc4, sc4)
    assert oracle.check_conserves(msc, full, sc4)
    assert not oracle.check_conserves(msc, sc3, sc4)
    lr = oracle.Msc.from_terms(case_terms(CASES['long_range_L7_full']))
    p0 = osub({'type': 'parity', 'L': 7, 'space': 0})
    assert not oracle.check_conserves(lr, p0, p0)       # sigma_x fields flip parity
    assert oracle.check_conserves(oracle.Msc.from_terms(case_terms(CASES['SYK_L5_par0'])),
                                  osub({'type': 'parity', 'L': 5, 'space': 0}), osub({'type': 'parity', 'L': 5, 'space': 0}))


def test_rdm_kats():
    k = kats()
    psi = np.array([complex(*v) for v in k['rdm_L4_state']])
    full4 = osub({'type': 'full', 'L': 4})
    for keep, key in (([0], 'rdm_L4_keep0'), ([2], 'rdm_L4_keep2')):
        want = np.array([[complex(*v) for v in row] for row in k[key]])
        got = oracle.rdm(psi, full4, keep)
        assert np.allclose(got, want, atol=2e-6)   # golden is printed to 6 digits
        assert np.allclose(oracle.rdm_dense(psi, 4, keep), got, atol=1e-15)
        w = np.linalg.eigvalsh(got)
        assert abs(-np.sum(w * np.log(w)) - k[key + '_entropy']) < 2e-5
    for keep in ([0, 2], [1, 3], [0, 1, 2, 3], []):
        assert np.allclose(oracle.rdm(psi, full4, keep), oracle.rdm_dense(psi, 4, keep), atol=1e-15)
    cs = k['rdm_complex_sign']
    got = oracle.rdm(np.array([complex(*v) for v in cs['state']]), osub({'type': 'full', 'L': 2}), cs['keep'])
    assert np.array_equal(got, np.array([[complex(*v) for v in row] for row in cs['dm']]))
    with pytest.raises(ValueError):
        oracle.rdm(psi, full4, [2, 1])


def test_rdm_subspace_skips_missing_states():
    spec = {'type': 'spinconserve', 'L': 6, 'k': 3}
    states = oracle.brute_states(spec)
    psi_sub = rand_state(states.size, 11)
    psi_full = np.zeros(64, complex)
    psi_full[states] = psi_sub
    for keep in ([0, 1], [1, 4, 5], [3]):
        assert np.allclose(oracle.rdm(psi_sub, osub(spec), keep), oracle.rdm_dense(psi_full, 6, keep), atol=1e-15)


def test_compute_rcm_finds_conserved_sector():
    c = CASES['heisenberg_L8_sc4']
    terms = case_terms(c)
    masks = [t[0] for t in terms]
    signs = [t[1] for t in terms]
    coeffs = [t[2] for t in terms]
    found = oracle.compute_rcm(masks, signs, coeffs, start=0b00001111, L=8)
    assert np.array_equal(np.sort(found), oracle.brute_states({'type': 'spinconserve', 'L': 8, 'k': 4}))
    assert found[0] == 0b00001111
    with pytest.raises(RuntimeError):
        oracle.compute_rcm(masks, signs, coeffs, start=0b00001111, L=8, max_states=10)
