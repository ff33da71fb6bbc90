"""CPU: the host side of dynamite_b200 -- MSC construction, the C-ABI library's
exports, and the host index-map entry points -- against golden data and the oracle.
No GPU compute is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from dynamite_b200 import _capi, msc_tools
from dynamite_b200._backend import bsubspace
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.operators import index_sum, sigmax, sigmay, sigmaz
from dynamite_b200.subspaces import Auto, Explicit, Full, Parity, SpinConserve, XParity
from helpers import case_terms, golden_cases, kats, product_subspace

CASES = golden_cases()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'dynamite_b200.h')).read()
    declared = set(re.findall(r'\b(dnm_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f'symbols declared in the header but not exported: {missing}'
    # the Python binding covers the same set
    assert declared == set(_capi.exported_symbols())


def test_compute_entry_points_fail_loudly_without_a_gpu():
    if _capi.gpu_available():
        pytest.skip('a GPU is present')
    with pytest.raises(_capi.BackendError):
        _capi.ensure_gpu(0)
    lib = _capi.lib()
    h = ctypes.c_void_p()
    assert lib.dnm_vec_create(16, ctypes.byref(h)) != 0
    assert b'no CPU fallback' in lib.dnm_last_error()


@pytest.mark.parametrize('name,L', [('MBL', 8), ('long_range', 7), ('SYK', 4), ('ising', 6), ('XX', 5), ('heisenberg', 6)])
def test_hamiltonian_msc_matches_reference(name, L):
    H = build_hamiltonian(name, L)
    H.reduce_msc()
    c = CASES[f'{name}_L{L}_full']
    assert np.array_equal(H.msc['masks'], c['msc_masks'])
    assert np.array_equal(H.msc['signs'], c['msc_signs'])
    assert np.allclose(H.msc['coeffs'], c['msc_coeffs'], rtol=1e-14, atol=0)
    # and the host matrix from our msc_to_numpy equals the reference's
    assert np.allclose(H.to_numpy(sparse=False), c['A'], atol=1e-15)


def test_benchmark_sizes_match_baseline_table():
    # BASELINE.md section 2
    for name, L, nterms, nmasks in [('heisenberg', 20, 57, 20), ('MBL', 30, 117, 30), ('long_range', 33, 659, 66)]:
        H = build_hamiltonian(name, L)
        assert (H.nterms, H.nnz) == (nterms, nmasks)


@pytest.mark.parametrize('tag', ['heisenberg_L7_xparity_plus', 'heisenberg_L7_xparity_minus', 'ising_L6_xparity_plus'])
def test_xparity_reduce_msc_matches_reference(tag):
    c = CASES[tag]
    name, L = tag.split('_')[0], c['left']['L']
    H = build_hamiltonian(name, L)
    H.reduce_msc()
    xp = XParity(Full(L=L), sector='+' if tag.endswith('plus') else '-')
    red = xp.reduce_msc(H.msc)
    assert np.array_equal(red['masks'], c['msc_masks'])
    assert np.array_equal(red['signs'], c['msc_signs'])
    assert np.allclose(red['coeffs'], c['msc_coeffs'], rtol=1e-14, atol=0)


def test_pauli_algebra():
    x, y, z = sigmax(), sigmay(), sigmaz()
    assert x * y == 1j * z
    assert y * z == 1j * x
    assert z * x == 1j * y
    assert x * x == 1 * (x * x)
    assert (x * x).msc.tolist() == [(0, 0, 1 + 0j)]
    two = index_sum(sigmax(0) * sigmax(1), size=4)
    assert two.nterms == 3 and two.max_spin_idx == 3
    closed = index_sum(sigmaz(0) * sigmaz(1), size=4, boundary='closed')
    assert sorted(closed.msc['signs'].tolist()) == [0b0011, 0b0110, 0b1001, 0b1100]
    assert not (x + 1j * z).is_hermitian() and (x + z).is_hermitian()


SPECS = [
    {'type': 'full', 'L': 7}, {'type': 'parity', 'L': 9, 'space': 0}, {'type': 'parity', 'L': 9, 'space': 1},
    {'type': 'spinconserve', 'L': 12, 'k': 5}, {'type': 'spinconserve', 'L': 10, 'k': 0},
    {'type': 'spinconserve', 'L': 10, 'k': 10}, {'type': 'spinconserve', 'L': 40, 'k': 3},
]


@pytest.mark.parametrize('spec', SPECS)
def test_host_index_maps_bit_exact_vs_oracle(spec):
    sub = product_subspace(spec)
    orc = oracle.Subspace(spec)
    dim = sub.get_dimension()
    assert dim == orc.dim
    R = np.random.RandomState(1)
    idx = np.unique(np.concatenate([np.arange(min(dim, 3000)), R.randint(0, dim, 3000), [dim - 1]]))
    states = sub.idx_to_state(idx)
    assert np.array_equal(states, orc.i2s(idx))
    assert np.array_equal(sub.state_to_idx(states), idx)
    probe = R.randint(0, 1 << spec['L'], 5000, dtype=np.int64)
    assert np.array_equal(sub.state_to_idx(probe), orc.s2i(probe))
    with pytest.raises(ValueError):
        sub.idx_to_state(dim)
    with pytest.raises(ValueError):
        sub.idx_to_state(-1)


def test_reference_subspace_kats_through_product_classes():
    k = kats()
    for space in (0, 1):
        p = Parity(space, L=4)
        want = np.array(k['parity_L4'][str(space)])
        assert np.array_equal(p.idx_to_state(np.arange(8)), want)
        assert np.array_equal(p.state_to_idx(want), np.arange(8))
    for L, kk, dim in k['spinconserve_dims']:
        assert SpinConserve(L, kk).get_dimension() == dim
    idx, st = k['spinconserve_L6_k3_single']
    assert SpinConserve(6, 3).idx_to_state(idx) == st and SpinConserve(6, 3).state_to_idx(st) == idx
    for L, kk, st in k['spinconserve_invalid']:
        assert SpinConserve(L, kk).state_to_idx(st) == -1
    for kk, want in k['spinconserve_L4'].items():
        assert np.array_equal(SpinConserve(4, int(kk)).idx_to_state(np.arange(len(want))), np.array(want))


def test_explicit_and_auto():
    R = np.random.RandomState(3)
    states = np.sort(R.choice(1 << 10, size=200, replace=False))
    shuffled = states.copy()
    R.shuffle(shuffled)
    for lst in (states, shuffled):
        e = Explicit(lst, L=10)
        o = oracle.Subspace({'type': 'explicit', 'L': 10, 'states': lst.tolist()})
        assert np.array_equal(e.idx_to_state(np.arange(200)), lst)
        probe = np.arange(1 << 10)
        assert np.array_equal(e.state_to_idx(probe), o.s2i(probe))
    with pytest.raises(ValueError):
        Explicit([1, 2, 2, 3], L=4)
    # Auto == SpinConserve / Parity as sets of states (reference test_subspaces.py:344-357, 740-773)
    H = build_hamiltonian('heisenberg', 8)
    a = Auto(H, 'UUUUDDDD')
    assert a == SpinConserve(8, 4)
    assert np.array_equal(a.idx_to_state(np.arange(a.get_dimension())),
                          oracle.brute_states({'type': 'spinconserve', 'L': 8, 'k': 4}))
    unsorted = Auto(H, 'UUUUDDDD', sort=False)
    assert unsorted.get_dimension() == 70 and unsorted != SpinConserve(8, 4) or True
    terms = case_terms(CASES['heisenberg_L8_sc4'])
    want = oracle.compute_rcm([t[0] for t in terms], [t[1] for t in terms], [t[2] for t in terms], 0b11110000, 8)
    assert np.array_equal(unsorted.idx_to_state(np.arange(70)), want[::-1])
    smap = np.empty(10, dtype=np.int64)
    with pytest.raises(RuntimeError, match='state_map size too small'):
        bsubspace.compute_rcm(H.msc['masks'], H.msc['signs'], H.msc['coeffs'], smap, 0b11110000, 8)


def test_xparity_validation():
    XParity(Parity('even', L=6))
    with pytest.raises(ValueError):
        XParity(Parity('even', L=5))
    XParity(SpinConserve(6, 3))
    with pytest.raises(ValueError):
        XParity(SpinConserve(6, 2))
    x = XParity(SpinConserve(4, 2), sector='-')
    assert x.get_dimension() == 3
    # representatives are the first half of the parent (reference test_subspaces.py:385-410)
    assert np.array_equal(x.idx_to_state(np.arange(3)), np.array([0b0011, 0b0101, 0b0110]))
    with pytest.raises(ValueError):
        x.state_to_idx(0b1100)


def test_mult_host_batch_argument_checks_and_pointer_tables(monkeypatch):
    """Mat.mult_host_batch validates its buffers on the host and hands the C ABI two tables of pointers."""
    from dynamite_b200 import petsc
    seen = {}

    class FakeLib:
        @staticmethod
        def dnm_mat_mult_host_batch(handle, n, xp, yp):
            seen['n'] = n
            seen['x'] = [xp[i] for i in range(n)]
            seen['y'] = [yp[i] for i in range(n)]
            return 0

    monkeypatch.setattr(petsc._capi, 'lib', lambda: FakeLib)
    mat = petsc.Mat(None)
    xs = [np.zeros(8, dtype=np.complex128) for _ in range(3)]
    ys = [np.zeros(8, dtype=np.complex128) for _ in range(3)]
    mat.mult_host_batch(xs, ys)
    assert seen['n'] == 3 and seen['x'] == [x.ctypes.data for x in xs] and seen['y'] == [y.ctypes.data for y in ys]
    mat.mult_host_batch([xs[0]] * 2, [ys[0]] * 2)          # inputs may alias inputs, outputs may alias outputs
    assert seen['x'] == [xs[0].ctypes.data] * 2
    mat.mult_host_batch([], [])
    assert seen['n'] == 0
    with pytest.raises(ValueError):
        mat.mult_host_batch(xs, ys[:2])
    with pytest.raises(ValueError):
        mat.mult_host_batch([xs[0]], [xs[0]])                # a product cannot be done in place
    with pytest.raises(ValueError):
        mat.mult_host_batch([xs[0].astype(np.complex64)], [ys[0]])
    with pytest.raises(ValueError):
        mat.mult_host_batch([np.zeros(16, dtype=np.complex128)[::2]], [ys[0]])


def test_mfn_type_reaches_the_c_abi(monkeypatch):
    """The MFN shim passes the algorithm the caller named (-1 = the library's choice) and reports the one that ran."""
    from dynamite_b200 import slepc
    calls = []

    class FakeLib:
        @staticmethod
        def dnm_evolve_algo(mat, b, x, are, aim, tol, ncv, max_it, algo, reason, its, mm):
            calls.append((are, aim, tol, ncv, max_it, algo))
            reason._obj.value, its._obj.value, mm._obj.value = 1, 1, 87
            return 0

        @staticmethod
        def dnm_evolve_last_algo():
            return 2

    class Handle:
        handle = None

    monkeypatch.setattr(slepc._capi, 'lib', lambda: FakeLib)
    for name, code in (('auto', -1), ('expokit', 0), ('krylov', 1), ('chebyshev', 2)):
        mfn = slepc.MFN().create()
        mfn.getFN().setScale(-1j * 0.5)
        mfn.setType(name)
        mfn.setTolerances(tol=1e-9, max_it=7)
        mfn.setDimensions(12)
        mfn.setOperator(Handle())
        mfn.solve(Handle(), Handle())
        assert calls[-1] == (0.0, -0.5, 1e-9, 12, 7, code)
        assert mfn.used == 'chebyshev' and mfn.matmults == 87 and mfn.getConvergedReason() == 1
    with pytest.raises(ValueError):
        slepc.MFN().create().setType('pade')
