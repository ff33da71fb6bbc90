"""GPU parity of the Krylov consumers (evolve, eigsolve) and rdm.
Oracles: the golden fixtures (scipy expm_multiply / numpy eigvalsh applied to the
reference's own matrices), scipy on the fly for bigger cases, and oracle.rdm.
Tolerances: evolve states and eigenvalues 1e-10 (north star)."""
import numpy as np
import pytest
import scipy.sparse.linalg

import oracle
from helpers import case_terms, golden_cases, kats, product_subspace, rand_state, rel_err

pytestmark = pytest.mark.gpu
CASES = golden_cases()


def operator_for(tag):
    """dynamite_b200 Operator + subspace + input State for a golden case"""
    from dynamite_b200.operators import Operator
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Full, XParity
    c = CASES[tag]
    L = c['left']['L']
    if c['xparity']:
        # golden MSC is already reduced; rebuild the unreduced operator from its name
        from dynamite_b200.hamiltonians import build_hamiltonian
        H = build_hamiltonian(tag.split('_')[0], L)
        sub = XParity(Full(L=L), sector='+' if tag.endswith('plus') else '-')
    else:
        H = Operator(msc=case_terms(c), L=L)
        sub = product_subspace(c['left'])
    H.allow_projection = True
    H.subspace = sub
    x = State(L=L, subspace=sub)
    x.vec[0:len(c['x'])] = c['x']
    x.set_initialized()
    return c, H, sub, x


EVOLVE_TAGS = sorted(t for t in CASES if 'evolved' in CASES[t])


@pytest.mark.parametrize('tag', EVOLVE_TAGS)
def test_evolve_vs_golden(gpu, tag):
    c, H, sub, x = operator_for(tag)
    t = float(c['evolve_t'])
    y = H.evolve(x, t, tol=1e-13)
    assert rel_err(y.to_numpy(), c['evolved']) < 1e-10
    assert abs(y.norm() - 1.0) < 1e-10
    # imaginary time (real scale): exp(-t H) x
    yi = H.evolve(x, -1j * t, tol=1e-13)
    assert rel_err(yi.to_numpy(), c['evolved_imag']) < 1e-10
    # default tolerance (1e-7) is what the reference's tests use with a 1e-9 overlap check
    yd = H.evolve(x, t)
    ov = np.vdot(c['evolved'], yd.to_numpy())
    assert abs(1 - ov) < 1e-7
    # backwards in time returns to the start
    back = H.evolve(y, -t, tol=1e-13)
    assert rel_err(back.to_numpy(), c['x']) < 1e-10


@pytest.mark.parametrize('name,L,ncv', [('MBL', 14, None), ('long_range', 13, 12), ('heisenberg', 15, 40), ('ising', 12, 5)])
def test_evolve_vs_scipy(gpu, name, L, ncv):
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State
    H = build_hamiltonian(name, L)
    x = State(L=L, state='random', seed=3)
    A = H.to_numpy()
    nrm = H.infinity_norm()
    assert abs(nrm - abs(A).sum(axis=1).max()) < 1e-11 * nrm
    for t in (1.0 / nrm, 7.0, 50.0 / nrm):
        want = scipy.sparse.linalg.expm_multiply(-1j * t * A, x.to_numpy())
        got = H.evolve(x, t, tol=1e-12, ncv=ncv)
        assert rel_err(got.to_numpy(), want) < 1e-10, (name, t)


def test_evolve_edge_cases(gpu):
    from dynamite_b200.computations import MaxIterationsError
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.operators import sigmax, sigmaz, index_sum
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Parity
    H = build_hamiltonian('long_range', 10)
    x = State(L=10, state='random', seed=1)
    # t = 0 is a copy
    assert np.array_equal(H.evolve(x, 0.0).to_numpy(), x.to_numpy())
    # too few iterations for a long evolution (reference test_evolve.py:195-202)
    with pytest.raises(MaxIterationsError):
        H.evolve(x, 500.0, max_its=2)
    # result vector reuse and subspace mismatch
    y = State(L=10)
    H.evolve(x, 0.3, result=y)
    with pytest.raises(ValueError):
        H.evolve(x, 0.3, result=State(L=10, subspace=Parity('even', L=10)))
    # pi pulse: exp(-i (pi/2) sum sigma_x) flips every spin (reference test_evolve.py:23-32)
    L = 8
    Hx = index_sum(sigmax(), size=L)
    Hx.L = L
    s = State(L=L, state='U' * L)
    out = Hx.evolve(s, np.pi / 2, tol=1e-12).to_numpy()
    want = np.zeros(1 << L, complex)
    want[-1] = (-1j) ** L
    assert np.allclose(out, want, atol=1e-10)
    # tiny dimension: Krylov space is exhausted (happy breakdown)
    Hz = sigmaz(0) + 0.5 * sigmax(0) * sigmax(1)
    Hz.L = 2
    s2 = State(L=2, state='random', seed=5)
    A = Hz.to_numpy(sparse=False)
    want = scipy.linalg.expm(-1j * 2.5 * A) @ s2.to_numpy()
    assert np.allclose(Hz.evolve(s2, 2.5).to_numpy(), want, atol=1e-10)
    # an eigenstate only picks up a phase (invariant subspace of dimension 1)
    Hd = index_sum(sigmaz(0) * sigmaz(1), size=6)
    Hd.L = 6
    e = State(L=6, state='UDUDUD')
    assert np.allclose(Hd.evolve(e, 1.7).to_numpy(), np.exp(-1j * 1.7 * -5) * e.to_numpy(), atol=1e-12)


EIG_TAGS = sorted(t for t in CASES if 'evals' in CASES[t])


@pytest.mark.parametrize('tag', EIG_TAGS)
def test_eigsolve_vs_golden(gpu, tag):
    c, H, sub, x = operator_for(tag)
    w = c['evals']
    n = w.size
    nev = min(4, n)
    for which, want in (('lowest', w[:nev]), ('highest', w[::-1][:nev]),
                        ('exterior', w[np.argsort(-np.abs(w), kind='stable')][:nev])):
        evals, evecs = H.eigsolve(nev=nev, which=which, getvecs=True, tol=1e-12)
        assert len(evals) >= nev
        scale = max(1.0, np.max(np.abs(w)))
        # Krylov methods may miss copies of a degenerate eigenvalue (documented in the
        # reference, computations.py:139-142): compare the DISTINCT values, in order,
        # and require every returned value to be a true eigenvalue.
        def distinct(vals):
            out = []
            for v in vals:
                if not out or abs(v - out[-1]) > 1e-8 * scale:
                    out.append(v)
            return np.array(out)
        key = (lambda v: -np.abs(v)) if which == 'exterior' else (lambda v: v if which == 'lowest' else -v)
        got_d = distinct(sorted(evals[:nev], key=key))
        want_d = distinct(sorted(w, key=key))[:got_d.size]
        if which == 'exterior':
            assert np.allclose(np.abs(got_d), np.abs(want_d), atol=1e-10 * scale), (tag, which, evals, want)
        else:
            assert np.allclose(got_d, want_d, atol=1e-10 * scale), (tag, which, evals, want)
        for lam in evals[:nev]:
            assert np.min(np.abs(w - lam)) < 1e-10 * scale
        A = c['A']
        V = np.array([v.to_numpy() for v in evecs[:nev]])
        for lam, v in zip(evals[:nev], V):
            assert abs(np.linalg.norm(v) - 1) < 1e-10
            assert np.linalg.norm(A @ v - lam * v) < 1e-8 * scale
        # orthonormal even inside degenerate multiplets
        assert np.allclose(V.conj() @ V.T, np.eye(nev), atol=1e-8)


def test_eigsolve_larger_and_errors(gpu):
    from dynamite_b200.computations import MaxIterationsError
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.operators import index_sum, sigmax
    from dynamite_b200.subspaces import SpinConserve, XParity
    L = 14
    H = build_hamiltonian('heisenberg', L)
    H.subspace = SpinConserve(L, L // 2)
    A = H.to_numpy()
    want = scipy.sparse.linalg.eigsh(A, k=5, which='SA')[0]
    want.sort()
    evals, evecs = H.eigsolve(nev=5, getvecs=True, tol=1e-12)
    assert np.allclose(evals[:5], want, atol=1e-10)
    for lam, v in zip(evals[:5], evecs[:5]):
        assert np.linalg.norm(A @ v.to_numpy() - lam * v.to_numpy()) < 1e-8
    # default tolerance
    e1 = H.eigsolve()
    assert abs(e1[0] - want[0]) < 1e-7
    # XParity on top of SpinConserve: the ground state lives in one of the two sectors
    lows = []
    for sector in '+-':
        Hx = build_hamiltonian('heisenberg', L)
        Hx.subspace = XParity(SpinConserve(L, L // 2), sector=sector)
        lows.append(Hx.eigsolve(nev=1, tol=1e-12)[0])
    assert abs(min(lows) - want[0]) < 1e-10
    # analytic spectrum of sum sigma_x: -L, -L+2, ... (reference test_eigsolve.py:95-123)
    Hs = index_sum(sigmax(), size=8)
    Hs.L = 8
    ev = Hs.eigsolve(nev=1, tol=1e-12)
    assert abs(ev[0] + 8) < 1e-10
    assert abs(Hs.eigsolve(nev=1, which='highest', tol=1e-12)[0] - 8) < 1e-10
    with pytest.raises(RuntimeError):
        H.eigsolve(target=0.1)
    with pytest.raises(MaxIterationsError):
        build_hamiltonian('MBL', 12).eigsolve(nev=6, tol=1e-14, max_its=2)


def test_rdm_kats_and_oracle(gpu):
    from dynamite_b200.computations import entanglement_entropy, reduced_density_matrix
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Full, XParity
    k = kats()
    psi = np.array([complex(*v) for v in k['rdm_L4_state']])
    s = State(L=4)
    s.vec[0:16] = psi
    s.set_initialized()
    for keep, key in (([0], 'rdm_L4_keep0'), ([2], 'rdm_L4_keep2')):
        want = np.array([[complex(*v) for v in row] for row in k[key]])
        got = reduced_density_matrix(s, keep)
        assert np.allclose(got, want, atol=2e-6)
        assert abs(entanglement_entropy(s, keep) - k[key + '_entropy']) < 2e-5
    for keep, key in (([0, 2], 'rdm_L4_keep02_entropy'), ([1, 3], 'rdm_L4_keep13_entropy')):
        assert abs(entanglement_entropy(s, keep) - k[key]) < 2e-5
    full4 = oracle.Subspace({'type': 'full', 'L': 4})
    for keep in ([0], [3], [0, 1], [1, 2, 3], [0, 1, 2, 3]):
        assert np.allclose(reduced_density_matrix(s, keep), oracle.rdm(psi, full4, keep), atol=1e-14)
    # an empty keep never reaches the backend in the reference (computations.py:331-332)
    assert np.array_equal(reduced_density_matrix(s, []), np.array([[1]], dtype=np.complex128))
    cs = k['rdm_complex_sign']
    s2 = State(L=2)
    s2.vec[0:4] = np.array([complex(*v) for v in cs['state']])
    s2.set_initialized()
    assert np.array_equal(reduced_density_matrix(s2, cs['keep']),
                          np.array([[complex(*v) for v in row] for row in cs['dm']]))
    with pytest.raises(ValueError):
        reduced_density_matrix(s, [2, 1])
    with pytest.raises(ValueError):
        reduced_density_matrix(State(subspace=XParity(Full(L=4)), state='uniform'), [0])


@pytest.mark.parametrize('spec', [
    {'type': 'full', 'L': 13}, {'type': 'parity', 'L': 12, 'space': 1}, {'type': 'spinconserve', 'L': 14, 'k': 6},
    {'type': 'explicit', 'L': 10, 'states': list(range(5, 900, 3))},
])
def test_rdm_subspaces_vs_oracle(gpu, spec):
    from dynamite_b200.computations import reduced_density_matrix
    from dynamite_b200.states import State
    sub = product_subspace(spec)
    osub = oracle.Subspace(spec)
    psi = rand_state(osub.dim, 13)
    s = State(subspace=sub)
    s.vec[0:osub.dim] = psi
    s.set_initialized()
    L = spec['L']
    for keep in ([0], [L - 1], [0, 1, 2], [1, 4, L - 2], list(range(0, L // 2)), list(range(L - 6, L))):
        got = reduced_density_matrix(s, keep)
        want = oracle.rdm(psi, osub, keep)
        assert np.allclose(got, want, atol=1e-13), keep
        assert abs(np.trace(got) - 1) < 1e-12


def test_state_api_and_checkpoint(gpu, tmp_path):
    """State surface used around the path: products, projection, XParity conversion, save/load."""
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State, UninitializedError
    from dynamite_b200.subspaces import Full, SpinConserve, XParity
    s = State(L=6, state='UDUDUD')
    assert s.to_numpy()[0b101010] == 1 and abs(s.norm() - 1) < 1e-15
    with pytest.raises(UninitializedError):
        State(L=6).assert_initialized()
    with pytest.raises(ValueError):
        State(subspace=SpinConserve(6, 3), state='UUUUUU')
    r = State(L=8, state='random', seed=4)
    R = np.random.RandomState(4)
    want = R.standard_normal(256) + 1j * R.standard_normal(256)     # reference stream, states.py:272-318
    assert np.allclose(r.to_numpy(), want / np.linalg.norm(want), atol=1e-15)
    # save / from_file round trip
    r.save(str(tmp_path / 'ckpt'))
    back = State.from_file(str(tmp_path / 'ckpt'))
    assert np.array_equal(back.to_numpy(), r.to_numpy()) and back.subspace == r.subspace
    # byte-level: PETSc's binary Vec = VEC_FILE_CLASSID (1211214) and the length as big-endian PetscInt,
    # then big-endian complex128.  A default PETSc build (32-bit PetscInt) is what save() writes; a
    # --with-64-bit-indices file must load too, as must metadata that names the reference's module.
    import pickle
    import struct
    raw = open(str(tmp_path / 'ckpt.vec'), 'rb').read()
    assert raw[:8] == struct.pack('>ii', 1211214, 256) and len(raw) == 8 + 16 * 256
    assert np.array_equal(np.frombuffer(raw[8:], dtype='>c16'), r.to_numpy())
    r.save(str(tmp_path / 'ckpt64'), indices=64)
    raw64 = open(str(tmp_path / 'ckpt64.vec'), 'rb').read()
    assert raw64[:16] == struct.pack('>qq', 1211214, 256) and raw64[16:] == raw[8:]
    assert np.array_equal(State.from_file(str(tmp_path / 'ckpt64')).to_numpy(), r.to_numpy())
    meta = pickle.dumps(r.subspace, protocol=0).replace(b'dynamite_b200.subspaces', b'dynamite.subspaces')
    assert b'dynamite.subspaces' in meta
    open(str(tmp_path / 'ckpt64.metadata'), 'wb').write(meta)
    assert State.from_file(str(tmp_path / 'ckpt64')).subspace == r.subspace
    open(str(tmp_path / 'bad.vec'), 'wb').write(struct.pack('>ii', 7, 256) + raw[8:])
    open(str(tmp_path / 'bad.metadata'), 'wb').write(pickle.dumps(r.subspace))
    with pytest.raises(RuntimeError):
        State.from_file(str(tmp_path / 'bad'))
    # projection
    p = r.copy()
    p.project(3, 1)
    v = p.to_numpy()
    idx = np.arange(256)
    assert np.all(v[((idx >> 3) & 1) == 0] == 0) and abs(p.norm() - 1) < 1e-14
    # XParity <-> parent conversion is an isometry onto the sector, and H commutes with it
    parent = SpinConserve(8, 4)
    xp = XParity(parent, sector='-')
    sx = State(subspace=xp, state='random', seed=9)
    up = xp.convert_state(sx)
    assert up.subspace is parent and abs(up.norm() - 1) < 1e-13
    down = xp.convert_state(up)
    assert np.allclose(down.to_numpy(), sx.to_numpy(), atol=1e-14)
    H = build_hamiltonian('heisenberg', 8)
    H.add_subspace(xp)
    H.add_subspace(parent)
    lhs = xp.convert_state(H.dot(sx))          # convert(H_x psi)
    rhs = H.dot(up)                            # H_parent convert(psi)
    assert np.allclose(lhs.to_numpy(), rhs.to_numpy(), atol=1e-13)


def test_expectation_pipeline(gpu):
    """Operator.expectation and an evolve -> expectation chain (reference operators.py:775-796,
    examples/scripts/SYK/run_syk.py:179-206) against dense numpy."""
    import scipy.linalg
    from dynamite_b200 import msc_tools
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Full
    L = 8
    H = build_hamiltonian('MBL', L)
    W = build_hamiltonian('long_range', L)
    for op in (H, W):
        op.subspace = Full(L=L)
    psi = State(L=L, state='random', seed=3)
    v = psi.to_numpy()
    full = oracle.Subspace({'type': 'full', 'L': L})
    dense = {}
    for name, op in (('H', H), ('W', W)):
        op.reduce_msc()
        A = msc_tools.msc_to_numpy(op.msc, (256, 256), full.i2s, full.s2i)
        dense[name] = np.asarray(A.todense() if hasattr(A, 'todense') else A, dtype=np.complex128)
    assert abs(H.expectation(psi) - np.vdot(v, dense['H'] @ v).real) < 1e-12
    tmp = State(L=L)
    assert abs(W.expectation(psi, tmp_state=tmp) - np.vdot(v, dense['W'] @ v).real) < 1e-12
    t = 0.7
    got = W.expectation(H.evolve(psi, t, tol=1e-12))
    vt = scipy.linalg.expm(-1j * t * dense['H']) @ v
    assert abs(got - np.vdot(vt, dense['W'] @ vt).real) < 1e-10
    bra, ket = H.create_states()
    assert bra.subspace == H.left_subspace and ket.subspace == H.right_subspace


@pytest.mark.parametrize('name,L,start', [('heisenberg', 16, 0b0101010101010101), ('XX', 14, 0b111), ('SYK', 7, 0),
                                          ('long_range', 10, 5), ('heisenberg', 22, (1 << 11) - 1)])
def test_compute_rcm_device(gpu, name, L, start):
    """Auto-subspace search on the device (csrc/rcm.cu) against the host port of the reference's queue
    walk (bsubspace.pyx:212-261): the same states in the same order, and the same size error."""
    import ctypes as C
    from dynamite_b200 import _capi
    from dynamite_b200.hamiltonians import build_hamiltonian
    H = build_hamiltonian(name, L)
    H.reduce_msc()
    masks = np.ascontiguousarray(H.msc['masks'], dtype=np.int64)
    signs = np.ascontiguousarray(H.msc['signs'], dtype=np.int64)
    coeffs = np.ascontiguousarray(H.msc['coeffs'], dtype=np.complex128)
    lib = _capi.lib()

    def run(fn, cap):
        out = np.full(cap, -7, dtype=np.int64)
        dim = C.c_int64()
        rc = fn(masks.size, _capi.ip(masks), _capi.ip(signs), _capi.fp(coeffs), _capi.ip(out), cap, int(start), L,
                C.byref(dim))
        return rc, dim.value, out

    cap = 1 << min(L, 21)
    rc_h, dim_h, host = run(lib.dnm_compute_rcm, cap)
    rc_d, dim_d, dev = run(lib.dnm_compute_rcm_device, cap)
    assert rc_h == 0 and rc_d == 0 and dim_h == dim_d and dim_h > 1
    assert np.array_equal(host[:dim_h], dev[:dim_d])
    if dim_h > 4:
        rc_h, _, _ = run(lib.dnm_compute_rcm, dim_h - 3)
        rc_d, _, _ = run(lib.dnm_compute_rcm_device, dim_h - 3)
        assert rc_h != 0 and rc_d != 0          # 'state_map size too small' from both


def test_vec_shift_normalize_and_evolve_algo(gpu):
    """C-ABI entries that complete the reference's Vec / MFN surface: Vec.shift (states.py:805),
    Vec.normalize on the device, and the MFN type passed to dnm_evolve_algo."""
    import scipy.linalg
    from dynamite_b200 import msc_tools
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Full
    L = 8
    s = State(L=L, state='random', seed=5)
    v = s.to_numpy().copy()
    s += 0.25 - 0.5j                       # a number: added to every amplitude
    assert np.allclose(s.to_numpy(), v + (0.25 - 0.5j), atol=1e-15)
    nrm = s.vec.normalize()
    assert abs(nrm - np.linalg.norm(v + (0.25 - 0.5j))) < 1e-13 and abs(s.norm() - 1) < 1e-14
    H = build_hamiltonian('MBL', L)
    H.subspace = Full(L=L)
    H.reduce_msc()
    full = oracle.Subspace({'type': 'full', 'L': L})
    A = msc_tools.msc_to_numpy(H.msc, (256, 256), full.i2s, full.s2i)
    A = np.asarray(A.todense() if hasattr(A, 'todense') else A, dtype=np.complex128)
    psi = State(L=L, state='random', seed=6)
    want = scipy.linalg.expm(-1.3j * A) @ psi.to_numpy()
    for algo in ('expokit', 'krylov'):
        got = H.evolve(psi, 1.3, tol=1e-12, algo=algo).to_numpy()
        assert np.linalg.norm(got - want) < 1e-10, algo
