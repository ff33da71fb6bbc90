"""GPU parity at the BASELINE.json sizes, through size-independent properties and sampled rows
(the oracle cannot finish a whole L=30 product in test time, and the reference's own L=30 test set
takes hours, tests/integration/test_sets/L30.tests:1-2):

* two independent GPU algorithms (window-tiled kernel vs plain gather kernel) agree at L=28;
* sampled blocks of rows of the L=30 product equal the oracle's fast path on the same rows;
* Hermiticity <x|Hy> = conj(<y|Hx>), linearity, evolve norm preservation and time reversal;
* the L=26 SpinConserve eigensolve (BASELINE C2) returns eigenpairs with small residuals.
"""
import ctypes as C
import mmap

import numpy as np
import pytest

import oracle
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _reserve(count, dtype):
    nbytes = int(count) * np.dtype(dtype).itemsize
    flags = mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS | getattr(mmap, 'MAP_NORESERVE', 0x4000)
    return np.frombuffer(mmap.mmap(-1, nbytes, flags=flags), dtype=dtype, count=int(count))


def _free_gib(gpu):
    f, t = C.c_int64(), C.c_int64()
    gpu.check(gpu.lib().dnm_mem_info(C.byref(f), C.byref(t)))
    return f.value / 2**30


def _setup(name, L, sub=None):
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Full
    H = build_hamiltonian(name, L)
    H.subspace = sub if sub is not None else Full(L=L)
    x = State(subspace=H.subspace)
    x.vec.setRandom(11)
    x.vec.normalize()
    x.set_initialized()
    return H, x


def test_tiled_vs_gather_kernel_L28(gpu):
    if _free_gib(gpu) < 20:
        pytest.skip('needs 20 GiB of device memory')
    from dynamite_b200.states import State
    H, x = _setup('MBL', 28)
    mat = H.get_mat()
    y1, y2 = State(subspace=H.subspace), State(subspace=H.subspace)
    mat.set_option('kernel', 2)
    H.dot(x, y1)
    assert mat.get_info('kernel') == 2
    mat.set_option('kernel', 1)
    H.dot(x, y2)
    assert mat.get_info('kernel') == 1
    nrm = y1.norm()
    y2.axpy(-1, y1)
    assert y2.norm() / nrm < 1e-12
    H.destroy_mat()


@pytest.mark.parametrize('L', [30])
def test_sampled_rows_vs_oracle_fullsize(gpu, L):
    """rows [0, S) and a block in the middle of the L=30 MBL product against the oracle."""
    if _free_gib(gpu) < 48:
        pytest.skip('needs 48 GiB of device memory')
    from dynamite_b200 import msc_tools
    from dynamite_b200.states import State
    H, x = _setup('MBL', L)
    y = State(subspace=H.subspace)
    H.dot(x, y)
    n = 1 << L
    H.reduce_msc()
    masks, offs = msc_tools.mask_offsets(H.msc)
    omsc = oracle.Msc(masks, offs, H.msc['signs'], H.msc['coeffs'])
    osub = oracle.Subspace({'type': 'full', 'L': L})
    S = 1 << 16
    xh, yh = _reserve(n, np.complex128), _reserve(n, np.complex128)
    for first in (0, (n // 2 + 12345 * 2048) & ~(S - 1), n - S):
        # the rows [first, first+S) only read x inside one S-long window per mask
        for m in masks:
            w = (first ^ int(m)) & ~(S - 1)
            xh[w:w + S] = x.vec[w:w + S]
        oracle.matmult_fast_range(omsc, osub, xh, yh, first // 2048, (first + S) // 2048, nthreads=8)
        assert rel_err(y.vec[first:first + S], yh[first:first + S]) < 1e-12
    H.destroy_mat()


def test_hermiticity_and_linearity_fullsize(gpu):
    if _free_gib(gpu) < 100:
        pytest.skip('needs 100 GiB of device memory')
    from dynamite_b200.states import State
    L = 30
    H, x = _setup('MBL', L)
    z = State(subspace=H.subspace)
    z.vec.setRandom(5)
    z.vec.normalize()
    z.set_initialized()
    Hx, Hz = H.dot(x), H.dot(z)
    a, b = z.dot(Hx), Hz.dot(x)          # <z|Hx> and <Hz|x>
    assert abs(a - b) < 1e-12 * H.infinity_norm()
    # linearity: H(x + 2i z) = Hx + 2i Hz
    x.axpy(2j, z)
    Hc = H.dot(x)
    Hx.axpy(2j, Hz)
    nrm = Hc.norm()
    Hc.axpy(-1, Hx)
    assert Hc.norm() / nrm < 1e-13
    H.destroy_mat()


def test_evolve_unitarity_and_reversal_L28(gpu):
    if _free_gib(gpu) < 80:
        pytest.skip('needs 80 GiB of device memory')
    H, x = _setup('MBL', 28)
    t = 2.0 / H.infinity_norm()
    y = H.evolve(x, t, tol=1e-11, ncv=12)
    assert abs(y.norm() - 1) < 1e-10
    back = H.evolve(y, -t, tol=1e-11, ncv=12)
    back.axpy(-1, x)
    assert back.norm() < 1e-9
    H.destroy_mat()


def test_eigsolve_C2_fullsize(gpu):
    """BASELINE C2: L=26 Heisenberg, SpinConserve(k=13), lowest 4 eigenpairs."""
    from dynamite_b200.subspaces import SpinConserve
    H, _ = _setup('heisenberg', 26, SpinConserve(26, 13))
    evals, evecs = H.eigsolve(nev=4, getvecs=True)
    assert np.all(np.diff(evals[:4]) >= -1e-9)
    # Bethe-ansatz-quality regression value of the open-chain ground state energy (this code, tol 1e-12)
    assert abs(evals[0] + 11.339579652755) < 1e-6
    nrmH = H.infinity_norm()
    for lam, v in zip(evals[:4], evecs[:4]):
        r = H.dot(v)
        r.axpy(-lam, v)
        assert r.norm() < 1e-6 * nrmH          # default tolerance 1e-8 relative
        assert abs(v.norm() - 1) < 1e-10
    assert abs(evecs[0].dot(evecs[1])) < 1e-8
    H.destroy_mat()
