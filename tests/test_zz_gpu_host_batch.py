"""GPU: the host-buffer entry points (dnm_mat_mult_host, dnm_mat_mult_host_batch) against the oracle.

The batch call pipelines the copies of consecutive products over the PCIe link (two device buffers per
direction, three streams); what it returns must not depend on that: every y_k equals the oracle's
product of x_k, for one, two and more products than buffers, with pageable and with pinned host
memory, and with the aliasing bench.py uses (one input and one output buffer for all products)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from dynamite_b200 import msc_tools

pytestmark = pytest.mark.gpu


def _problem(name, L):
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.subspaces import Full
    H = build_hamiltonian(name, L)
    H.subspace = Full(L=L)
    H.reduce_msc()
    masks, offs = msc_tools.mask_offsets(H.msc)
    omsc = oracle.Msc(masks, offs, H.msc['signs'], H.msc['coeffs'])
    osub = oracle.Subspace({'type': 'full', 'L': L})
    return H, omsc, osub


def _inputs(n, count, seed=3):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(count)]


def _pinned(gpu, n):
    ptr = C.c_void_p()
    gpu.check(gpu.lib().dnm_host_alloc(n * 16, C.byref(ptr)))
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
    return ptr, arr


@pytest.mark.parametrize('name,L', [('heisenberg', 14), ('MBL', 22)])
@pytest.mark.parametrize('count', [1, 2, 5])
def test_host_batch_vs_oracle(gpu, name, L, count):
    H, omsc, osub = _problem(name, L)
    mat = H.get_mat()
    n = 1 << L
    xs = _inputs(n, count)
    ys = [np.full(n, np.nan + 0j) for _ in range(count)]
    mat.mult_host_batch(xs, ys)
    single = np.empty(n, dtype=np.complex128)
    for x, y in zip(xs, ys):
        want = oracle.matmult(omsc, osub, osub, x)
        assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
        mat.mult_host(x, single)
        assert np.abs(single - y).max() <= 1e-14 * np.abs(want).max()     # the same kernels on the same input
    mat.mult_host_batch([], [])                       # nothing to do is not an error
    H.destroy_mat()


def test_host_batch_pinned_and_aliased_buffers(gpu):
    """bench.py's use: one pinned input and one pinned output for every product of the call."""
    L = 22
    H, omsc, osub = _problem('MBL', L)
    mat = H.get_mat()
    n = 1 << L
    px, x = _pinned(gpu, n)
    py, y = _pinned(gpu, n)
    try:
        x[:] = _inputs(n, 1, seed=9)[0]
        y[:] = np.nan
        mat.mult_host_batch([x] * 6, [y] * 6)
        want = oracle.matmult(omsc, osub, osub, x)
        assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
        # a second call reuses the device buffers and the streams
        x2 = 2.0 * x
        y[:] = np.nan
        mat.mult_host_batch([x, x2, x], [y, y, y])
        assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
        with pytest.raises(ValueError):
            mat.mult_host_batch([x], [x])
        with pytest.raises(ValueError):
            mat.mult_host_batch([x, x], [y])
    finally:
        H.destroy_mat()
        gpu.lib().dnm_host_free(px)
        gpu.lib().dnm_host_free(py)
