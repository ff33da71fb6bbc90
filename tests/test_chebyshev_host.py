"""CPU: the Chebyshev propagator of csrc/chebyshev.h -- Bessel coefficients, truncation and the three-term
recurrence -- instantiated on host vectors (tests/cheb_host.cpp) and checked against scipy.  krylov.cu
instantiates the same template on device vectors; the GPU tests then only have to show that the device
operations it is given do what their names say (tests/test_zz_gpu_chebyshev.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg
import scipy.special

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def cheb(tmp_path_factory):
    so = tmp_path_factory.mktemp('cheb') / 'cheb_host.so'
    res = subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'dynamite_b200', 'csrc'),
                          '-o', str(so), os.path.join(ROOT, 'tests', 'cheb_host.cpp')], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lib.cheb_bessel.argtypes = [C.c_int, C.c_double, dp]
    lib.cheb_plan.restype = C.c_longlong
    lib.cheb_plan.argtypes = [C.c_double, C.c_double, C.c_double, C.c_longlong, dp, C.c_longlong, dp]
    lib.cheb_apply_dense.restype = C.c_longlong
    lib.cheb_apply_dense.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_longlong]
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize('x', [0.0, 1e-3, 0.7, 14.0, 50.0, 333.3, 5000.0])
def test_bessel_values(cheb, x):
    nmax = int(x + 12 * (x + 1) ** (1 / 3) + 60)
    out = np.empty(nmax + 1)
    cheb.cheb_bessel(nmax, x, _dp(out))
    want = scipy.special.jv(np.arange(nmax + 1), x)
    assert np.abs(out - want).max() < 2e-15 * max(1.0, x / 10)     # (scipy's own error grows with the argument)
    big = np.abs(want) > 1e-200
    assert np.abs(out[big] / want[big] - 1).max() < 1e-9 * max(1.0, x / 10)   # relative accuracy, tiny orders included


@pytest.mark.parametrize('z,expected', [(14.0, (36, 44)), (50.0, (80, 92)), (100.0, (138, 152))])
def test_truncation(cheb, z, expected):
    c = np.empty(2 * 4096)
    tail = C.c_double()
    n = cheb.cheb_plan(1.0, z, 1e-14, -1, _dp(c), 4096, C.byref(tail))
    assert expected[0] <= n - 1 <= expected[1]                      # MatMults for fourteen digits
    k = np.arange(n, n + 400)
    assert 2 * np.abs(scipy.special.jv(k, z)).sum() <= 1e-14 and tail.value <= 1e-14
    assert cheb.cheb_plan(1.0, z, 1e-14, 10, _dp(c), 4096, None) == 0   # max_terms too small


def _hermitian(n, seed, norm):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = A + A.conj().T
    return A * (norm / np.abs(A).sum(axis=1).max())


@pytest.mark.parametrize('n,s,norm', [(24, -1.0, 14.0), (40, -50.0 / 7.3, 7.3), (40, 50.0 / 7.3, 7.3), (16, -0.01, 3.0),
                                      (32, -400.0, 1.0), (8, 0.0, 2.0), (8, 5000.0, 1.0), (12, 1e-9, 3.0)])
def test_propagator_vs_expm(cheb, n, s, norm):
    A = np.ascontiguousarray(_hermitian(n, n, norm))
    rng = np.random.default_rng(1)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = np.empty(n, dtype=np.complex128)
    a = np.abs(A).sum(axis=1).max() * (1 + 1e-10)
    k = cheb.cheb_apply_dense(n, _dp(A.view(np.float64)), _dp(x.view(np.float64)), _dp(y.view(np.float64)), s, a, 1e-14, -1)
    assert k >= 0
    want = scipy.linalg.expm(1j * s * A) @ x
    assert np.abs(y - want).max() <= 2e-13 * np.linalg.norm(x) * max(1.0, abs(s) * norm / 50)
    assert abs(np.linalg.norm(y) / np.linalg.norm(x) - 1) < 1e-12          # unitary
    assert k <= abs(s) * a + 12 * (abs(s) * a + 1) ** (1 / 3) + 40


def test_loose_bound_costs_terms_not_accuracy(cheb):
    """a may overestimate the spectral radius (||A||_inf does): more terms, the same answer."""
    n = 20
    A = np.ascontiguousarray(_hermitian(n, 5, 4.0))
    x = np.ones(n, dtype=np.complex128)
    want = scipy.linalg.expm(-2j * A) @ x
    counts = []
    for a in (4.0 * (1 + 1e-10), 12.0):
        y = np.empty(n, dtype=np.complex128)
        counts.append(cheb.cheb_apply_dense(n, _dp(A.view(np.float64)), _dp(x.view(np.float64)), _dp(y.view(np.float64)),
                                            -2.0, a, 1e-14, -1))
        assert np.abs(y - want).max() < 1e-12
    assert counts[1] > counts[0]
