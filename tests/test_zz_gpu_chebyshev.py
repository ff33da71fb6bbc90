"""GPU: the Chebyshev propagator (csrc/chebyshev.h instantiated on device vectors in krylov.cu).

The recurrence and its coefficients are checked on the CPU (tests/test_chebyshev_host.py, the same
template on host vectors); here the device instantiation is compared with scipy, with the golden
fixtures and with the expokit path.  The library picks it by itself only when a Krylov basis does not
fit device memory (L >= 30 on one B200), so the tests ask for it: algo='chebyshev' or DNM_EVOLVE_CHEB=1."""
import numpy as np
import pytest
import scipy.sparse.linalg

from helpers import rel_err
from test_gpu_krylov import EVOLVE_TAGS, operator_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,L', [('MBL', 14), ('long_range', 13), ('heisenberg', 15), ('ising', 12), ('XX', 12)])
def test_chebyshev_vs_scipy(gpu, name, L):
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State
    H = build_hamiltonian(name, L)
    x = State(L=L, state='random', seed=3)
    A = H.to_numpy()
    nrm = H.infinity_norm()
    for t in (1.0 / nrm, 7.0, -50.0 / nrm):
        want = scipy.sparse.linalg.expm_multiply(-1j * t * A, x.to_numpy())
        got = H.evolve(x, t, algo='chebyshev')
        assert rel_err(got.to_numpy(), want) < 1e-10, (name, t)
        assert abs(got.norm() - 1.0) < 1e-12


@pytest.mark.parametrize('tag', EVOLVE_TAGS)
def test_chebyshev_vs_golden(gpu, tag):
    """every subspace type of the fixtures (Full, Parity, SpinConserve, Explicit, XParity)"""
    c, H, sub, x = operator_for(tag)
    t = float(c['evolve_t'])
    y = H.evolve(x, t, algo='chebyshev')
    assert rel_err(y.to_numpy(), c['evolved']) < 1e-10
    back = H.evolve(y, -t, algo='chebyshev')
    assert rel_err(back.to_numpy(), c['x']) < 1e-10


def test_chebyshev_selection_and_limits(gpu, monkeypatch):
    from dynamite_b200 import petsc
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.states import State
    L = 20
    H = build_hamiltonian('heisenberg', L)
    x = State(L=L, state='random', seed=2)
    t = 30.0 / H.infinity_norm()
    from dynamite_b200 import computations
    ref = H.evolve(x, t, tol=1e-13, algo='expokit').to_numpy()
    assert computations.last_evolve['algo'] == 'expokit'
    assert rel_err(H.evolve(x, t, algo='chebyshev').to_numpy(), ref) < 1e-10
    assert computations.last_evolve['algo'] == 'chebyshev' and computations.last_evolve['iterations'] == 1
    z = 30.0 * (1 + 1e-9)
    assert z < computations.last_evolve['matmults'] < z + 12 * (z + 1) ** (1 / 3) + 40
    # the default takes the expokit path while the basis fits ...
    assert rel_err(H.evolve(x, t, tol=1e-13).to_numpy(), ref) < 1e-10
    assert computations.last_evolve['algo'] == 'expokit' and computations.last_evolve['requested'] == 'auto'
    # ... and the propagator when told to through the environment
    monkeypatch.setenv('DNM_EVOLVE_CHEB', '1')
    assert rel_err(H.evolve(x, t).to_numpy(), ref) < 1e-10
    assert computations.last_evolve['algo'] == 'chebyshev'
    monkeypatch.delenv('DNM_EVOLVE_CHEB')
    # imaginary time is not unitary: the propagator refuses, the default still serves it
    with pytest.raises(petsc.Error):
        H.evolve(x, -0.1j, algo='chebyshev')
    H.evolve(x, -0.1j)
    # a zero evolution time is a copy
    assert np.array_equal(H.evolve(x, 0.0, algo='chebyshev').to_numpy(), x.to_numpy())
