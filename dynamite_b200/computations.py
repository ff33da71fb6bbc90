"""
``evolve``, ``eigsolve``, ``reduced_density_matrix`` and the entropies (host
mirror of reference ``computations.py``).  The Krylov loops run device-resident
inside the backend; these wrappers do argument checking and map solver status
to the reference's exceptions (``computations.py:114-122, 261-275``).
"""
import warnings

import numpy as np

from . import config
from .msc_tools import dnm_int_t


class ConvergenceError(Exception):
    pass


class MaxIterationsError(ConvergenceError):
    pass


# what the most recent evolve() did (diagnostics: bench.py reports the MatMult count next to the seconds)
last_evolve = {}


def evolve(H, state, t, result=None, tol=None, ncv=None, algo=None, max_its=None):
    r"""``result = exp(-i H t) state`` (reference ``computations.py:10-126``).
    Imaginary ``t`` gives imaginary-time evolution."""
    state.assert_initialized()
    config._initialize()
    from . import slepc as SLEPc
    from .states import State

    H.establish_L()
    if not H.has_subspace(state.subspace, state.subspace):
        raise ValueError('Hamiltonian and state are defined on different subspaces.')
    if result is None:
        result = State(L=H.L, subspace=state.subspace)
    elif state.subspace != result.subspace:
        raise ValueError('input and result states are on different subspaces.')
    if t == 0.0:
        state.copy(result)
        return result

    mfn = SLEPc.MFN().create()
    f = mfn.getFN()
    f.setType(SLEPc.FN.Type.EXP)
    f.setScale(-1j * t)
    # (the reference defaults to 'expokit'; 'auto' is expokit unless its Krylov basis does not fit the GPU)
    mfn.setType(algo if algo is not None else 'auto')
    if ncv is not None:
        mfn.setDimensions(ncv)
    mfn.setTolerances(tol=tol, max_it=max_its)
    mfn.setFromOptions()
    mfn.setOperator(H.get_mat(subspaces=(state.subspace, state.subspace)))
    mfn.solve(state.vec, result.vec)
    last_evolve.update(algo=mfn.used, requested=mfn.type, iterations=mfn.getIterationNumber(), matmults=mfn.matmults)

    conv = mfn.getConvergedReason()
    if conv == SLEPc.MFN.ConvergedReason.DIVERGED_ITS:
        raise MaxIterationsError('solver reached maximum number of iterations without converging. '
                                 'perhaps try increasing the max iterations with the "max_its" argument.')
    if conv == SLEPc.MFN.ConvergedReason.DIVERGED_BREAKDOWN:
        raise ConvergenceError('solver failed to converge with MFN_DIVERGED_BREAKDOWN.')
    if conv <= 0:
        raise ConvergenceError('solver failed to converge.')
    result.set_initialized()
    return result


def eigsolve(H, getvecs=False, nev=1, which='lowest', target=None, tol=None, subspace=None, max_its=None):
    """A few extremal eigenpairs of the Hermitian operator ``H``
    (reference ``computations.py:128-292``)."""
    H.establish_L()
    if subspace is None:
        subspace = H.subspace
    elif not H.has_subspace(subspace):
        raise ValueError('Requested subspace has not been added to operator.')
    config._initialize()
    from . import slepc as SLEPc
    from .states import State

    eps = SLEPc.EPS().create()
    eps.setProblemType(SLEPc.EPS.ProblemType.HEP)
    if target is not None:
        # shell matrices and GPUs both rule out shift-invert in the reference (:214-220)
        raise RuntimeError('Shift-invert ("target") not supported for shell matrices.')
    if which == 'target':
        raise ValueError("Must specify target when setting which='target'")
    eps.setOperators(H.get_mat(subspaces=(subspace, subspace)))
    eps.setDimensions(nev)
    if which in ('smallest', 'largest'):
        warnings.warn('values "smallest" and "largest" for eigsolve parameter "which" are deprecated, '
                      'and have been replaced by "lowest" and "highest" respectively.',
                      DeprecationWarning, stacklevel=2)
        which = {'smallest': 'lowest', 'largest': 'highest'}[which]
    try:
        eps.setWhichEigenpairs({'lowest': SLEPc.EPS.Which.SMALLEST_REAL,
                                'highest': SLEPc.EPS.Which.LARGEST_REAL,
                                'exterior': SLEPc.EPS.Which.LARGEST_MAGNITUDE}[which])
    except KeyError:
        raise ValueError(f'invalid value "{which}" for parameter "which"') from None
    eps.setTolerances(tol=tol, max_it=max_its)
    eps.setFromOptions()
    eps.keep_vectors = bool(getvecs)
    eps.solve()
    nconv = eps.getConverged()
    reason = eps.getConvergedReason()
    try:
        if reason == SLEPc.EPS.ConvergedReason.DIVERGED_ITS:
            raise MaxIterationsError('eigensolver reached maximum number of iterations without converging. '
                                     'Try increasing the maximum iterations of the eigensolver via the '
                                     f'"max_its" argument to eigsolve() (current value: {eps.getTolerances()[1]})')
        if reason == SLEPc.EPS.ConvergedReason.DIVERGED_BREAKDOWN:
            raise ConvergenceError('eigsolver failed to converge with reason EPS_DIVERGED_BREAKDOWN')
        if reason == SLEPc.EPS.ConvergedReason.DIVERGED_SYMMETRY_LOST:
            raise ConvergenceError('eigsolver failed to converge with reason EPS_DIVERGED_SYMMETRY_LOST')
        if reason <= 0 or nconv < nev:
            raise ConvergenceError('eigsolver failed to converge')

        evals = np.array([eps.getEigenpair(i, None).real for i in range(nconv)], dtype=float)
        evecs = []
        if getvecs:
            for i in range(nconv):
                v = State(L=H.L, subspace=subspace)
                eps.getEigenpair(i, v.vec)
                v.set_initialized()
                evecs.append(v)
    finally:
        eps.destroy()
    return (evals, evecs) if getvecs else evals


def reduced_density_matrix(state, keep):
    """Trace out every spin not in ``keep`` (reference ``computations.py:294-350``).
    Computed on the device.  As in the reference the matrix is returned on rank 0 and every other
    rank gets the 1x1 matrix [[-1]]; an empty ``keep`` gives [[1]]."""
    if not state.subspace.product_state_basis:
        raise ValueError('reduced density matrices only supported for product state subspaces')
    keep = np.array(keep, dtype=dnm_int_t).reshape(-1)
    state.assert_initialized()
    if keep.size != np.unique(keep).size:
        raise ValueError('values in keep must be unique')
    if keep.size and (keep.min() < 0 or keep.max() >= state.L):
        raise ValueError('values in keep must be between 0 and L-1')
    if np.any(np.diff(keep) <= 0):
        raise ValueError('keep array must be strictly increasing')
    if keep.size == 0:
        return np.array([[1]], dtype=np.complex128)
    from ._backend import bpetsc
    from .petsc import COMM_WORLD
    dm = bpetsc.reduced_density_matrix(state.vec, state.subspace._to_c(), keep)
    if COMM_WORLD.size > 1 and COMM_WORLD.rank != 0:
        return np.array([[-1]], dtype=np.complex128)
    return dm


def entanglement_entropy(state, keep):
    return dm_entanglement_entropy(reduced_density_matrix(state, keep))


def dm_entanglement_entropy(dm):
    w = np.linalg.eigvalsh(dm)
    w = w[w > 0]
    return float(-np.sum(w * np.log(w)))


def renyi_entropy(state, keep, alpha, method='eigsolve'):
    return dm_renyi_entropy(reduced_density_matrix(state, keep), alpha, method)


def dm_renyi_entropy(dm, alpha, method='eigsolve'):
    if alpha == 0:
        return float(np.log(np.sum(np.linalg.eigvalsh(dm) > 1e-10)))
    if alpha == 1:
        return dm_entanglement_entropy(dm)
    if alpha == 'inf':
        return float(-np.log(np.max(np.linalg.eigvalsh(dm))))
    if method == 'matrix_power':
        if alpha != int(alpha):
            raise TypeError('alpha must be an integer for matrix_power method.')
        trace = np.trace(np.linalg.matrix_power(dm, int(alpha))).real
    elif method == 'eigsolve':
        trace = np.sum(np.linalg.eigvalsh(dm) ** alpha)
    else:
        raise ValueError('Valid methods are "eigsolve" and "matrix_power"')
    return float(np.log(trace) / (1 - alpha))


def get_tstep(ncv, nrm, tol=1e-7):
    """expokit's initial sub-step for a unit-norm start vector (reference ``computations.py:511-520``)"""
    f = ((ncv + 1) / 2.72) ** (ncv + 1) * np.sqrt(2 * np.pi * (ncv + 1))
    t = ((1 / nrm) * (f * tol) / (4.0 * nrm)) ** (1 / ncv)
    s = 10.0 ** (np.floor(np.log10(t)) - 1)
    return np.ceil(t / s) * s


def estimate_compute_time(t, ncv, nrm, tol=1e-7):
    return ncv * np.ceil(t / get_tstep(ncv, nrm, tol))
