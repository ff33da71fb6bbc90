"""
The slice of the slepc4py surface that ``computations.evolve`` / ``eigsolve``
drive (reference ``computations.py:89-122, 208-287``): ``MFN`` (expokit
exponential action) and ``EPS`` (Hermitian Krylov-Schur), implemented by the
device-resident Krylov loops of the C ABI (``dnm_evolve`` / ``dnm_eigsolve``).
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, fp


class FN:
    class Type:
        EXP = 'exp'

    def __init__(self):
        self.type = None
        self.alpha = 1.0

    def setType(self, t):
        if t != FN.Type.EXP:
            raise ValueError('only the exponential is implemented')
        self.type = t

    def setScale(self, alpha, beta=1.0):
        self.alpha = complex(alpha) * complex(beta)


class MFN:
    class ConvergedReason:
        CONVERGED_TOL = 1
        CONVERGED_ITS = 2
        DIVERGED_ITS = -1
        DIVERGED_BREAKDOWN = -2
        ITERATING = 0

    # 'expokit' = sub-stepped scheme with the Lanczos recurrence, 'krylov' = the same with full Arnoldi
    # orthogonalisation, 'chebyshev' = Jacobi-Anger expansion (real time only), 'auto' = the library's
    # choice: expokit unless its basis would not fit device memory (include/dynamite_b200.h)
    _ALGO = {'auto': -1, 'expokit': 0, 'krylov': 1, 'chebyshev': 2}

    def create(self):
        self.fn = FN()
        self.type = 'expokit'
        self.ncv = 0
        self.tol = 0.0
        self.max_it = 0
        self.mat = None
        self.reason = 0
        self.its = 0
        self.matmults = 0
        return self

    def getFN(self):
        return self.fn

    def setType(self, t):
        if t not in self._ALGO:
            raise ValueError(f'unknown MFN type {t}')
        self.type = t

    def setDimensions(self, ncv):
        self.ncv = int(ncv)

    def setTolerances(self, tol=None, max_it=None):
        if tol is not None:
            self.tol = float(tol)
        if max_it is not None:
            self.max_it = int(max_it)

    def setFromOptions(self):
        pass

    def setOperator(self, mat):
        self.mat = mat

    def solve(self, b, x):
        reason, its, mm = C.c_int(), C.c_int(), C.c_int()
        a = self.fn.alpha
        check(_capi.lib().dnm_evolve_algo(self.mat.handle, b.handle, x.handle, a.real, a.imag,
                                          self.tol, self.ncv, self.max_it, self._ALGO[self.type],
                                          C.byref(reason), C.byref(its), C.byref(mm)))
        self.reason, self.its, self.matmults = reason.value, its.value, mm.value
        self.used = {0: 'expokit', 1: 'krylov', 2: 'chebyshev'}.get(_capi.lib().dnm_evolve_last_algo(), self.type)

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its


class ST:
    class Type:
        SINVERT = 'sinvert'

    def setType(self, t):
        raise RuntimeError('Shift-invert ("target") not supported for shell matrices.')


class EPS:
    class ProblemType:
        HEP = 1

    class Which:
        LARGEST_MAGNITUDE = 2
        SMALLEST_REAL = 0
        LARGEST_REAL = 1
        TARGET_MAGNITUDE = 7

    class ConvergedReason:
        CONVERGED_TOL = 1
        DIVERGED_ITS = -1
        DIVERGED_BREAKDOWN = -2
        DIVERGED_SYMMETRY_LOST = -3
        ITERATING = 0

    def create(self):
        self.mat = None
        self.nev = 1
        self.ncv = 0
        self.which = EPS.Which.SMALLEST_REAL
        self.tol = 0.0
        self.max_it = 0
        self.seed = 0
        self.reason = 0
        self.nconv = 0
        self.its = 0
        self.matmults = 0
        self._evals = None
        self._errest = None
        self._evecs = []
        return self

    def setProblemType(self, t):
        if t != EPS.ProblemType.HEP:
            raise ValueError('only Hermitian problems are implemented')

    def getST(self):
        return ST()

    def setOperators(self, mat):
        self.mat = mat

    def setDimensions(self, nev=None, ncv=None):
        if nev is not None:
            self.nev = int(nev)
        if ncv is not None:
            self.ncv = int(ncv)

    def setWhichEigenpairs(self, which):
        if which not in (0, 1, 2):
            raise ValueError('unsupported selection of eigenpairs')
        self.which = which

    def setTolerances(self, tol=None, max_it=None):
        if tol is not None:
            self.tol = float(tol)
        if max_it is not None:
            self.max_it = int(max_it)

    def getTolerances(self):
        return self.tol, self.max_it

    def setFromOptions(self):
        pass

    def setRandomSeed(self, seed):
        self.seed = int(seed)

    def solve(self):
        from .petsc import Vec
        m, _ = self.mat.getSize()
        for v in self._evecs:
            v.destroy()
        # room for a few more pairs than requested: more may converge together
        cap = self.nev + 8
        # eigenvector storage only when somebody will ask for vectors (EPS.keep_vectors, set by
        # computations.eigsolve from its getvecs argument): at L=30 nine unused 16 GiB vectors would
        # otherwise eat the memory the Lanczos basis needs
        if getattr(self, 'keep_vectors', True):
            self._evecs = [Vec(m) for _ in range(cap)]
            handles = (C.c_void_p * cap)(*[v.handle for v in self._evecs])
        else:
            self._evecs = []
            handles = None
        evals = np.zeros(cap, dtype=np.float64)
        errest = np.zeros(cap, dtype=np.float64)
        nconv, reason, its, mm = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(_capi.lib().dnm_eigsolve(self.mat.handle, self.nev, self.which, self.tol, self.max_it,
                                       self.ncv, self.seed, cap, C.byref(nconv), fp(evals), fp(errest),
                                       handles, C.byref(reason), C.byref(its), C.byref(mm)))
        self.nconv, self.reason, self.its, self.matmults = nconv.value, reason.value, its.value, mm.value
        self._evals, self._errest = evals, errest

    def getConverged(self):
        return min(self.nconv, len(self._evals))

    def getConvergedReason(self):
        return self.reason

    def getIterationNumber(self):
        return self.its

    def getEigenpair(self, i, vr=None):
        if not 0 <= i < self.getConverged():
            raise IndexError('eigenpair index out of range')
        if vr is not None:
            if not self._evecs:
                raise RuntimeError('eigenvectors were not kept (EPS.keep_vectors = False)')
            self._evecs[i].copy(vr)
        return complex(self._evals[i], 0.0)

    def getErrorEstimate(self, i):
        return float(self._errest[i])

    def destroy(self):
        for v in self._evecs:
            v.destroy()
        self._evecs = []
