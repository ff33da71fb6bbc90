"""
Device-resident ``Vec`` and shell ``Mat`` with the slice of the petsc4py method
surface that dynamite's Python layer uses: the objects
``State.vec`` and ``Operator.get_mat()`` return.

A ``Vec`` owns a complex128 buffer in HBM (``dnm_vec_t``); host access goes
through explicit copies (``vec[a:b]``, ``vec[idx] = v``), everything else runs
as CUDA kernels on the library stream.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import BackendError, as_c128, as_i64, check, fp, ip

Error = BackendError


class NormType:
    NORM_1 = 1
    NORM_2 = 0
    NORM_INFINITY = 2
    INFINITY = 2
    FROBENIUS = 3


def garbage_cleanup():
    pass


class _Comm:
    @property
    def rank(self):
        r, n = C.c_int(), C.c_int()
        _capi.lib().dnm_comm_rank(C.byref(r), C.byref(n))
        return r.value

    @property
    def size(self):
        r, n = C.c_int(), C.c_int()
        _capi.lib().dnm_comm_rank(C.byref(r), C.byref(n))
        return n.value

    def getRank(self):
        return self.rank

    def getSize(self):
        return self.size

    def barrier(self):
        check(_capi.lib().dnm_comm_barrier())


COMM_WORLD = _Comm()


class Vec:
    """complex128 vector in device memory, block-distributed over the ranks."""

    def __init__(self, n=None, handle=None):
        self.handle = handle
        if n is not None:
            self.create(n)

    # -- lifetime ----------------------------------------------------------
    def create(self, n):
        _capi.ensure_gpu()
        h = C.c_void_p()
        check(_capi.lib().dnm_vec_create(int(n), C.byref(h)))
        self.handle = h
        return self

    def destroy(self):
        if self.handle is not None:
            _capi.lib().dnm_vec_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def duplicate(self):
        return Vec(self.getSize())

    # -- layout --------------------------------------------------------------
    def _sizes(self):
        n, a, b = C.c_int64(), C.c_int64(), C.c_int64()
        check(_capi.lib().dnm_vec_size(self.handle, C.byref(n), C.byref(a), C.byref(b)))
        return n.value, a.value, b.value

    def getSize(self):
        return self._sizes()[0]

    def getLocalSize(self):
        _, a, b = self._sizes()
        return b - a

    def getOwnershipRange(self):
        _, a, b = self._sizes()
        return a, b

    @property
    def device_ptr(self):
        return _capi.lib().dnm_vec_device_ptr(self.handle)

    # -- host access ---------------------------------------------------------
    def _local(self, key):
        """translate a global int / slice / index array into local form"""
        n, a, b = self._sizes()
        if isinstance(key, slice):
            start, stop, step = key.indices(n)
            if step != 1:
                raise IndexError('only contiguous slices are supported')
            if start < a or stop > b:
                raise IndexError(f'slice [{start}:{stop}) not owned by this rank [{a}:{b})')
            return 'slice', start - a, max(0, stop - start)
        if np.isscalar(key) or (isinstance(key, np.ndarray) and key.ndim == 0):
            k = int(key)
            if k < 0:
                k += n
            if not a <= k < b:
                raise IndexError(f'index {k} not owned by this rank [{a}:{b})')
            return 'slice', k - a, 1
        idx = as_i64(key) - a
        return 'index', idx, idx.size

    def __getitem__(self, key):
        kind, first, count = self._local(key)
        out = np.empty(count, dtype=np.complex128)
        if kind == 'slice':
            check(_capi.lib().dnm_vec_get_host(self.handle, first, count, fp(out)))
            scalar = np.isscalar(key) or (isinstance(key, np.ndarray) and key.ndim == 0)
            return out[0] if scalar else out
        check(_capi.lib().dnm_vec_get_values(self.handle, count, ip(first), fp(out)))
        return out

    def __setitem__(self, key, value):
        kind, first, count = self._local(key)
        vals = np.empty(count, dtype=np.complex128)
        vals[:] = value
        if kind == 'slice':
            check(_capi.lib().dnm_vec_set_host(self.handle, first, count, fp(vals)))
        else:
            check(_capi.lib().dnm_vec_set_values(self.handle, count, ip(first), fp(vals), 0))

    def setValues(self, idxs, values, addv=False):
        _, a, _ = self._sizes()
        idx = as_i64(np.atleast_1d(idxs)) - a
        vals = np.empty(idx.size, dtype=np.complex128)
        vals[:] = values
        check(_capi.lib().dnm_vec_set_values(self.handle, idx.size, ip(idx), fp(vals), int(bool(addv))))

    def getArray(self):
        """host copy of the local block"""
        return self[slice(*self.getOwnershipRange())]

    def assemblyBegin(self):
        pass

    def assemblyEnd(self):
        pass

    def assemble(self):
        pass

    # -- algebra (device) ------------------------------------------------------
    def set(self, value):
        value = complex(value)
        check(_capi.lib().dnm_vec_set(self.handle, value.real, value.imag))

    def setRandom(self, seed=0):
        """device-side pseudo-random fill (synthetic data; not numpy's stream)"""
        check(_capi.lib().dnm_vec_set_random(self.handle, int(seed)))

    def zeroEntries(self):
        self.set(0)

    def copy(self, result=None):
        if result is None:
            result = self.duplicate()
        check(_capi.lib().dnm_vec_copy(self.handle, result.handle))
        return result

    def scale(self, alpha):
        alpha = complex(alpha)
        check(_capi.lib().dnm_vec_scale(self.handle, alpha.real, alpha.imag))

    def axpby(self, alpha, beta, x):
        """self = alpha*x + beta*self"""
        alpha, beta = complex(alpha), complex(beta)
        check(_capi.lib().dnm_vec_axpby(self.handle, alpha.real, alpha.imag, beta.real, beta.imag,
                                        x.handle))

    def axpy(self, alpha, x):
        self.axpby(alpha, 1.0, x)

    def dot(self, other):
        """``VecDot(self, other)`` = sum_i self_i * conj(other_i)"""
        out = np.empty(2, dtype=np.float64)
        check(_capi.lib().dnm_vec_dot(self.handle, other.handle, fp(out)))
        return complex(out[0], out[1])

    def norm(self, norm_type=None):
        if norm_type is None:
            norm_type = NormType.NORM_2
        out = C.c_double()
        check(_capi.lib().dnm_vec_norm(self.handle, int(norm_type), C.byref(out)))
        return out.value

    def shift(self, alpha):
        """self[i] += alpha for every entry (petsc4py Vec.shift)"""
        alpha = complex(alpha)
        check(_capi.lib().dnm_vec_shift(self.handle, alpha.real, alpha.imag))

    def normalize(self):
        nrm = C.c_double()
        check(_capi.lib().dnm_vec_normalize(self.handle, C.byref(nrm)))
        return nrm.value

    def equal(self, other):
        if self.getSize() != other.getSize():
            return False
        return bool(np.array_equal(self.getArray(), other.getArray()))

    def __sub__(self, other):
        out = self.copy()
        out.axpby(-1.0, 1.0, other)
        return out

    def __add__(self, other):
        out = self.copy()
        out.axpby(1.0, 1.0, other)
        return out

    def __len__(self):
        return self.getSize()


class Mat:
    """Matrix-free MSC operator on the device (a PETSc shell ``Mat`` in the
    reference, ``_backend/bcuda_template_2.cu:4-44``)."""

    def __init__(self, handle):
        self.handle = handle

    def getSize(self):
        m, n = C.c_int64(), C.c_int64()
        check(_capi.lib().dnm_mat_size(self.handle, C.byref(m), C.byref(n)))
        return m.value, n.value

    def createVecs(self):
        m, n = self.getSize()
        return Vec(n), Vec(m)

    def mult(self, x, y):
        """y = A x (MATOP_MULT)"""
        check(_capi.lib().dnm_mat_mult(self.handle, x.handle, y.handle))

    def mult_host(self, x, y):
        """same product with host numpy buffers (H2D, multiply, D2H)"""
        if not (x.dtype == np.complex128 and y.dtype == np.complex128
                and x.flags.c_contiguous and y.flags.c_contiguous):
            raise ValueError('host buffers must be contiguous complex128')
        check(_capi.lib().dnm_mat_mult_host(self.handle, x.ctypes.data, y.ctypes.data))

    def mult_host_batch(self, xs, ys):
        """``ys[k] = A xs[k]`` for host numpy buffers, the copies of consecutive products overlapped with
        one another and with the multiplies (``dnm_mat_mult_host_batch``); pinned buffers for full overlap."""
        if len(xs) != len(ys):
            raise ValueError('as many outputs as inputs are needed')
        for a in list(xs) + list(ys):
            if not (a.dtype == np.complex128 and a.flags.c_contiguous):
                raise ValueError('host buffers must be contiguous complex128')
        for y in ys:
            if any(np.may_share_memory(y, x) for x in xs):
                raise ValueError('an output buffer aliases an input buffer')
        n = len(xs)
        xp = (C.c_void_p * max(n, 1))(*[x.ctypes.data for x in xs])
        yp = (C.c_void_p * max(n, 1))(*[y.ctypes.data for y in ys])
        check(_capi.lib().dnm_mat_mult_host_batch(self.handle, n, xp, yp))

    def norm(self, norm_type=None):
        if norm_type not in (None, NormType.INFINITY):
            raise BackendError(1, 'Only NORM_INFINITY is implemented for shell matrices.')
        out = C.c_double()
        check(_capi.lib().dnm_mat_norm_inf(self.handle, C.byref(out)))
        return out.value

    def set_option(self, key, value):
        check(_capi.lib().dnm_mat_set_option(self.handle, key.encode(), int(value)))

    def get_info(self, key):
        out = C.c_double()
        check(_capi.lib().dnm_mat_get_info(self.handle, key.encode(), C.byref(out)))
        return out.value

    def destroy(self):
        if self.handle is not None:
            _capi.lib().dnm_mat_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
