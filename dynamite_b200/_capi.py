"""
ctypes binding of ``libdynamite_b200.so`` (C ABI declared in ``include/dynamite_b200.h``).

This plays the role of dynamite's Cython layer (``_backend/bpetsc.pyx``,
``bsubspace.pyx``): numpy arrays in, opaque handles out, and a non-zero return
code becomes a Python exception (``bpetsc.pyx:135-136`` raises
``petsc4py.PETSc.Error(ierr)``; here :class:`BackendError`).

There is no CPU fallback: if the shared library is missing or no B200-class
device can be bound, the compute entry points raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DNM_LIB overrides the library path (A/B timing of builds); normally the in-tree build is used
LIB_PATH = os.environ.get('DNM_LIB') or os.path.join(_HERE, 'libdynamite_b200.so')

i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)


class BackendError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status
    (the role of ``petsc4py.PETSc.Error``)."""

    def __init__(self, ierr, message=''):
        self.ierr = ierr
        super().__init__(f'dynamite_b200 backend error {ierr}: {message}')


class SubspaceDesc(C.Structure):
    """``dnm_subspace_t``"""
    _fields_ = [('type', C.c_int32), ('L', C.c_int64), ('space', C.c_int64), ('k', C.c_int64),
                ('ld_nchoosek', C.c_int64), ('nchoosek', i64p), ('dim', C.c_int64),
                ('state_map', i64p), ('rmap_indices', i64p), ('rmap_states', i64p)]


_sp = C.POINTER(SubspaceDesc)
_vec = C.c_void_p
_mat = C.c_void_p

# name -> (restype, argtypes); every symbol the header declares
_SIGNATURES = {
    'dnm_init': (C.c_int, [C.c_int]),
    'dnm_finalize': (C.c_int, []),
    'dnm_last_error': (C.c_char_p, []),
    'dnm_have_gpu': (C.c_int, []),
    'dnm_device_count': (C.c_int, [C.POINTER(C.c_int)]),
    'dnm_stream': (C.c_void_p, []),
    'dnm_synchronize': (C.c_int, []),
    'dnm_timer_start': (C.c_int, []),
    'dnm_timer_stop': (C.c_int, [C.POINTER(C.c_float)]),
    'dnm_mem_info': (C.c_int, [i64p, i64p]),
    'dnm_launch_count': (C.c_int64, [C.c_int]),
    'dnm_host_alloc': (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    'dnm_host_free': (C.c_int, [C.c_void_p]),
    'dnm_comm_unique_id': (C.c_int, [C.c_char_p]),
    'dnm_comm_init': (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    'dnm_comm_rank': (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'dnm_comm_barrier': (C.c_int, []),
    'dnm_shard_plan': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, i64p, C.POINTER(C.c_int32), i64p]),
    'dnm_subspace_dim': (C.c_int, [_sp, i64p]),
    'dnm_subspace_s2i': (C.c_int, [_sp, C.c_int64, i64p, i64p]),
    'dnm_subspace_i2s': (C.c_int, [_sp, C.c_int64, i64p, i64p]),
    'dnm_compute_rcm': (C.c_int, [C.c_int64, i64p, i64p, f64p, i64p, C.c_int64, C.c_int64,
                                  C.c_int64, i64p]),
    'dnm_jit_dryrun': (C.c_int, [C.c_int64, i64p, i64p, i64p, f64p, _sp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_char_p, C.c_int64, i64p, i64p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'dnm_jit_set_host_emulation': (C.c_int, [C.c_int]),
    'dnm_compute_rcm_device': (C.c_int, [C.c_int64, i64p, i64p, f64p, i64p, C.c_int64, C.c_int64,
                                  C.c_int64, i64p]),
    'dnm_subspace_s2i_device': (C.c_int, [_sp, C.c_int64, i64p, i64p]),
    'dnm_subspace_i2s_device': (C.c_int, [_sp, C.c_int64, i64p, i64p]),
    'dnm_vec_create': (C.c_int, [C.c_int64, C.POINTER(_vec)]),
    'dnm_vec_destroy': (C.c_int, [_vec]),
    'dnm_vec_size': (C.c_int, [_vec, i64p, i64p, i64p]),
    'dnm_vec_device_ptr': (C.c_void_p, [_vec]),
    'dnm_vec_set_host': (C.c_int, [_vec, C.c_int64, C.c_int64, f64p]),
    'dnm_vec_get_host': (C.c_int, [_vec, C.c_int64, C.c_int64, f64p]),
    'dnm_vec_get_host_global': (C.c_int, [_vec, C.c_int64, C.c_int64, f64p]),
    'dnm_vec_set_values': (C.c_int, [_vec, C.c_int64, i64p, f64p, C.c_int]),
    'dnm_vec_get_values': (C.c_int, [_vec, C.c_int64, i64p, f64p]),
    'dnm_vec_set': (C.c_int, [_vec, C.c_double, C.c_double]),
    'dnm_vec_set_random': (C.c_int, [_vec, C.c_uint64]),
    'dnm_vec_copy': (C.c_int, [_vec, _vec]),
    'dnm_vec_scale': (C.c_int, [_vec, C.c_double, C.c_double]),
    'dnm_vec_shift': (C.c_int, [_vec, C.c_double, C.c_double]),
    'dnm_vec_normalize': (C.c_int, [_vec, C.POINTER(C.c_double)]),
    'dnm_vec_axpby': (C.c_int, [_vec, C.c_double, C.c_double, C.c_double, C.c_double, _vec]),
    'dnm_vec_dot': (C.c_int, [_vec, _vec, f64p]),
    'dnm_vec_norm': (C.c_int, [_vec, C.c_int, f64p]),
    'dnm_mat_create': (C.c_int, [C.c_int64, i64p, i64p, i64p, f64p, _sp, _sp, C.c_int,
                                 C.POINTER(_mat)]),
    'dnm_mat_precompute_diagonal': (C.c_int, [_mat]),
    'dnm_mat_mult': (C.c_int, [_mat, _vec, _vec]),
    'dnm_mat_mult_host': (C.c_int, [_mat, C.c_void_p, C.c_void_p]),
    'dnm_mat_mult_host_batch': (C.c_int, [_mat, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    'dnm_mat_norm_inf': (C.c_int, [_mat, f64p]),
    'dnm_mat_size': (C.c_int, [_mat, i64p, i64p]),
    'dnm_mat_destroy': (C.c_int, [_mat]),
    'dnm_mat_set_option': (C.c_int, [_mat, C.c_char_p, C.c_int64]),
    'dnm_mat_get_info': (C.c_int, [_mat, C.c_char_p, f64p]),
    'dnm_check_conserves': (C.c_int, [C.c_int64, i64p, i64p, i64p, f64p, _sp, _sp, C.c_int,
                                      C.POINTER(C.c_int)]),
    'dnm_evolve': (C.c_int, [_mat, _vec, _vec, C.c_double, C.c_double, C.c_double, C.c_int,
                             C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'dnm_evolve_algo': (C.c_int, [_mat, _vec, _vec, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'dnm_evolve_last_algo': (C.c_int, []),
    'dnm_eigsolve': (C.c_int, [_mat, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64,
                               C.c_int, C.POINTER(C.c_int), f64p, f64p, C.POINTER(_vec),
                               C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'dnm_rdm': (C.c_int, [_vec, _sp, C.c_int64, i64p, f64p]),
}

_LIB = None


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load the shared library (once).  Raises if it has not been built:
    run ``python -c 'import __graft_entry__ as g; g.build()'`` or
    ``make -C dynamite_b200/csrc``."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} not found: build the CUDA extension first '
                '(make -C dynamite_b200/csrc). dynamite_b200 has no CPU fallback.')
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(ierr):
    if ierr != 0:
        msg = lib().dnm_last_error()
        raise BackendError(ierr, msg.decode() if msg else '')


def as_i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def as_c128(a):
    return np.ascontiguousarray(a, dtype=np.complex128)


def ip(a):
    return a.ctypes.data_as(i64p)


def fp(a):
    """pointer to the doubles of a float64 or complex128 array"""
    return a.ctypes.data_as(f64p)


_initialized_device = None


def ensure_gpu(device=None):
    """Bind this process to a GPU (``dnm_init``).  With torchrun the device is
    ``LOCAL_RANK``.  Raises :class:`BackendError` when there is no device."""
    global _initialized_device
    if _initialized_device is not None:
        return _initialized_device
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0'))
    check(lib().dnm_init(int(device)))
    _initialized_device = device
    return device


def gpu_available():
    """True if the library is built and a CUDA device is visible (no init)."""
    try:
        n = C.c_int(0)
        lib().dnm_device_count(C.byref(n))
        return n.value > 0
    except (ImportError, OSError):
        return False
