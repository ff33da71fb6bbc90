"""
Matrix construction and the native routines that take a Mat or Vec
(reference ``_backend/bpetsc.pyx:78-276``).
"""
import ctypes as C

import numpy as np

from .. import _capi
from .._capi import as_c128, as_i64, check, fp, ip
from ..petsc import Mat, Vec


def build_mat(masks, mask_offsets, signs, coeffs, left_subspace, right_subspace,
              xparity, shell, gpu):
    """``bpetsc.build_mat`` (``bpetsc.pyx:78-138``) -> :class:`dynamite_b200.petsc.Mat`.

    ``left_subspace`` / ``right_subspace`` are ``{'type': SubspaceType, 'data': C*}``
    dicts as produced by ``Subspace._to_c()``.  Only shell GPU matrices exist here.
    """
    if not shell:
        raise RuntimeError('dynamite_b200 only builds shell (matrix-free) matrices; '
                           'set Operator.shell = True')
    if not gpu:
        raise RuntimeError('dynamite_b200 has no CPU shell implementation')
    _capi.ensure_gpu()
    masks, mask_offsets, signs = as_i64(masks), as_i64(mask_offsets), as_i64(signs)
    coeffs = as_c128(coeffs)
    handle = C.c_void_p()
    check(_capi.lib().dnm_mat_create(
        masks.size, ip(masks), ip(mask_offsets), ip(signs), fp(coeffs),
        C.byref(left_subspace['data'].desc), C.byref(right_subspace['data'].desc),
        int(bool(xparity)), C.byref(handle)))
    return Mat(handle)


def precompute_diagonal(A):
    """``bpetsc.precompute_diagonal`` (``bpetsc.pyx:141-147``)"""
    check(_capi.lib().dnm_mat_precompute_diagonal(A.handle))


def check_conserves(masks, mask_offsets, signs, coeffs, left_subspace, right_subspace, xparity):
    """``bpetsc.check_conserves`` (``bpetsc.pyx:150-193``)"""
    _capi.ensure_gpu()
    masks, mask_offsets, signs = as_i64(masks), as_i64(mask_offsets), as_i64(signs)
    coeffs = as_c128(coeffs)
    out = C.c_int()
    check(_capi.lib().dnm_check_conserves(
        masks.size, ip(masks), ip(mask_offsets), ip(signs), fp(coeffs),
        C.byref(left_subspace['data'].desc), C.byref(right_subspace['data'].desc),
        int(bool(xparity)), C.byref(out)))
    return bool(out.value)


def reduced_density_matrix(v, subspace, keep, triang=True):
    """``bpetsc.reduced_density_matrix`` (``bpetsc.pyx:245-276``).  ``triang`` is
    accepted and ignored, as in the reference (``bpetsc_template_1.c:92``).
    The matrix is returned on every rank."""
    keep = as_i64(keep)
    d = 1 << keep.size
    out = np.zeros((d, d), dtype=np.complex128, order='C')
    check(_capi.lib().dnm_rdm(v.handle, C.byref(subspace['data'].desc), keep.size, ip(keep), fp(out)))
    return out


def track_memory():
    pass


def get_max_memory_usage():
    return get_cur_memory_usage()


def get_cur_memory_usage():
    """device bytes in use (reference reports host RSS; the state lives in HBM here)"""
    free, total = C.c_int64(), C.c_int64()
    check(_capi.lib().dnm_mem_info(C.byref(free), C.byref(total)))
    return float(total.value - free.value)
