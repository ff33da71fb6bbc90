"""
Subspace data holders and index maps (reference ``_backend/bsubspace.pyx:60-261``).

The ``C*`` classes own a ``dnm_subspace_t`` descriptor; the array functions
call the host entry points of the C ABI (``dnm_subspace_*``), which evaluate
the same ``__host__ __device__`` rank/unrank code the CUDA kernels use.
"""
import ctypes as C

import numpy as np

from .. import _capi
from .._capi import SubspaceDesc, as_i64, check, ip

dnm_int_t = np.int64


class SubspaceType:
    # numeric values of _backend/bsubspace_impl.h:17-23
    FULL = 0
    PARITY = 1
    EXPLICIT = 2
    SPIN_CONSERVE = 3


class _CData:
    def __init__(self):
        self.desc = SubspaceDesc()
        self._keep = []

    def _hold(self, a):
        a = as_i64(a)
        self._keep.append(a)
        return a


class CFull(_CData):
    def __init__(self, L):
        super().__init__()
        self.desc.type = SubspaceType.FULL
        self.desc.L = int(L)


class CParity(_CData):
    def __init__(self, L, space):
        super().__init__()
        self.desc.type = SubspaceType.PARITY
        self.desc.L = int(L)
        self.desc.space = int(space)


class CSpinConserve(_CData):
    def __init__(self, L, k, nchoosek):
        super().__init__()
        nck = self._hold(nchoosek)
        if nck.ndim != 2:
            raise ValueError('nchoosek must be a 2D array')
        self.desc.type = SubspaceType.SPIN_CONSERVE
        self.desc.L = int(L)
        self.desc.k = int(k)
        self.desc.ld_nchoosek = nck.shape[1]
        self.desc.nchoosek = ip(nck)


class CExplicit(_CData):
    def __init__(self, L, state_map, rmap_indices, rmap_states):
        super().__init__()
        smap = self._hold(state_map)
        ridx = self._hold(rmap_indices)
        rst = self._hold(rmap_states)
        self.desc.type = SubspaceType.EXPLICIT
        self.desc.L = int(L)
        self.desc.dim = smap.size
        self.desc.state_map = ip(smap)
        # sentinel: rmap_indices[0] == -1 means "state_map is sorted" (bsubspace.pyx:110-113)
        self.desc.rmap_indices = ip(ridx) if ridx[0] != -1 else None
        self.desc.rmap_states = ip(rst)


def _dim(data):
    out = C.c_int64()
    check(_capi.lib().dnm_subspace_dim(C.byref(data.desc), C.byref(out)))
    return out.value


def _i2s(idxs, data):
    idxs = as_i64(idxs)
    out = np.empty(idxs.size, dtype=dnm_int_t)
    check(_capi.lib().dnm_subspace_i2s(C.byref(data.desc), idxs.size, ip(idxs), ip(out)))
    return out


def _s2i(states, data):
    states = as_i64(states)
    out = np.empty(states.size, dtype=dnm_int_t)
    check(_capi.lib().dnm_subspace_s2i(C.byref(data.desc), states.size, ip(states), ip(out)))
    return out


get_dimension_Full = get_dimension_Parity = get_dimension_SpinConserve = get_dimension_Explicit = _dim
idx_to_state_Full = idx_to_state_Parity = idx_to_state_SpinConserve = idx_to_state_Explicit = _i2s
state_to_idx_Full = state_to_idx_Parity = state_to_idx_SpinConserve = state_to_idx_Explicit = _s2i


def compute_rcm(masks, signs, coeffs, state_map, start, L):
    """Fill ``state_map`` with the states reachable from ``start`` under the
    operator (breadth-first); returns how many were found
    (reference ``bsubspace.pyx:212-261``)."""
    masks, signs = as_i64(masks), as_i64(signs)
    coeffs = _capi.as_c128(coeffs)
    if not (state_map.dtype == np.int64 and state_map.flags.c_contiguous):
        raise ValueError('state_map must be a contiguous int64 array')
    out = C.c_int64()
    # big searches run on the GPU when one is initialised (same output, element for element);
    # DNM_RCM_DEVICE=0 / 1 forces the host / device version
    import os
    where = os.environ.get('DNM_RCM_DEVICE')
    on_device = _capi.lib().dnm_have_gpu() and (where == '1' or (where != '0' and state_map.size >= (1 << 16)))
    fn = _capi.lib().dnm_compute_rcm_device if on_device else _capi.lib().dnm_compute_rcm
    ierr = fn(masks.size, ip(masks), ip(signs),
                                       _capi.fp(coeffs), ip(state_map), state_map.size,
                                       int(start), int(L), C.byref(out))
    if ierr:
        raise RuntimeError(_capi.lib().dnm_last_error().decode())
    return out.value
