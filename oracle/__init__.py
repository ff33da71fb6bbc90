"""
CPU oracle for the dynamite MSC shell path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``dynamite_b200/`` imports this package.  It is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` as the checker / timed CPU baseline.

Two layers:

* ``liboracle.so`` (``dnm_oracle.c``): plain-C restatement of the reference's
  native algorithms, one function per reference routine (citations there).
* numpy restatements in this file of the *definition* of the MSC format
  (``/root/reference/src/dynamite/msc_tools.py:63-80``) and brute-force
  definitions of the subspaces, which are independent of the C code and are
  used to pin it.

Parity pin: see ``dnm_oracle.h``.  The SLEPc Krylov algorithms are third-party
(slepc 3.20.2, absent here); for evolve/eigsolve the oracle is scipy
(``expm_multiply`` / dense ``eigh``), as in the reference's own tests
(``tests/integration/test_evolve.py:35-57``, ``test_eigsolve.py:17-88``).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FULL, PARITY, EXPLICIT, SPIN_CONSERVE = 0, 1, 2, 3
_TYPES = {'full': FULL, 'parity': PARITY, 'explicit': EXPLICIT, 'spinconserve': SPIN_CONSERVE}

_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


class _Sub(C.Structure):
    _fields_ = [('type', C.c_int32), ('L', C.c_int64), ('space', C.c_int64), ('k', C.c_int64),
                ('ld_nchoosek', C.c_int64), ('nchoosek', _i64p), ('dim', C.c_int64),
                ('state_map', _i64p), ('rmap_indices', _i64p), ('rmap_states', _i64p)]


class _Msc(C.Structure):
    _fields_ = [('nmasks', C.c_int64), ('masks', _i64p), ('mask_offsets', _i64p),
                ('signs', _i64p), ('coeffs', _f64p)]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, '_build', 'liboracle.so')
    src = [os.path.join(_HERE, f) for f in ('dnm_oracle.c', 'dnm_oracle.h')]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        sp = C.POINTER(_Sub)
        mp = C.POINTER(_Msc)
        L.orc_dim.restype = C.c_int64
        L.orc_dim.argtypes = [sp]
        L.orc_s2i_array.argtypes = [sp, C.c_int64, _i64p, _i64p]
        L.orc_i2s_array.argtypes = [sp, C.c_int64, _i64p, _i64p]
        L.orc_next_state.restype = C.c_int64
        L.orc_next_state.argtypes = [sp, C.c_int64, C.c_int64]
        L.orc_matmult.argtypes = [mp, sp, sp, C.c_int, _f64p, _f64p, _f64p]
        L.orc_precompute_diag.argtypes = [mp, sp, C.c_int, _f64p]
        L.orc_precompute_diag_range.argtypes = [mp, sp, C.c_int64, C.c_int64, _f64p]
        L.orc_norm_inf.argtypes = [mp, sp, sp, C.c_int, _f64p]
        L.orc_check_conserves.argtypes = [mp, sp, sp, C.c_int, C.POINTER(C.c_int)]
        L.orc_rdm.argtypes = [_f64p, sp, C.c_int64, _i64p, C.c_int64, _f64p]
        L.orc_compute_rcm.restype = C.c_int64
        L.orc_compute_rcm.argtypes = [C.c_int64, _i64p, _i64p, _f64p, _i64p, C.c_int64,
                                      C.c_int64, C.c_int64]
        L.orc_matmult_fast.argtypes = [mp, sp, _f64p, _f64p, _f64p, C.c_int]
        L.orc_matmult_fast_range.argtypes = [mp, sp, _f64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_int]
        _LIB = L
    return _LIB


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ip(a):
    return a.ctypes.data_as(_i64p)


def _fp(a):
    return a.ctypes.data_as(_f64p)


class Subspace:
    """Oracle-side subspace descriptor built from a plain spec:

    ``{'type': 'full'|'parity'|'spinconserve'|'explicit', 'L': .., 'space': ..,
    'k': .., 'states': [...]}``
    """

    def __init__(self, spec):
        self.spec = dict(spec)
        t = _TYPES[spec['type']]
        s = _Sub()
        s.type = t
        s.L = int(spec['L'])
        self._keep = []
        if t == PARITY:
            s.space = int(spec['space'])
        elif t == SPIN_CONSERVE:
            from math import comb
            k, L = int(spec['k']), int(spec['L'])
            # table layout per /root/reference/src/dynamite/subspaces.py:341-352
            tab = _i64([[comb(n, kk) for n in range(L + 1)] for kk in range(k + 1)])
            s.k = k
            s.ld_nchoosek = L + 1
            s.nchoosek = _ip(tab)
            self._keep.append(tab)
        elif t == EXPLICIT:
            # rmap construction per subspaces.py:391-399
            smap = _i64(spec['states'])
            s.dim = smap.size
            if np.all(smap[:-1] <= smap[1:]):
                rstates = smap
                s.rmap_indices = None
            else:
                order = _i64(np.argsort(smap, kind='stable'))
                rstates = _i64(smap[order])
                s.rmap_indices = _ip(order)
                self._keep.append(order)
            s.state_map = _ip(smap)
            s.rmap_states = _ip(rstates)
            self._keep += [smap, rstates]
        self.c = s

    @property
    def dim(self):
        return int(lib().orc_dim(C.byref(self.c)))

    def s2i(self, states):
        states = _i64(np.atleast_1d(states))
        out = np.empty_like(states)
        lib().orc_s2i_array(C.byref(self.c), states.size, _ip(states), _ip(out))
        return out

    def i2s(self, idxs):
        idxs = _i64(np.atleast_1d(idxs))
        out = np.empty_like(idxs)
        lib().orc_i2s_array(C.byref(self.c), idxs.size, _ip(idxs), _ip(out))
        return out

    def next_state(self, prev, idx):
        return int(lib().orc_next_state(C.byref(self.c), int(prev), int(idx)))


class Msc:
    """Oracle-side MSC in the backend's 'CSR by mask' layout
    (/root/reference/src/dynamite/_backend/shell_context.h:4-10)."""

    def __init__(self, masks, mask_offsets, signs, coeffs):
        self.masks = _i64(masks)
        self.mask_offsets = _i64(mask_offsets)
        self.signs = _i64(signs)
        self.coeffs = np.ascontiguousarray(coeffs, dtype=np.complex128)
        m = _Msc()
        m.nmasks = self.masks.size
        m.masks = _ip(self.masks)
        m.mask_offsets = _ip(self.mask_offsets)
        m.signs = _ip(self.signs)
        m.coeffs = self.coeffs.view(np.float64).ctypes.data_as(_f64p)
        self.c = m

    @classmethod
    def from_terms(cls, terms):
        """terms: structured/record array or list of (mask, sign, coeff), sorted by mask."""
        masks = _i64([t[0] for t in terms])
        signs = _i64([t[1] for t in terms])
        coeffs = np.array([t[2] for t in terms], dtype=np.complex128)
        um, first = np.unique(masks, return_index=True)
        offs = np.append(first, masks.size)
        return cls(um, offs, signs, coeffs)

    def flat_masks(self):
        return np.repeat(self.masks, np.diff(self.mask_offsets))


def _dims(left, right, xparity):
    M, N = left.dim, right.dim
    return (M // 2, N // 2) if xparity else (M, N)


def matmult(msc, left, right, x, xparity=False, diag=None):
    M, N = _dims(left, right, xparity)
    x = np.ascontiguousarray(x, dtype=np.complex128)
    assert x.size == N
    y = np.empty(M, dtype=np.complex128)
    d = None
    if diag is not None:
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        d = _fp(diag)
    lib().orc_matmult(C.byref(msc.c), C.byref(left.c), C.byref(right.c), int(xparity), d,
                      _fp(x.view(np.float64)), _fp(y.view(np.float64)))
    return y


def matmult_fast(msc, sub, x, diag=None, nthreads=1):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.empty(sub.dim, dtype=np.complex128)
    d = None
    if diag is not None:
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        d = _fp(diag)
    used = lib().orc_matmult_fast(C.byref(msc.c), C.byref(sub.c), d, _fp(x.view(np.float64)),
                                  _fp(y.view(np.float64)), int(nthreads))
    if used < 0:
        raise ValueError('fast path needs Full/Parity and dim >= 2048')
    return y, used


def matmult_fast_range(msc, sub, x, y, blk_first, blk_last, diag=None, nthreads=1):
    """Rows [2048*blk_first, 2048*blk_last) of y = A x with the fast path, in place in ``y``
    (complex128, full length).  Used by bench.py to time a bounded sample of a big multiply."""
    assert x.dtype == np.complex128 and y.dtype == np.complex128 and x.size == y.size == sub.dim
    d = None
    if diag is not None:
        assert diag.dtype == np.float64
        d = _fp(diag)
    used = lib().orc_matmult_fast_range(C.byref(msc.c), C.byref(sub.c), d, _fp(x.view(np.float64)),
                                        _fp(y.view(np.float64)), int(blk_first), int(blk_last), int(nthreads))
    if used < 0:
        raise ValueError('fast path needs Full/Parity, dim >= 2048 and a non-empty block range')
    return used


def precompute_diag(msc, sub, xparity=False):
    M = sub.dim // 2 if xparity else sub.dim
    d = np.empty(M, dtype=np.float64)
    rc = lib().orc_precompute_diag(C.byref(msc.c), C.byref(sub.c), int(xparity), _fp(d))
    return None if rc else d


def precompute_diag_range(msc, sub, diag, row_first, row_last):
    """fill diag[row_first:row_last] in place (float64, full length)"""
    assert diag.dtype == np.float64
    return lib().orc_precompute_diag_range(C.byref(msc.c), C.byref(sub.c), int(row_first), int(row_last), _fp(diag))


def norm_inf(msc, left, right, xparity=False):
    out = C.c_double()
    lib().orc_norm_inf(C.byref(msc.c), C.byref(left.c), C.byref(right.c), int(xparity),
                       C.byref(out))
    return out.value


def check_conserves(msc, left, right, xparity=False):
    out = C.c_int()
    lib().orc_check_conserves(C.byref(msc.c), C.byref(left.c), C.byref(right.c), int(xparity),
                              C.byref(out))
    return bool(out.value)


def rdm(x, sub, keep):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    keep = _i64(keep)
    d = 1 << keep.size
    out = np.empty((d, d), dtype=np.complex128)
    rc = lib().orc_rdm(_fp(x.view(np.float64)), C.byref(sub.c), keep.size, _ip(keep), d,
                       _fp(out.view(np.float64)))
    if rc:
        raise ValueError('keep array must be strictly increasing')
    return out


def compute_rcm(masks, signs, coeffs, start, L, max_states=None):
    masks, signs = _i64(masks), _i64(signs)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.complex128)
    if max_states is None:
        max_states = 1 << L
    smap = np.empty(max_states, dtype=np.int64)
    n = lib().orc_compute_rcm(masks.size, _ip(masks), _ip(signs), _fp(coeffs.view(np.float64)),
                              _ip(smap), max_states, int(start), int(L))
    if n < 0:
        raise RuntimeError('state_map size too small')
    return smap[:n].copy()


# ----------------------------------------------------------------------------
# numpy restatements, independent of the C code
# ----------------------------------------------------------------------------

def parity(v):
    """popcount parity of int64 array(s)."""
    v = np.array(v, dtype=np.uint64, copy=True)
    for s in (32, 16, 8, 4, 2, 1):
        v ^= v >> np.uint64(s)
    return (v & np.uint64(1)).astype(np.int64)


def brute_states(spec):
    """The subspace as an explicit list of states, from its *definition*:
    Parity / SpinConserve enumerate the states with the given popcount parity /
    popcount in increasing integer order (pinned by the KATs in
    /root/reference/tests/unit/test_subspaces.py:140-190, 294-342)."""
    L = spec['L']
    t = spec['type']
    if t == 'explicit':
        return _i64(spec['states'])
    allst = np.arange(1 << L, dtype=np.int64)
    if t == 'full':
        return allst
    if t == 'parity':
        return allst[parity(allst) == spec['space']]
    if t == 'spinconserve':
        pc = np.array([bin(int(s)).count('1') for s in allst])
        return allst[pc == spec['k']]
    raise ValueError(t)


def msc_to_dense(terms, left_states, right_states):
    """Dense matrix of an MSC operator between two explicit state lists.

    Definition (/root/reference/src/dynamite/msc_tools.py:63-80):
    ``A[row, S2I_R(I2S_L(row) ^ mask)] += coeff * (-1)^popcount(sign & (I2S_L(row) ^ mask))``
    """
    left_states = _i64(left_states)
    right_states = _i64(right_states)
    lookup = {int(s): i for i, s in enumerate(right_states)}
    A = np.zeros((left_states.size, right_states.size), dtype=np.complex128)
    for row, ket in enumerate(left_states):
        for (m, s, c) in terms:
            bra = int(ket) ^ int(m)
            col = lookup.get(bra)
            if col is None:
                continue
            A[row, col] += complex(c) * (1 - 2 * (bin(bra & int(s)).count('1') & 1))
    return A


def rdm_dense(psi, L, keep):
    """rho_keep = Tr_rest |psi><psi| by reshape-and-contract (Full space), as in
    /root/reference/tests/integration/test_rdm.py:242-272."""
    keep = list(keep)
    rest = [i for i in range(L) if i not in keep]
    t = np.asarray(psi, dtype=np.complex128).reshape([2] * L)  # axis a <-> spin L-1-a
    axes = [L - 1 - i for i in reversed(keep)] + [L - 1 - i for i in reversed(rest)]
    t = t.transpose(axes).reshape(1 << len(keep), 1 << len(rest))
    return t @ t.conj().T
