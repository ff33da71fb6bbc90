"""
Host-side helpers for the MSC (mask, sign, coefficient) operator format.

A term ``(mask, sign, coeff)`` is the matrix ``coeff * X^mask * Z^sign``:
``A[col ^ mask, col] = coeff * (-1)^popcount(sign & col)`` -- the definition
given by ``msc_to_numpy`` in the reference
(``/root/reference/src/dynamite/msc_tools.py:19-92``).  The structured dtype
has the same field names as the reference's ``msc_dtype`` because it is the
data format that crosses the backend boundary.

Only what the shell-matrix path needs is here (building the benchmark
Hamiltonians, sorting/merging terms, the Hermiticity test); this is numpy on
the host, as in the reference.
"""
import numpy as np
import scipy.sparse

dnm_int_t = np.int64

msc_dtype = np.dtype([('masks', dnm_int_t), ('signs', dnm_int_t), ('coeffs', np.complex128)])


def parity(v):
    """popcount parity, elementwise, for non-negative int64."""
    v = np.array(v, dtype=np.uint64, copy=True)
    for s in (32, 16, 8, 4, 2, 1):
        v ^= v >> np.uint64(s)
    return (v & np.uint64(1)).astype(np.int64)


def make_msc(terms):
    """structured array from an iterable of ``(mask, sign, coeff)``."""
    if isinstance(terms, np.ndarray) and terms.dtype == msc_dtype:
        return terms
    terms = list(terms)
    out = np.zeros(len(terms), dtype=msc_dtype)
    for i, (m, s, c) in enumerate(terms):
        out[i] = (m, s, c)
    return out


def combine_and_sort(msc):
    """Sort by (mask, sign), merge equal (mask, sign) pairs, drop zero terms
    (same contract as reference ``msc_tools.py:225-252``)."""
    msc = make_msc(msc)
    if msc.size == 0:
        return msc.copy()
    order = np.lexsort((msc['signs'], msc['masks']))
    m, s, c = msc['masks'][order], msc['signs'][order], msc['coeffs'][order]
    new_group = np.ones(m.size, dtype=bool)
    new_group[1:] = (m[1:] != m[:-1]) | (s[1:] != s[:-1])
    starts = np.flatnonzero(new_group)
    out = np.zeros(starts.size, dtype=msc_dtype)
    out['masks'] = m[starts]
    out['signs'] = s[starts]
    out['coeffs'] = np.add.reduceat(c, starts)
    return out[out['coeffs'] != 0]


def msc_sum(parts):
    """operator sum: concatenate the term lists and merge."""
    parts = [make_msc(p) for p in parts]
    if not parts:
        return np.zeros(0, dtype=msc_dtype)
    return combine_and_sort(np.concatenate(parts))


def msc_product(parts):
    """operator product ``parts[0] @ parts[1] @ ...``.

    ``(m1,s1,c1)(m2,s2,c2) = (m1^m2, s1^s2, c1*c2*(-1)^popcount(s1 & m2))``:
    moving ``Z^s1`` through ``X^m2`` costs a sign per overlapping bit.
    """
    parts = [make_msc(p) for p in parts]
    acc = parts[0]
    for nxt in parts[1:]:
        if acc.size == 1 and nxt.size == 1:
            # single Pauli strings (Majorana products: ~10^5 of these for SYK): plain integers
            m1, s1, c1 = int(acc['masks'][0]), int(acc['signs'][0]), complex(acc['coeffs'][0])
            m2, s2, c2 = int(nxt['masks'][0]), int(nxt['signs'][0]), complex(nxt['coeffs'][0])
            c = c1 * c2
            if bin(s1 & m2).count('1') & 1:
                c = -c
            acc = np.zeros(1 if c != 0 else 0, dtype=msc_dtype)
            if c != 0:
                acc[0] = (m1 ^ m2, s1 ^ s2, c)
            continue
        m = acc['masks'][:, None] ^ nxt['masks'][None, :]
        s = acc['signs'][:, None] ^ nxt['signs'][None, :]
        sgn = 1 - 2 * parity(acc['signs'][:, None] & nxt['masks'][None, :])
        c = acc['coeffs'][:, None] * nxt['coeffs'][None, :] * sgn
        out = np.zeros(m.size, dtype=msc_dtype)
        out['masks'], out['signs'], out['coeffs'] = m.ravel(), s.ravel(), c.ravel()
        acc = combine_and_sort(out)
    return acc


def shift(msc, by, wrap_L=None):
    """Translate the operator ``by`` sites to higher index.  With ``wrap_L``
    bits pushed past site L-1 re-enter at site 0 (periodic boundary)."""
    msc = make_msc(msc).copy()
    for field in ('masks', 'signs'):
        v = msc[field] << by
        if wrap_L is not None:
            full = (1 << wrap_L) - 1
            v = (v & full) | (v >> wrap_L)
        msc[field] = v
    return msc


def max_spin_idx(msc):
    msc = make_msc(msc)
    if msc.size == 0:
        return -1
    top = int(np.max(msc['masks'] | msc['signs']))
    return top.bit_length() - 1


def is_hermitian(msc):
    """Term-by-term Hermiticity: a term with odd popcount(mask & sign) must be
    purely imaginary, any other purely real (reference ``msc_tools.py:94-118``)."""
    msc = make_msc(msc)
    odd = parity(msc['masks'] & msc['signs']) == 1
    return not (np.any(msc['coeffs'][odd].real != 0) or np.any(msc['coeffs'][~odd].imag != 0))


def mask_offsets(msc):
    """unique masks and CSR-style offsets of a sorted MSC array
    (reference ``operators.py:653-669``)."""
    msc = make_msc(msc)
    if msc.size and np.any(np.diff(msc['masks']) < 0):
        raise ValueError('msc must be sorted first')
    masks, first = np.unique(msc['masks'], return_index=True)
    offsets = np.empty(first.size + 1, dtype=dnm_int_t)
    offsets[:-1] = first
    offsets[-1] = msc.size
    return masks.astype(dnm_int_t), offsets


def msc_to_numpy(msc, dims, idx_to_state=None, state_to_idx=None, sparse=True):
    """Matrix of an MSC operator (the format's definition, reference
    ``msc_tools.py:19-92``); vectorised over rows here.  ``idx_to_state`` maps
    row indices to states of the left subspace, ``state_to_idx`` maps states to
    column indices of the right subspace (-1 = not in the subspace)."""
    msc = make_msc(msc)
    rows = np.arange(dims[0], dtype=dnm_int_t)
    kets = rows if idx_to_state is None else np.asarray(idx_to_state(rows), dtype=dnm_int_t)
    data, ri, ci = [], [], []
    for m, s, c in msc:
        bras = kets ^ m
        cols = bras if state_to_idx is None else np.asarray(state_to_idx(bras), dtype=dnm_int_t)
        ok = cols != -1
        data.append(c * (1 - 2 * parity(bras[ok] & s)))
        ri.append(rows[ok])
        ci.append(cols[ok])
    if data:
        data, ri, ci = np.concatenate(data), np.concatenate(ri), np.concatenate(ci)
    else:
        data, ri, ci = np.zeros(0, complex), np.zeros(0, int), np.zeros(0, int)
    mat = scipy.sparse.csc_matrix((data, (ri, ci)), shape=dims, dtype=np.complex128)
    return mat if sparse else mat.toarray()


def nnz(msc):
    """number of distinct flip masks = non-zeros per row in the full space."""
    return int(np.unique(make_msc(msc)['masks']).size)
