"""
Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (where /root/reference exists):

    python tests/golden/generate_golden.py

The reference's Python layer is imported unmodified from /root/reference/src;
its compiled backend (PETSc/SLEPc/Cython, absent here) is replaced by stubs
that only provide `dnm_int_t`.  Everything numerical in the fixtures comes
from reference code or from numpy/scipy applied to reference output:

  * MSC arrays of the benchmark/test Hamiltonians: reference `operators.py`
    (`benchmarking/benchmark.py:129-178` definitions), incl. XParity-reduced
    versions from reference `subspaces.XParity.reduce_msc`.
  * dense matrices: reference `msc_tools.msc_to_numpy` (the definition of MSC)
    with index maps given by brute-force enumeration of each subspace
    (sorted states with the defining property).
  * y = A x, infinity norms, diagonals: numpy on those matrices.
  * evolve: scipy.sparse.linalg.expm_multiply; eigsolve: numpy.linalg.eigvalsh
    (the reference's own test oracles, tests/integration/test_evolve.py:35-57,
    test_eigsolve.py).
  * KATs copied as data from the reference's tests (index-map tables, the
    8x8 MSC matrices, the QuTiP-generated L=4 reduced density matrices).

The GPU box has no /root/reference: tests read only the .npz/.json written here.
"""
import json
import os
import sys
import types
from itertools import combinations
from random import seed, uniform

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- numpy 2 compatibility for the (numpy 1 era) reference ---------------------
_orig_array = np.array


def _array(*a, **k):
    if k.get('copy') is False:
        k['copy'] = None
    return _orig_array(*a, **k)


np.array = _array

# ---- stubs for the compiled / absent modules ------------------------------------
sys.path.insert(0, '/root/reference/src')
_sl = types.ModuleType('slepc4py')
_sl.init = lambda *a, **k: None
sys.modules['slepc4py'] = _sl
_be = types.ModuleType('dynamite._backend')
_be.__path__ = []
_bb = types.ModuleType('dynamite._backend.bbuild')
_bb.dnm_int_t = np.int64
_bb.have_gpu_shell = lambda: False
_bb.complex_enabled = lambda: True
_bs = types.ModuleType('dynamite._backend.bsubspace')
_bs.dnm_int_t = np.int64


class _ST:
    FULL, PARITY, EXPLICIT, SPIN_CONSERVE = 0, 1, 2, 3


_bs.SubspaceType = _ST
for _n in ('Full', 'Parity', 'SpinConserve', 'Explicit'):
    for _f in ('get_dimension_', 'idx_to_state_', 'state_to_idx_', 'C'):
        setattr(_bs, _f + _n, None)
_be.bbuild, _be.bsubspace = _bb, _bs
sys.modules['dynamite._backend'] = _be
sys.modules['dynamite._backend.bbuild'] = _bb
sys.modules['dynamite._backend.bsubspace'] = _bs

import scipy.sparse.linalg  # noqa: E402
from dynamite import config  # noqa: E402
from dynamite import msc_tools  # noqa: E402
from dynamite.extras import majorana  # noqa: E402
from dynamite.operators import (index_sum, op_product, op_sum, sigmax, sigmay,  # noqa: E402
                                sigmaz)
from dynamite.subspaces import Full, XParity  # noqa: E402


def build_hamiltonian(name, L):
    """benchmarking/benchmark.py:129-178, verbatim semantics."""
    config._L = L
    if name == 'MBL':
        rtn = index_sum(op_sum(0.25 * s(0) * s(1) for s in (sigmax, sigmay, sigmaz)))
        seed(0)
        for i in range(L):
            rtn += uniform(-3, 3) * 0.5 * sigmaz(i)
    elif name == 'long_range':
        rtn = op_sum(index_sum(0.25 * sigmaz(0) * sigmaz(i)) for i in range(1, L))
        rtn += 0.5 * index_sum(0.25 * sigmax(0) * sigmax(1))
        rtn += sum(0.05 * index_sum(s()) for s in [sigmax, sigmay, sigmaz])
    elif name == 'SYK':
        seed(0)
        majoranas = [majorana(i) for i in range(L * 2)]

        def gen_products(L):
            for idxs in combinations(range(L * 2), 4):
                p = op_product(majoranas[idx] for idx in idxs)
                p.scale(uniform(-1, 1))
                yield p
        rtn = op_sum(gen_products(L))
        rtn.scale(np.sqrt(6 / (L * 2) ** 3))
    elif name == 'ising':
        rtn = index_sum(0.25 * sigmaz(0) * sigmaz(1)) + 0.1 * index_sum(sigmax())
    elif name == 'XX':
        rtn = index_sum(0.25 * sigmax(0) * sigmax(1))
    elif name == 'heisenberg':
        rtn = index_sum(op_sum(0.25 * s(0) * s(1) for s in (sigmax, sigmay, sigmaz)))
    rtn.L = L
    rtn.reduce_msc()
    return rtn


def popcount(a):
    return np.array([bin(int(v)).count('1') for v in a])


def brute_states(spec):
    L = spec['L']
    allst = np.arange(1 << L, dtype=np.int64)
    t = spec['type']
    if t == 'full':
        return allst
    if t == 'parity':
        return allst[popcount(allst) % 2 == spec['space']]
    if t == 'spinconserve':
        return allst[popcount(allst) == spec['k']]
    if t == 'explicit':
        return np.array(spec['states'], dtype=np.int64)
    raise ValueError(t)


def maps_for(states):
    lookup = {int(s): i for i, s in enumerate(states)}

    def i2s(idx):
        return states[idx]

    def s2i(st):
        return np.array([lookup.get(int(s), -1) for s in np.atleast_1d(st)], dtype=np.int64)
    return i2s, s2i


def dense(msc, left_spec, right_spec, xparity=False):
    ls, rs = brute_states(left_spec), brute_states(right_spec)
    if xparity:
        ls, rs_half = ls[:ls.size // 2], rs[:rs.size // 2]
    i2s, _ = maps_for(ls)
    _, s2i_full = maps_for(rs)
    if xparity:
        half = rs.size // 2

        def s2i(st):
            out = s2i_full(st)
            assert np.all(out < half), 'reduced operator must stay in the representative half'
            return out
        dims = (ls.size, half)
    else:
        s2i = s2i_full
        dims = (ls.size, rs.size)
    return msc_tools.msc_to_numpy(msc, dims, idx_to_state=i2s, state_to_idx=s2i, sparse=False)


def rand_state(n, sd):
    R = np.random.RandomState(sd)
    v = R.standard_normal(n) + 1j * R.standard_normal(n)
    return v / np.linalg.norm(v)


def main():
    out = {}
    meta = {'cases': []}

    def add_case(tag, H, left, right, xparity=False, sector=None, evolve_t=None, nev=0):
        msc = H.msc
        if xparity:
            xp = XParity(Full(L=H.L), sector=sector)
            msc = xp.reduce_msc(msc)
        A = dense(msc, left, right, xparity)
        x = rand_state(A.shape[1], 1234)
        case = {'tag': tag, 'left': left, 'right': right, 'xparity': bool(xparity)}
        out[tag + '.msc_masks'] = msc['masks']
        out[tag + '.msc_signs'] = msc['signs']
        out[tag + '.msc_coeffs'] = msc['coeffs']
        out[tag + '.A'] = A
        out[tag + '.x'] = x
        out[tag + '.y'] = A @ x
        out[tag + '.norm_inf'] = np.array(np.max(np.sum(np.abs(A), axis=1)))
        if A.shape[0] == A.shape[1]:
            out[tag + '.diag'] = np.real(np.diag(A)).copy()
            if evolve_t is not None:
                out[tag + '.evolve_t'] = np.array(evolve_t)
                out[tag + '.evolved'] = scipy.sparse.linalg.expm_multiply(-1j * evolve_t * A, x)
                out[tag + '.evolved_imag'] = scipy.sparse.linalg.expm_multiply(-1.0 * evolve_t * A, x)
            if nev:
                w = np.linalg.eigvalsh(A)
                out[tag + '.evals'] = w
        meta['cases'].append(case)

    full = lambda L: {'type': 'full', 'L': L}  # noqa: E731
    par = lambda L, s: {'type': 'parity', 'L': L, 'space': s}  # noqa: E731
    sc = lambda L, k: {'type': 'spinconserve', 'L': L, 'k': k}  # noqa: E731

    for name, L in [('MBL', 8), ('long_range', 7), ('SYK', 4), ('ising', 6), ('XX', 5), ('heisenberg', 6)]:
        H = build_hamiltonian(name, L)
        add_case(f'{name}_L{L}_full', H, full(L), full(L), evolve_t=1.3, nev=4)

    H = build_hamiltonian('heisenberg', 8)
    add_case('heisenberg_L8_sc4', H, sc(8, 4), sc(8, 4), evolve_t=2.0, nev=4)
    add_case('heisenberg_L8_sc3', H, sc(8, 3), sc(8, 3), evolve_t=2.0, nev=4)
    add_case('heisenberg_L8_par0', H, par(8, 0), par(8, 0), evolve_t=0.7, nev=4)
    add_case('heisenberg_L8_par1', H, par(8, 1), par(8, 1))
    # projections between different subspaces (reference test_multiply.py:108-282)
    add_case('heisenberg_L8_full_to_sc4', H, sc(8, 4), full(8))
    add_case('heisenberg_L8_sc4_to_full', H, full(8), sc(8, 4))
    add_case('heisenberg_L8_par0_to_sc4', H, sc(8, 4), par(8, 0))
    H = build_hamiltonian('SYK', 5)
    add_case('SYK_L5_par0', H, par(5, 0), par(5, 0), evolve_t=0.9, nev=4)
    add_case('SYK_L5_par1', H, par(5, 1), par(5, 1))
    H = build_hamiltonian('long_range', 8)
    add_case('long_range_L8_par1_to_par0', H, par(8, 0), par(8, 1))   # sigma_x/y fields flip parity
    # explicit subspaces: sorted and unsorted (same set as SpinConserve(6,3), shuffled)
    H = build_hamiltonian('heisenberg', 6)
    st = brute_states(sc(6, 3))
    R = np.random.RandomState(7)
    shuffled = st.copy()
    R.shuffle(shuffled)
    add_case('heisenberg_L6_explicit_sorted', H, {'type': 'explicit', 'L': 6, 'states': st.tolist()},
             {'type': 'explicit', 'L': 6, 'states': st.tolist()}, nev=4)
    add_case('heisenberg_L6_explicit_shuffled', H, {'type': 'explicit', 'L': 6, 'states': shuffled.tolist()},
             {'type': 'explicit', 'L': 6, 'states': shuffled.tolist()}, evolve_t=1.1)
    # XParity on top of Full, both sectors
    H = build_hamiltonian('heisenberg', 7)
    add_case('heisenberg_L7_xparity_plus', H, full(7), full(7), xparity=True, sector='+', nev=4)
    add_case('heisenberg_L7_xparity_minus', H, full(7), full(7), xparity=True, sector='-', evolve_t=1.0)
    H = build_hamiltonian('ising', 6)
    add_case('ising_L6_xparity_plus', H, full(6), full(6), xparity=True, sector='+')

    np.savez_compressed(os.path.join(HERE, 'reference_cases.npz'), **out)

    # ---- KATs copied as data from the reference's tests ------------------------------
    kat = {
        # tests/unit/test_subspaces.py:140-190
        'parity_L4': {'0': [0b0000, 0b0011, 0b0101, 0b0110, 0b1001, 0b1010, 0b1100, 0b1111],
                      '1': [0b0001, 0b0010, 0b0100, 0b0111, 0b1000, 0b1011, 0b1101, 0b1110]},
        # :107-118
        'parity_L5_even_single': [[5, 0b01010], [7, 0b01111]],
        # :132-139
        'parity_L5_invalid': [[0, 0b01011], [1, 0b01010]],
        # :226-239
        'spinconserve_dims': [[2, 1, 2], [10, 2, 45], [10, 5, 252]],
        # :263-273
        'spinconserve_L6_k3_single': [5, 0b010101],
        # :275-292
        'spinconserve_invalid': [[5, 1, 0b01011], [5, 2, 0b01011], [5, 3, 0b01010]],
        # :294-342
        'spinconserve_L4': {'1': [0b0001, 0b0010, 0b0100, 0b1000],
                            '2': [0b0011, 0b0101, 0b0110, 0b1001, 0b1010, 0b1100]},
        # tests/unit/test_msc_tools.py:142-172 (msc [(1,5,0.5j),(4,3,-2)], 8x8)
        'msc_full_terms': [[1, 5, [0.0, 0.5]], [4, 3, [-2.0, 0.0]]],
        # tests/integration/test_rdm.py:123-183 (QuTiP)
        'rdm_L4_state': [[0.03, -0.293], [0.131, 0.203], [0.063, 0.17], [0.027, 0.226], [-0.047, 0.292],
                         [0.089, 0.183], [-0.038, -0.024], [0.239, 0.171], [0.283, -0.233], [-0.071, 0.085],
                         [0.178, -0.218], [0.018, 0.271], [0.042, 0.013], [0.26, -0.052], [-0.262, -0.144],
                         [-0.252, 0.167]],
        'rdm_L4_keep0': [[[0.51401, 0.0], [-0.022913, 0.007162]], [[-0.022913, -0.007162], [0.485675, 0.0]]],
        'rdm_L4_keep0_entropy': 0.6916884573920534,
        'rdm_L4_keep2': [[[0.52941, 0.0], [0.011921, 0.059947]], [[0.011921, -0.059947], [0.470275, 0.0]]],
        'rdm_L4_keep2_entropy': 0.6839923299240713,
        'rdm_L4_keep02_entropy': 0.9691946314869655,
        'rdm_L4_keep13_entropy': 0.9691946314869655,
        # :104-121
        'rdm_complex_sign': {'state': [[1, 0], [0, 1], [0, 0], [0, 0]], 'keep': [0],
                             'dm': [[[1, 0], [0, -1]], [[0, 1], [1, 0]]]},
    }
    full_golden = msc_tools.msc_to_numpy([(1, 5, 0.5j), (4, 3, -2)], (8, 8), sparse=False)
    kat['msc_full_dense'] = [[[float(v.real), float(v.imag)] for v in row] for row in full_golden]
    with open(os.path.join(HERE, 'reference_kats.json'), 'w') as f:
        json.dump(kat, f, indent=1)
    with open(os.path.join(HERE, 'reference_cases.json'), 'w') as f:
        json.dump(meta, f, indent=1)
    print('wrote', len(meta['cases']), 'cases')


if __name__ == '__main__':
    main()
