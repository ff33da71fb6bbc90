"""
The benchmark Hamiltonians of the reference (``benchmarking/benchmark.py:129-178``):
same operators, same ``random.seed(0)`` streams, so sizes and term counts match
BASELINE.md section 2.
"""
from itertools import combinations
from random import seed, uniform

import numpy as np

from .extras import majorana
from .operators import index_sum, op_product, op_sum, sigmax, sigmay, sigmaz

NAMES = ('MBL', 'long_range', 'SYK', 'ising', 'XX', 'heisenberg')


def build_hamiltonian(name, L):
    """``name`` in :data:`NAMES`; ``L`` spins (SYK: ``2L`` Majoranas)."""
    def s0(op, i=0):
        return op(i)

    def isum(op):
        return index_sum(op, size=L)

    if name == 'MBL':
        H = isum(op_sum(0.25 * s0(s) * s0(s, 1) for s in (sigmax, sigmay, sigmaz)))
        seed(0)
        for i in range(L):
            H += uniform(-3, 3) * 0.5 * sigmaz(i)
    elif name == 'long_range':
        H = op_sum(isum(0.25 * s0(sigmaz) * s0(sigmaz, i)) for i in range(1, L))
        H += 0.5 * isum(0.25 * s0(sigmax) * s0(sigmax, 1))
        H += op_sum(0.05 * isum(s0(s)) for s in (sigmax, sigmay, sigmaz))
    elif name == 'SYK':
        seed(0)
        majoranas = [majorana(i) for i in range(2 * L)]

        def products():
            for idxs in combinations(range(2 * L), 4):
                p = op_product(majoranas[i] for i in idxs)
                p.scale(uniform(-1, 1))
                yield p

        H = op_sum(products())
        H.scale(np.sqrt(6 / (2 * L) ** 3))
    elif name == 'ising':
        H = isum(0.25 * s0(sigmaz) * s0(sigmaz, 1)) + 0.1 * isum(s0(sigmax))
    elif name == 'XX':
        H = isum(0.25 * s0(sigmax) * s0(sigmax, 1))
    elif name == 'heisenberg':
        H = isum(op_sum(0.25 * s0(s) * s0(s, 1) for s in (sigmax, sigmay, sigmaz)))
    else:
        raise ValueError('Unrecognized Hamiltonian.')
    H.L = L
    H.allow_projection = True   # as benchmark.py:174-176
    return H
