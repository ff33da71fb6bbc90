"""Operator helpers used by the benchmark Hamiltonians (reference ``extras.py``)."""
from .operators import index_product, sigmax, sigmay, sigmaz


def commutator(op1, op2):
    return op1 * op2 - op2 * op1


def majorana(idx):
    """Majorana ``idx`` under Jordan-Wigner: a sigma_z string on sites below
    ``idx//2``, then sigma_x (even idx) or sigma_y (odd idx) on site ``idx//2``."""
    site = idx // 2
    rtn = sigmay(site) if idx % 2 else sigmax(site)
    if site > 0:
        rtn = index_product(sigmaz(), size=site) * rtn
    return rtn
