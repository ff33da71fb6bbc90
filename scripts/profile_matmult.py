"""A handful of MatMults for ncu: python scripts/profile_matmult.py MBL 26 [tile_bits] [reps]"""
import sys

import numpy as np

sys.path.insert(0, '.')
from dynamite_b200 import _capi
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.states import State
from dynamite_b200.subspaces import Full, Parity, SpinConserve

name, L = sys.argv[1], int(sys.argv[2])
tile_bits = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
subname = sys.argv[5] if len(sys.argv) > 5 else 'full'
_capi.ensure_gpu(0)
H = build_hamiltonian(name, L)
sub = {'full': lambda: Full(L=L), 'parity': lambda: Parity('even', L=L), 'spinconserve': lambda: SpinConserve(L, L // 2)}[subname]()
H.subspace = sub
x = State(L=L, subspace=sub)
x.set_uniform()
y = State(L=L, subspace=sub)
mat = H.get_mat()
if tile_bits:
    mat.set_option('tile_bits', tile_bits)
for _ in range(reps):
    H.dot(x, y)
_capi.lib().dnm_synchronize()
print('done', mat.get_info('kernel'), mat.get_info('passes'))
