"""Exploratory timing of the MatMult kernels (not the bench contract)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from dynamite_b200 import _capi, msc_tools
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.petsc import Vec
from dynamite_b200.subspaces import Full, Parity, SpinConserve
from dynamite_b200._backend import bpetsc

_capi.ensure_gpu(0)
lib = _capi.lib()


def time_mult(mat, n, reps=5):
    x, y = Vec(n), Vec(n)
    x.set(1.0 / np.sqrt(n))
    for _ in range(2):
        mat.mult(x, y)
    lib.dnm_synchronize()
    lib.dnm_timer_start()
    for _ in range(reps):
        mat.mult(x, y)
    ms = C.c_float()
    lib.dnm_timer_stop(C.byref(ms))
    x.destroy(); y.destroy()
    return ms.value / reps


def build(H, sub, diag):
    H.reduce_msc()
    masks, offs = msc_tools.mask_offsets(H.msc)
    mat = bpetsc.build_mat(masks, offs, np.ascontiguousarray(H.msc['signs']), np.ascontiguousarray(H.msc['coeffs']),
                           sub._to_c(), sub._to_c(), False, True, True)
    if diag:
        bpetsc.precompute_diagonal(mat)
    return mat


import os
for name, L in [(a.split(':')[0], int(a.split(':')[1])) for a in sys.argv[1:]]:
    H = build_hamiltonian(name, L)
    sub = Full(L=L)
    n = 1 << L
    for diag in (True,):
        mat = build(H, sub, diag)
        model = mat.get_info('model_bytes')
        for kern, tb, rb, rows, fuse, lag in [(2, 11, 3, 8, 0, 2), (2, 11, 3, 8, 18, 2), (2, 11, 3, 8, 18, 1), (2, 11, 3, 8, 18, 4),
                                              (2, 11, 3, 8, 20, 2), (2, 11, 3, 8, 16, 2), (2, 12, 3, 8, 18, 2), (2, 12, 3, 8, 20, 2),
                                              (2, 12, 2, 8, 20, 2), (2, 11, 2, 8, 18, 2), (2, 11, 3, 16, 18, 2), (2, 0, -1, 0, 18, 2)]:
            if rb >= 0:
                os.environ['DNM_TILE_RUN_BITS'] = str(rb)
            else:
                os.environ.pop('DNM_TILE_RUN_BITS', None)
            os.environ['DNM_FUSE_BITS'] = str(fuse)
            os.environ['DNM_FUSE_LAG'] = str(lag)
            mat.set_option('kernel', kern)
            if kern == 2:
                mat.set_option('tile_bits', tb)
                mat.set_option('tile_rows', rows)
            ms = time_mult(mat, n)
            print(f'{name} L={L} diag={diag} T={tb} B={rb} R={rows} fuse={fuse} lag={lag} launches={mat.get_info("launches_per_mult"):.0f} passes={mat.get_info("passes"):.0f} '
                  f'{ms:.3f} ms  model {model/ms/1e6:.0f} GB/s  compulsory {mat.get_info("compulsory_bytes")/ms/1e6:.0f} GB/s',
                  flush=True)
        mat.destroy()
