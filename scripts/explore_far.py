"""Time the tiled MatMult for (tile_bits, far_bits[, jit]) combinations on one GPU and check every
variant against the first one (max |dy| / |y|).   python scripts/explore_far.py MBL 30 13,0 11,8 12,7"""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from dynamite_b200 import _capi, msc_tools
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.petsc import Vec
from dynamite_b200.subspaces import Full
from dynamite_b200._backend import bpetsc
_capi.ensure_gpu(0)
lib = _capi.lib()
model, L = sys.argv[1], int(sys.argv[2])
H = build_hamiltonian(model, L); H.reduce_msc()
sub = Full(L=L)
masks, offs = msc_tools.mask_offsets(H.msc)
n = 1 << L
x, y, y0 = Vec(n), Vec(n), Vec(n)
x.setRandom(0)
first = True
reps = int(os.environ.get('REPS', '5'))
import time
for tb, far, *rest in [tuple(int(v) for v in a.split(',')) for a in sys.argv[3:]]:
    mat = bpetsc.build_mat(masks, offs, np.ascontiguousarray(H.msc['signs']), np.ascontiguousarray(H.msc['coeffs']), sub._to_c(), sub._to_c(), False, True, True)
    bpetsc.precompute_diagonal(mat)
    mat.set_option('tile_bits', tb)
    mat.set_option('far_bits', far)
    mat.set_option('verbose', int(os.environ.get('VERBOSE', '0')))
    mat.set_option('jit', rest[0] if rest else 0)
    mat.set_option('pipeline', rest[1] if len(rest) > 1 else 0)
    t0 = time.perf_counter()
    mat.mult(x, y); lib.dnm_synchronize()
    t_first = time.perf_counter() - t0
    for _ in range(2): mat.mult(x, y)
    lib.dnm_synchronize(); lib.dnm_timer_start()
    for _ in range(reps): mat.mult(x, y)
    ms = C.c_float(); lib.dnm_timer_stop(C.byref(ms))
    if first:
        mat.mult(x, y0); err = 0.0; first = False
    else:
        y.axpy(-1.0, y0); err = y.norm() / y0.norm()
    print(f'{model} L={L} T={tb} far={far} pipe={rest[1] if len(rest) > 1 else 0} jit={mat.get_info("jit_passes"):.0f} passes={mat.get_info("passes"):.0f} first={t_first:.2f}s {ms.value/reps:.3f} ms  diff_vs_first={err:.2e}', flush=True)
    mat.destroy()
