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
L = int(sys.argv[1])
H = build_hamiltonian(os.environ.get('DNM_MODEL', 'MBL'), L); H.reduce_msc()
sub = Full(L=L)
masks, offs = msc_tools.mask_offsets(H.msc)
n = 1 << L
x, y = Vec(n), Vec(n)
x.setRandom(0)
for fuse, lag, tb, *rest in [tuple(int(v) for v in a.split(',')) for a in sys.argv[2:]]:
    os.environ['DNM_FUSE_BITS'] = str(fuse); os.environ['DNM_FUSE_LAG'] = str(lag)
    mat = bpetsc.build_mat(masks, offs, np.ascontiguousarray(H.msc['signs']), np.ascontiguousarray(H.msc['coeffs']), sub._to_c(), sub._to_c(), False, True, True)
    bpetsc.precompute_diagonal(mat)
    mat.set_option('tile_bits', tb)
    if rest: mat.set_option('tile_rows', rest[0])
    if len(rest) > 1: mat.set_option('pipeline', rest[1])
    for _ in range(2): mat.mult(x, y)
    lib.dnm_synchronize(); lib.dnm_timer_start()
    for _ in range(5): mat.mult(x, y)
    ms = C.c_float(); lib.dnm_timer_stop(C.byref(ms))
    print(f'L={L} T={tb} R={rest[0] if rest else 0} pipe={rest[1] if len(rest) > 1 else 0} fuse={fuse} lag={lag} launches={mat.get_info("launches_per_mult"):.0f} {ms.value/5:.3f} ms', flush=True)
    mat.destroy()
