"""Explicit/Auto-subspace MatMult: the shared-memory staged search (k_mult_explicit) against the global
binary search of the general kernel and against the SpinConserve kernel on the same space.
   python scripts/explore_explicit.py 26"""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from dynamite_b200 import _capi, msc_tools
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.petsc import Vec
from dynamite_b200.subspaces import SpinConserve, Explicit
from dynamite_b200._backend import bpetsc
_capi.ensure_gpu(0)
lib = _capi.lib()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 26
H = build_hamiltonian('heisenberg', L); H.reduce_msc()
masks, offs = msc_tools.mask_offsets(H.msc)
sc = SpinConserve(L, L // 2)
n = sc.get_dimension()
states = sc.idx_to_state(np.arange(n, dtype=np.int64))
ex = Explicit(states, L=L)
x = Vec(n); x.setRandom(0)
ys = []
def run(sub, label, env=None):
    if env: os.environ[env] = '1'
    mat = bpetsc.build_mat(masks, offs, np.ascontiguousarray(H.msc['signs']), np.ascontiguousarray(H.msc['coeffs']), sub._to_c(), sub._to_c(), False, True, True)
    bpetsc.precompute_diagonal(mat)
    y = Vec(n)
    for _ in range(2): mat.mult(x, y)
    lib.dnm_synchronize(); lib.dnm_timer_start()
    for _ in range(10): mat.mult(x, y)
    ms = C.c_float(); lib.dnm_timer_stop(C.byref(ms))
    if env: os.environ.pop(env)
    ys.append(y)
    err = 0.0
    if len(ys) > 1:
        d = y.copy() if hasattr(y, 'copy') else y
        d.axpy(-1.0, ys[0]); err = d.norm() / ys[0].norm()
    print(f'L={L} dim={n} {label}: {ms.value/10:.3f} ms per MatMult  diff_vs_first={err:.2e}', flush=True)
    mat.destroy()
run(sc, 'SpinConserve kernel (rank arithmetic)')
run(ex, 'Explicit, shared-memory staged search (k_mult_explicit)')
run(ex, 'Explicit, global binary search (k_mult_general)', 'DNM_NO_EXPLICIT_KERNEL')
