import os, sys, ctypes as C, time
import numpy as np
sys.path.insert(0, '.')
from dynamite_b200 import _capi
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.states import State
from dynamite_b200.subspaces import Parity
_capi.ensure_gpu(0)
lib = _capi.lib()
L = int(sys.argv[1]) if len(sys.argv) > 1 else 20
H = build_hamiltonian('SYK', L)
sub = Parity('even', L=L)
H.subspace = sub
x = State(subspace=sub); x.vec.setRandom(1); x.set_initialized()
y = State(subspace=sub)
mat = H.get_mat()
ref = None
for kern, tb, rows in [(2, 12, 8), (1, 0, 0), (2, 11, 8), (2, 13, 8), (2, 12, 16), (2, 10, 8)]:
    mat.set_option('kernel', kern)
    if kern == 2:
        mat.set_option('tile_bits', tb); mat.set_option('tile_rows', rows)
    H.dot(x, y); H.dot(x, y)
    lib.dnm_synchronize(); lib.dnm_timer_start()
    for _ in range(5): H.dot(x, y)
    ms = C.c_float(); lib.dnm_timer_stop(C.byref(ms))
    out = y.to_numpy()
    if ref is None: ref = out
    print(f'SYK L={L} kernel={kern} T={tb} R={rows} passes={mat.get_info("passes"):.0f} {ms.value/5:.3f} ms  diff vs first {np.linalg.norm(out-ref)/np.linalg.norm(ref):.1e}', flush=True)
