import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, '.')
from dynamite_b200 import _capi
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.states import State
from dynamite_b200.subspaces import Parity
_capi.ensure_gpu(0)
lib = _capi.lib()
L = 20
H = build_hamiltonian('SYK', L)
sub = Parity('even', L=L)
H.subspace = sub
x = State(subspace=sub); x.vec.setRandom(1); x.set_initialized()
y = State(subspace=sub)
mat = H.get_mat()
for rep in range(3):
    for stage in (1, 0):
        if stage: os.environ.pop('DNM_NO_STAGE', None)
        else: os.environ['DNM_NO_STAGE'] = '1'
        mat.set_option('tile_bits', 12); mat.set_option('tile_rows', 8)   # invalidates the plan
        H.dot(x, y); H.dot(x, y)
        lib.dnm_synchronize(); lib.dnm_timer_start()
        for _ in range(5): H.dot(x, y)
        ms = C.c_float(); lib.dnm_timer_stop(C.byref(ms))
        print(f'rep {rep} staged={stage} {ms.value/5:.3f} ms', flush=True)
