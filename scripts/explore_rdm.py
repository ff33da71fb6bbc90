import sys, time
import numpy as np
sys.path.insert(0, '.')
from dynamite_b200 import _capi
from dynamite_b200.computations import reduced_density_matrix
from dynamite_b200.states import State
from dynamite_b200.subspaces import Full, SpinConserve
_capi.ensure_gpu(0)
lib = _capi.lib()
for L, k, sub in [(20, 10, None), (24, 12, None), (26, 13, None), (24, 12, 'sc'), (28, 8, None)]:
    s = State(subspace=Full(L=L) if sub is None else SpinConserve(L, L // 2))
    s.vec.setRandom(3); s.vec.normalize(); s.set_initialized()
    keep = list(range(k))
    reduced_density_matrix(s, keep[:2])
    lib.dnm_synchronize(); t0 = time.perf_counter()
    lib.dnm_launch_count(1); lib.dnm_timer_start()
    rho = reduced_density_matrix(s, keep)
    import ctypes as C
    ms = C.c_float(); lib.dnm_timer_stop(C.byref(ms))
    dt = time.perf_counter() - t0
    print(f'   device time {ms.value/1e3:.3f} s', end='')
    macs = 2.0 ** (L + k)
    print(f'L={L} keep={k} {sub or "full"}: {dt:.3f} s  trace={np.trace(rho).real:.12f}  {macs*8/dt/1e12:.2f} TFLOP/s (fp64, dense-equivalent)', flush=True)
