"""C1 (L=20 Heisenberg evolve t=1): wall time of repeated calls, with the phase trace (DNM_TRACE=1)."""
import sys, time
sys.path.insert(0, '.')
from dynamite_b200 import _capi
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.states import State
from dynamite_b200.subspaces import Full
_capi.ensure_gpu(0)
lib = _capi.lib()
H = build_hamiltonian('heisenberg', 20); H.subspace = Full(L=20)
s = State(L=20, subspace=H.subspace); s.vec.setRandom(1); s.vec.normalize(); s.set_initialized()
r = State(L=20, subspace=H.subspace)
H.get_mat()
for k in range(6):
    lib.dnm_synchronize(); t0 = time.perf_counter(); H.evolve(s, 1.0, result=r); lib.dnm_synchronize()
    print('evolve', k, round(time.perf_counter() - t0, 4), flush=True)
