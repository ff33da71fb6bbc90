"""C2 (L=26 Heisenberg, SpinConserve(26,13), lowest 4 eigenpairs): wall time of repeated eigsolve calls
with the phase trace (DNM_TRACE=1)."""
import sys, time
sys.path.insert(0, '.')
from dynamite_b200 import _capi
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.subspaces import SpinConserve
_capi.ensure_gpu(0)
lib = _capi.lib()
H = build_hamiltonian('heisenberg', 26); H.subspace = SpinConserve(26, 13)
H.get_mat()
for k in range(3):
    lib.dnm_synchronize(); t0 = time.perf_counter(); ev = H.eigsolve(nev=4); lib.dnm_synchronize()
    print('eigsolve', k, round(time.perf_counter() - t0, 4), ev[:2], flush=True)
