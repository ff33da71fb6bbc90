"""Timing breakdown of evolve / eigsolve on the BASELINE configs (exploration, not the bench contract)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from dynamite_b200 import _capi, slepc
from dynamite_b200.hamiltonians import build_hamiltonian
from dynamite_b200.states import State
from dynamite_b200.subspaces import Full, Parity, SpinConserve

_capi.ensure_gpu(0)
lib = _capi.lib()


def sync():
    lib.dnm_synchronize()


def timeit(fn, reps=1):
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    sync()
    return (time.perf_counter() - t0) / reps, r


def report(name, L, sub, evolve_t=None, nev=None, ncv=None):
    t0 = time.perf_counter()
    H = build_hamiltonian(name, L)
    H.subspace = sub
    tb = time.perf_counter() - t0
    tm, mat = timeit(H.get_mat)
    n = sub.get_dimension()
    x = State(subspace=sub)
    x.vec.setRandom(1)
    x.vec.normalize()
    x.set_initialized()
    y = State(subspace=sub)
    H.dot(x, y)
    tmm, _ = timeit(lambda: H.dot(x, y), 20)
    print(f'{name} L={L} dim={n} host-build {tb:.2f}s mat-build {tm:.3f}s matmult {tmm*1e3:.3f} ms kernel={mat.get_info("kernel"):.0f} passes={mat.get_info("passes"):.0f}', flush=True)
    if evolve_t is not None:
        nrm = H.infinity_norm()
        t = evolve_t if evolve_t > 0 else -evolve_t / nrm
        mfn = slepc.MFN().create()
        mfn.getFN().setScale(-1j * t)
        if ncv:
            mfn.setDimensions(ncv)
        mfn.setOperator(mat)
        lib.dnm_launch_count(1)
        te, _ = timeit(lambda: mfn.solve(x.vec, y.vec))
        print(f'   evolve t={t:.4g} (|H|={nrm:.3f}): {te:.3f}s its={mfn.its} matmults={mfn.matmults} '
              f'-> {te/mfn.matmults*1e3:.3f} ms per matmult-equivalent, launches={lib.dnm_launch_count(0)}', flush=True)
    if nev is not None:
        eps = slepc.EPS().create()
        eps.setOperators(mat)
        eps.setDimensions(nev)
        lib.dnm_launch_count(1)
        te, _ = timeit(eps.solve)
        print(f'   eigsolve nev={nev}: {te:.3f}s its={eps.its} matmults={eps.matmults} nconv={eps.nconv} '
              f'-> {te/max(eps.matmults,1)*1e3:.3f} ms per matmult-equivalent, launches={lib.dnm_launch_count(0)} '
              f'evals={eps._evals[:nev]}', flush=True)
        eps.destroy()
    H.destroy_mat()


which = sys.argv[1:] or ['C1', 'C2']
if 'C1' in which:
    report('heisenberg', 20, Full(L=20), evolve_t=1.0)
if 'C2' in which:
    report('heisenberg', 26, SpinConserve(26, 13), nev=4)
if 'C3' in which:
    report('MBL', 30, Full(L=30), evolve_t=-50.0)
if 'C3s' in which:
    report('MBL', 26, Full(L=26), evolve_t=-50.0)
if 'C4' in which:
    report('SYK', 20, Parity('even', L=20), evolve_t=-50.0)
if 'C4s' in which:
    report('SYK', 12, Parity('even', L=12), evolve_t=-50.0)
