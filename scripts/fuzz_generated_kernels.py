"""Fuzz the planner + kernel generator on the CPU (no GPU needed).

  python scripts/fuzz_generated_kernels.py emulate COUNT [FIRST_SEED]
      random operators (Pauli strings, two-pattern masks, the benchmark models), random L = 13..17, tile /
      far / pipeline / autotuner shape, Full or Parity, 1..8 ranks: the generated source must be accepted by
      NVRTC (CUDA mode) and, compiled for the host (tests/test_jit_emulation.py), reproduce the oracle's product
  python scripts/fuzz_generated_kernels.py compile COUNT [FIRST_SEED]
      the same families at L = 24..33 on 1..8 ranks: NVRTC must accept what the generator emits (a rejected
      source silently costs a plan its generated kernels)

The first run of this (round 2) found the contiguous accumulating pass bug fixed in csrc/jit.cu
(tests/test_jit_emulation.py::test_contiguous_accumulating_pass)."""
import os
import pathlib
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from test_jit_emulation import NotGenerated, check, random_pair_operator, random_pauli_operator  # noqa: E402
from test_jit_generator import dryrun  # noqa: E402

from dynamite_b200.hamiltonians import build_hamiltonian  # noqa: E402
from dynamite_b200.subspaces import Parity  # noqa: E402

MODELS = ['MBL', 'heisenberg', 'long_range', 'ising', 'XX']


def draw(seed, big):
    rng = np.random.default_rng(seed)
    if big:
        p = int(rng.choice([0, 0, 1, 2, 3]))
        L = int(rng.choice([24, 27, 30])) + p
        nranks = 1 << p
    else:
        L = int(rng.choice([13, 14, 15, 16, 17]))
        nranks = int(rng.choice([1, 1, 2, 4, 8]))
        if L - (nranks.bit_length() - 1) < 12:
            nranks = 1
    fam = int(rng.integers(0, 3))
    ns, ng = int(rng.integers(4, 40 if big else 30)), int(rng.integers(3, 13))
    model = str(rng.choice(MODELS))

    def make():
        if fam == 0:
            return random_pauli_operator(L, seed, nstrings=ns)
        if fam == 1:
            return random_pair_operator(L, seed, ngroups=ng)
        return build_hamiltonian(model, L)
    if big:
        kw = dict(tile_bits=int(rng.choice([0, 9, 10, 11, 12, 13])), far_bits=int(rng.integers(-1, 11)),
                  pipeline=int(rng.integers(0, 4) == 0))
    else:
        kw = dict(tile_bits=int(rng.choice([9, 10, 11, 12])), far_bits=int(rng.integers(0, 6)),
                  pipeline=int(rng.integers(0, 4) == 0))
    if rng.integers(0, 3) == 0:
        kw = dict(tune=int(rng.integers(0, 4)))
    sub = None
    if fam == 2 and model in ('heisenberg', 'MBL', 'XX') and rng.integers(0, 3) == 0:
        sub = Parity('even' if rng.integers(0, 2) == 0 else 'odd', L=L)
    desc = dict(seed=seed, L=L, family=('pauli', 'pairs', model)[fam], nranks=nranks, sub='parity' if sub else 'full', **kw)
    return make, L, nranks, sub, kw, desc, rng


def main():
    mode, count = sys.argv[1], int(sys.argv[2])
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    tmp = pathlib.Path(tempfile.mkdtemp(prefix='dnm_fuzz_'))
    ok = skipped = bad = 0
    t0 = time.time()
    for seed in range(first, first + count):
        make, L, nranks, sub, kw, desc, rng = draw(seed, mode == 'compile')
        try:
            for rank in sorted({0, nranks - 1, int(rng.integers(0, nranks))}):
                r = dryrun(make(), L, sub=sub, nranks=nranks, rank=rank, **kw)
                assert r['kernels'] == 0 or r['cubin'] > 0
            if mode == 'emulate':
                check(tmp, make(), L, sub=sub, nranks=nranks, seed=seed, **kw)
            ok += 1
        except NotGenerated:
            skipped += 1        # a pass kept the table-driven kernel: nothing to emulate
        except Exception as exc:  # noqa: BLE001
            bad += 1
            print('FAIL', desc, repr(exc)[:500], flush=True)
    print(f'{mode}: ok {ok}, not fully generated {skipped}, FAILED {bad}, {time.time() - t0:.0f} s')
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
