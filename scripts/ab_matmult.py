"""A/B timing of one MatMult under different environment knobs, one fresh process per variant
(the knobs are read when the plan is built and some launch attributes are set once per process).

usage: python scripts/ab_matmult.py MODEL L [SUBSPACE] -- "" "DNM_ROWOFF_ARITH=1" "DNM_NO_PAIR=1 DNM_TILE_RUN_BITS=2" ...

Each variant is a space-separated list of NAME=VALUE; the empty string is the default build.  The
special names tile_bits / tile_rows / pipeline / kernel are passed to dnm_mat_set_option instead.
Prints one line per variant: milliseconds per MatMult (CUDA events, 2 warm-ups, 5 timed), passes.
"""
import os
import subprocess
import sys

OPTIONS = ('tile_bits', 'tile_rows', 'pipeline', 'kernel')


def child(model, L, subname, opts):
    import ctypes as C
    import numpy as np
    sys.path.insert(0, '.')
    from dynamite_b200 import _capi, msc_tools
    from dynamite_b200._backend import bpetsc
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.petsc import Vec
    from dynamite_b200.subspaces import Full, Parity
    _capi.ensure_gpu(0)
    lib = _capi.lib()
    H = build_hamiltonian(model, L)
    H.reduce_msc()
    sub = Parity('even', L=L) if subname == 'parity' else Full(L=L)
    masks, offs = msc_tools.mask_offsets(H.msc)
    n = sub.get_dimension()
    x, y = Vec(n), Vec(n)
    x.setRandom(0)
    mat = bpetsc.build_mat(masks, offs, np.ascontiguousarray(H.msc['signs']), np.ascontiguousarray(H.msc['coeffs']),
                           sub._to_c(), sub._to_c(), False, True, True)
    bpetsc.precompute_diagonal(mat)
    for k, v in opts.items():
        mat.set_option(k, int(v))
    for _ in range(2):
        mat.mult(x, y)
    lib.dnm_synchronize()
    lib.dnm_timer_start()
    for _ in range(5):
        mat.mult(x, y)
    ms = C.c_float()
    lib.dnm_timer_stop(C.byref(ms))
    print(f'{ms.value / 5:9.3f} ms  kernel={mat.get_info("kernel"):.0f} passes={mat.get_info("passes"):.0f} '
          f'launches={mat.get_info("launches_per_mult"):.0f}', flush=True)


def main():
    if sys.argv[1] == '--child':
        model, L, subname = sys.argv[2], int(sys.argv[3]), sys.argv[4]
        opts = dict(a.split('=') for a in sys.argv[5:])
        return child(model, L, subname, opts)
    split = sys.argv.index('--')
    head, variants = sys.argv[1:split], sys.argv[split + 1:]
    model, L = head[0], head[1]
    subname = head[2] if len(head) > 2 else 'full'
    for v in variants or ['']:
        env = dict(os.environ)
        opts = []
        for item in v.split():
            k, val = item.split('=', 1)
            if k in OPTIONS:
                opts.append(item)
            else:
                env[k] = val
        res = subprocess.run([sys.executable, __file__, '--child', model, L, subname] + opts, env=env,
                             capture_output=True, text=True)
        out = res.stdout.strip().splitlines()
        print(f'{v or "(default)":48s} {out[-1] if out else "FAILED: " + res.stderr.strip().splitlines()[-1]}', flush=True)


if __name__ == '__main__':
    main()
