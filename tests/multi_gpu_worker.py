"""Sharded-state checks, one process per GPU.  Launched by test_gpu_multi.py (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_worker.py

Every rank builds the same small problem, holds only its shard on its GPU, and compares its
slice with the single-process oracle / scipy result.  Exit code 0 = all checks passed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import scipy.sparse.linalg
    import torch.distributed as dist

    import oracle
    from dynamite_b200 import _capi, msc_tools
    from dynamite_b200.computations import reduced_density_matrix
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.sharding import init_comm
    from dynamite_b200.states import State
    from dynamite_b200.subspaces import Full, Parity
    from helpers import rand_state, rel_err

    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist.init_process_group('gloo')
    _capi.ensure_gpu(local_rank)
    rank, world = init_comm(dist)
    failures = []

    def check(name, ok, detail=''):
        if not ok:
            failures.append(f'[rank {rank}] {name} {detail}')

    def sharded_state(sub, full_vec):
        s = State(subspace=sub)
        a, b = s.vec.getOwnershipRange()
        s.vec[a:b] = full_vec[a:b]
        s.set_initialized()
        return s, a, b

    cases = [('MBL', 16, 'full'), ('long_range', 15, 'full'), ('heisenberg', 16, 'parity'), ('SYK', 9, 'parity')]
    for name, L, kind in cases:
        H = build_hamiltonian(name, L)
        sub = Full(L=L) if kind == 'full' else Parity('even', L=L)
        H.subspace = sub
        spec = {'type': kind, 'L': L, 'space': 0}
        osub = oracle.Subspace(spec)
        H.reduce_msc()
        masks, offs = msc_tools.mask_offsets(H.msc)
        omsc = oracle.Msc(masks, offs, H.msc['signs'], H.msc['coeffs'])
        xfull = rand_state(osub.dim, 7)
        want = oracle.matmult(omsc, osub, osub, xfull)
        for diag in (True, False):
            H.precompute_diagonal = diag
            x, a, b = sharded_state(sub, xfull)
            y = H.dot(x)
            check(f'matmult {name} L={L} {kind} diag={diag}', rel_err(y.vec[a:b], want[a:b]) < 1e-12,
                  f'err={rel_err(y.vec[a:b], want[a:b]):.2e}')
            # the other two ways of fetching remote amplitudes: NVLink peer loads inside the MatMult
            # kernel, and the same on a side stream into a separate buffer
            for mode in ('peer', 'peer_overlap', 'dma'):
                os.environ['DNM_REMOTE'] = mode
                H.get_mat().set_option('tile_bits', 0)      # drops the cached plan
                y2 = H.dot(x)
                check(f'matmult[{mode}] {name} L={L}', rel_err(y2.vec[a:b], want[a:b]) < 1e-12)
            # remote masks folded into the local passes of generated code (FAR groups reading the partner's
            # shard over NVLink inside the pass): forced on at this small size, several tile shapes
            os.environ['DNM_REMOTE'] = 'fold'
            mat = H.get_mat()
            mat.set_option('jit', 1)
            for tile_bits, far in ((0, -1), (9, 3), (10, 0), (11, 2)):
                mat.set_option('tile_bits', tile_bits)
                mat.set_option('far_bits', far)
                y5 = H.dot(x)
                check(f'matmult[fold T={tile_bits} far={far}] {name} L={L} diag={diag}',
                      rel_err(y5.vec[a:b], want[a:b]) < 1e-12, f'err={rel_err(y5.vec[a:b], want[a:b]):.2e}')
                if name != 'SYK' and (diag or name == 'XX'):
                    check(f'generated kernels ran [{name}]', mat.get_info('jit_passes') >= 1)
            mat.set_option('far_bits', -1)
            os.environ.pop('DNM_REMOTE', None)
            # the first-use autotuner: every rank runs the same trials, the slowest rank decides
            mat.set_option('tile_bits', 0)
            mat.set_option('autotune', 1)
            y6 = H.dot(x)
            check(f'matmult[autotuned] {name} L={L} diag={diag}', rel_err(y6.vec[a:b], want[a:b]) < 1e-12)
            mat.set_option('autotune', -1)
            mat.set_option('jit', -1)
            mat.set_option('tile_bits', 0)
            # DMA staging copies only the part of a partner shard this rank can touch (XX+YY terms:
            # half of it or nothing); the rest of the staging buffer is poisoned with NaNs here
            os.environ['DNM_POISON_STAGE'] = '1'
            y3 = H.dot(x)
            check(f'matmult[dma, poisoned staging] {name} L={L}', rel_err(y3.vec[a:b], want[a:b]) < 1e-12)
            os.environ['DNM_FULL_STAGE'] = '1'
            H.get_mat().set_option('tile_bits', 0)
            y4 = H.dot(x)
            check(f'matmult[dma, full staging] {name} L={L}', rel_err(y4.vec[a:b], want[a:b]) < 1e-12)
            os.environ.pop('DNM_POISON_STAGE', None)
            os.environ.pop('DNM_FULL_STAGE', None)
            os.environ.pop('DNM_REMOTE', None)
            H.get_mat().set_option('tile_bits', 0)
            nrm = H.infinity_norm()
            check(f'norm {name}', abs(nrm - oracle.norm_inf(omsc, osub, osub)) < 1e-12 * nrm)
            # whole vector on every rank / on rank 0 only (read through the peer mappings)
            got_all = y.to_numpy(to_all=True)
            check('to_numpy(to_all)', rel_err(got_all, want) < 1e-12)
            got0 = y.to_numpy()
            check('to_numpy rank 0 only', (got0 is None) == (rank != 0) and (rank != 0 or rel_err(got0, want) < 1e-12))
            # global reductions
            check('dot', abs(x.dot(y) - np.vdot(xfull, want)) < 1e-12)
            check('norm2', abs(y.norm() - np.linalg.norm(want)) < 1e-12)
        # Krylov consumers on the sharded state
        if L <= 15 or name == 'heisenberg':
            A = msc_tools.msc_to_numpy(H.msc, (osub.dim, osub.dim), osub.i2s, osub.s2i)
            x, a, b = sharded_state(sub, xfull)
            t = 3.0 / H.infinity_norm()
            ev = H.evolve(x, t, tol=1e-12)
            ref = scipy.sparse.linalg.expm_multiply(-1j * t * A, xfull)
            check(f'evolve {name}', rel_err(ev.vec[a:b], ref[a:b]) < 1e-10, f'err={rel_err(ev.vec[a:b], ref[a:b]):.2e}')
            ec = H.evolve(x, t, algo='chebyshev')
            check(f'evolve chebyshev {name}', rel_err(ec.vec[a:b], ref[a:b]) < 1e-10, f'err={rel_err(ec.vec[a:b], ref[a:b]):.2e}')
            evals, evecs = H.eigsolve(nev=3, getvecs=True, tol=1e-11)
            w = scipy.sparse.linalg.eigsh(A, k=3, which='SA')[0]
            check(f'eigsolve {name}', abs(evals[0] - w.min()) < 1e-9, f'{evals[:3]} vs {np.sort(w)}')
            v0 = evecs[0]
            Hv = H.dot(v0)
            Hv.axpy(-evals[0], v0)
            check(f'eigvec residual {name}', Hv.norm() < 1e-7, f'{Hv.norm():.2e}')
        # checkpoint: rank 0 writes the header and the metadata, every rank its own block of the
        # PETSc binary Vec; reading back gives every rank its block again
        if name in ('MBL', 'heisenberg'):
            import tempfile
            x, a, b = sharded_state(sub, xfull)
            path = os.path.join(tempfile.gettempdir(), f'dnm_ckpt_{name}_{world}')
            x.save(path)
            back = State.from_file(path)
            check(f'save/from_file {name}', np.array_equal(back.vec[a:b], xfull[a:b]) and back.subspace == sub)
            if rank == 0:
                raw = np.fromfile(path + '.vec', dtype='>c16', offset=8)
                check('file holds the whole vector', np.array_equal(raw, xfull))
            dist.barrier()
            if rank == 0:
                os.remove(path + '.vec')
                os.remove(path + '.metadata')
        if kind == 'full':
            x, a, b = sharded_state(sub, xfull)
            for keep in ([0, 1, 2], [L - 1], [1, L - 2, L - 1], list(range(L - 5, L))):
                got = reduced_density_matrix(x, keep)
                if rank == 0:
                    check(f'rdm {keep}', np.allclose(got, oracle.rdm(xfull, osub, keep), atol=1e-12))
                else:
                    check(f'rdm {keep} off rank 0', got.shape == (1, 1) and got[0, 0] == -1)
        H.destroy_mat()

    dist.barrier()
    if failures:
        print('\n'.join(failures), flush=True)
        sys.exit(1)
    if rank == 0:
        print(f'multi-gpu worker ok on {world} ranks', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
