"""CPU, world_size 2 and 4 over gloo: the sharding layout the library plans
(dnm_shard_plan) reproduces the single-rank product when every rank combines the
shards it is told to fetch.  The per-shard arithmetic is done by the oracle; what
is under test is the partition / partner / local-mask logic of the host layer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, L, sub_type, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import torch
    import torch.distributed as dist

    import oracle
    from dynamite_b200 import msc_tools
    from dynamite_b200.hamiltonians import build_hamiltonian
    from dynamite_b200.sharding import local_range, shard_plan
    from helpers import rand_state

    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        H = build_hamiltonian(name, L)
        H.reduce_msc()
        terms = [(int(m), int(s), complex(c)) for m, s, c in H.msc]
        spec = {'type': sub_type, 'L': L, 'space': 0}
        osub = oracle.Subspace(spec)
        nbits = L if sub_type == 'full' else L - 1
        x = rand_state(osub.dim, 42)                                  # same on every rank
        want = oracle.matmult(oracle.Msc.from_terms(terms), osub, osub, x)

        # index-space rewrite of the terms (Full: identity; Parity: drop odd masks, halve)
        idx_terms = []
        for m, s, c in terms:
            if sub_type == 'parity':
                if bin(m).count('1') & 1:
                    continue
                full = (1 << nbits) - 1
                sp = s >> 1
                if s & 1:
                    sp ^= full       # bit 0 of the state is parity(idx): fold it into the sign mask
                idx_terms.append((m >> 1, sp & full, c))
            else:
                idx_terms.append((m, s, c))
        masks = np.array(sorted({t[0] for t in idx_terms}), dtype=np.int64)
        partner, local = shard_plan(nbits, world, rank, masks)
        a, b = local_range(nbits, world, rank)
        nloc_bits = nbits - int(np.log2(world))
        assert b - a == 1 << nloc_bits

        # halo exchange: every rank sends its shard to the ranks that gather from it (XOR pattern)
        mine = torch.from_numpy(np.ascontiguousarray(x[a:b]).view(np.float64).copy())
        shards = {rank: mine}
        for h in sorted({int(p) ^ rank for p in partner} - {0}):
            peer = rank ^ h
            recv = torch.empty_like(mine)
            if rank < peer:
                dist.send(mine, peer)
                dist.recv(recv, peer)
            else:
                dist.recv(recv, peer)
                dist.send(mine, peer)
            shards[peer] = recv
        y = np.zeros(b - a, dtype=np.complex128)
        lowmask = (1 << nloc_bits) - 1
        for peer in shards:
            grp = []
            for (m, s, c) in idx_terms:
                k = int(np.searchsorted(masks, m))
                if int(partner[k]) != peer:
                    continue
                # sign bits above the local index see the PEER's rank bits (column state)
                hi_sign = bin((s >> nloc_bits) & peer).count('1') & 1
                grp.append((int(local[k]), s & lowmask, -c if hi_sign else c))
            if not grp:
                continue
            xs = shards[peer].numpy().view(np.complex128)
            # column-state signs inside the shard: the MSC definition on the local bits.  (The
            # C oracle's real/imaginary shortcut needs whole Hermitian terms, truncated ones are not.)
            idx = np.arange(1 << nloc_bits)
            y += oracle.msc_to_dense(grp, idx, idx) @ xs
        err = np.linalg.norm(y - want[a:b]) / np.linalg.norm(want[a:b])
        result_q.put((rank, float(err), len(shards)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
@pytest.mark.parametrize('name,L,sub_type', [('MBL', 9, 'full'), ('long_range', 8, 'full'), ('SYK', 5, 'parity'),
                                             ('heisenberg', 9, 'parity')])
def test_sharded_product_matches_single_rank(world, name, L, sub_type):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, L, sub_type, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    results = sorted(q.get(timeout=5) for _ in range(world))
    for rank, err, nshards in results:
        assert err < 1e-13, (rank, err)
    if name in ('long_range', 'SYK'):
        assert max(r[2] for r in results) > 1   # some masks really cross shards


def test_shard_plan_basics():
    from dynamite_b200._capi import BackendError
    from dynamite_b200.sharding import shard_plan
    partner, local = shard_plan(6, 4, 1, [0, 0b000011, 0b010000, 0b110001, 0b100000])
    assert partner.tolist() == [1, 1, 0, 2, 3]
    assert local.tolist() == [0, 3, 0, 1, 0]
    with pytest.raises(BackendError):
        shard_plan(6, 3, 0, [1])
    with pytest.raises(BackendError):
        shard_plan(6, 4, 0, [1 << 6])
