"""
Host-side description of how a state vector is sharded over GPUs (one process
per GPU): rank r of 2^p owns the indices whose top p index bits equal r --
the reference's power-of-two block layout (``_backend/bpetsc_template_2.c:768-797``).

``shard_plan`` wraps the host-only C-ABI entry ``dnm_shard_plan`` (no device
needed) and is what the multi-rank CPU tests exercise; ``init_comm`` boots the
library's NCCL communicator from a ``torch.distributed`` process group.
"""
import ctypes as C

import numpy as np

from . import _capi


def shard_plan(n_index_bits, nranks, rank, index_masks):
    """For each index-space flip mask: (partner rank, local mask)."""
    masks = _capi.as_i64(index_masks)
    partner = np.empty(masks.size, dtype=np.int32)
    local = np.empty(masks.size, dtype=np.int64)
    _capi.check(_capi.lib().dnm_shard_plan(int(n_index_bits), int(nranks), int(rank), masks.size, _capi.ip(masks),
                                            partner.ctypes.data_as(C.POINTER(C.c_int32)), _capi.ip(local)))
    return partner, local


def local_range(n_index_bits, nranks, rank):
    n = 1 << n_index_bits
    per = n // nranks
    return rank * per, (rank + 1) * per


def init_comm(dist=None):
    """Create the library communicator for the current ``torch.distributed`` group
    (rank 0 makes the NCCL id, the group broadcasts it).  Call after ``ensure_gpu``."""
    if dist is None:
        import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = [None]
    if rank == 0:
        buf = C.create_string_buffer(128)
        _capi.check(_capi.lib().dnm_comm_unique_id(buf))
        ident[0] = buf.raw
    dist.broadcast_object_list(ident, src=0)
    _capi.check(_capi.lib().dnm_comm_init(rank, world, ident[0]))
    return rank, world
