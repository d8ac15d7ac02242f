"""Multi-GPU host logic: one process per GPU, independent channels sharded over ranks, no data-path collective.

The reference has no parallelism; its unit of independence is the filter object, i.e. the channel.
Channel c of a job lives on rank c % world (b2d_shard_count in the C-ABI).  The only exchange on the path is the
coefficient set at load() time: rank `root` owns it and b2d_fir_load broadcasts it with one ncclBroadcast over the
engine's own communicator.  torch.distributed is used for rendezvous only: it carries the 128-byte NCCL unique id
from rank 0 to the other ranks (any backend -- the CPU tests run this with gloo).
"""
import os

from . import _lib as L


def rank_world():
    """(rank, world, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def local_channels(n_channels, rank, world):
    """Global channel ids served by `rank` (c % world == rank), in local order."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    return list(range(rank, int(n_channels), world))


def owner_of(channel, world):
    """(rank, local index) of a global channel id."""
    return channel % world, channel // world


def exchange_unique_id(make_id, rank, src=0):
    """Rank `src` creates the communicator id (make_id()), every rank returns the same bytes.
    Needs an initialised torch.distributed process group when world > 1."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return make_id()
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return bytes(box[0])


def make_comm(rank, world, device):
    """The engine's NCCL communicator for this rank (None when world == 1)."""
    if world == 1:
        return None
    from .filters import Comm
    uid = exchange_unique_id(Comm.unique_id, rank)
    return Comm(uid, rank, world, device)


def check_partition(n_channels, world):
    """Every channel is served exactly once and the C-ABI agrees with the Python view (used by the tests)."""
    import ctypes as C
    lib = L.load()
    seen = []
    for r in range(world):
        ids = local_channels(n_channels, r, world)
        n = C.c_uint32(0)
        L.check(lib.b2d_shard_count(int(n_channels), r, world, C.byref(n)))
        if n.value != len(ids):
            raise AssertionError((r, n.value, ids))
        seen += ids
    return sorted(seen) == list(range(n_channels))
