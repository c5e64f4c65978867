"""Multi-GPU plumbing: one process per GPU (torchrun), samples sharded per class, one NCCL
all-reduce of the partial gradient per bond inside the library.  torch.distributed is
used only to agree on the NCCL unique id."""
import numpy as np


def shard_ranges(class_counts, rank, world):
    """Each class's contiguous (sorted) sample range is split evenly over the ranks.
    Returns [(begin, end)] per class in GLOBAL sample indices."""
    out = []
    off = 0
    for n in class_counts:
        n = int(n)
        b = off + (n * rank) // world
        e = off + (n * (rank + 1)) // world
        out.append((b, e))
        off += n
    return out


def shard_samples(X_TxN, class_counts, rank, world):
    """Local (T, N_local) slice (still class-sorted) and its local class counts."""
    rngs = shard_ranges(class_counts, rank, world)
    idx = np.concatenate([np.arange(b, e) for b, e in rngs]) if rngs else np.zeros(0, dtype=np.int64)
    return X_TxN[:, idx], np.array([e - b for b, e in rngs], dtype=np.int64), idx


def rank_world():
    try:
        import torch.distributed as td
        if td.is_available() and td.is_initialized():
            return td.get_rank(), td.get_world_size()
    except Exception:
        pass
    return 0, 1


def init_comm(ctx):
    """Create the library's NCCL communicator over the ranks of the default process group."""
    rank, world = rank_world()
    if world == 1:
        return rank, world
    import torch.distributed as td
    box = [ctx.comm_unique_id() if rank == 0 else None]
    td.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)
    return rank, world


def shard_instances(n, rank, world):
    """Imputation / classification shard by instance with no data-path collective: rank r takes the
    contiguous block [n r / world, n (r + 1) / world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_instances(local):
    """Concatenate every rank's block (leading axis = instances) in rank order on every rank: host-language plumbing
    (torch.distributed object all-gather), results only -- the series, masks and cores never cross ranks."""
    rank, world = rank_world()
    if world == 1:
        return local
    import torch.distributed as td
    box = [None] * world
    td.all_gather_object(box, local)
    if isinstance(local, tuple):
        return tuple(np.concatenate([b[k] for b in box], axis=0) for k in range(len(local)))
    return np.concatenate(box, axis=0)
