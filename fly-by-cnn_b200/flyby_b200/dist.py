"""Multi-GPU plumbing: one process per GPU, work sharded by viewpoint or by mesh.

The feature path needs no exchange: every rank renders its own
views / meshes.  The only collective is the integer sum of label votes.  It runs
either through the library's own NCCL communicator on the context stream
(``init_nccl`` + ``Context.votes_allreduce``: enqueued right behind the scatter
kernel, no host synchronisation) or through ``torch.distributed`` (``gloo`` for
the CPU tests, ``nccl`` otherwise).
"""
import numpy as np


def world():
    """(rank, world_size) of the default torch.distributed group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous block [begin, end) of n items for this rank (sizes differ by at most one)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def shard_items(items, rank, world_size):
    """The items (meshes) this rank owns: contiguous blocks, like BASELINE config 3 (125 meshes per GPU)."""
    b, e = shard_range(len(items), rank, world_size)
    return list(items[b:e])


def init_nccl(ctx, group=None):
    """Create the library-side NCCL communicator: rank 0 makes the unique id, torch.distributed carries it."""
    import torch.distributed as dist
    rank, ws = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.nccl_init(box[0], rank, ws)
    return rank, ws


def allreduce_votes(votes, ctx=None, group=None):
    """Sum the vote table over all ranks, in place where possible; returns the summed table.

    * torch CUDA tensor + ``ctx`` with an NCCL communicator -> ncclAllReduce on the context stream;
    * torch tensor + initialised torch.distributed -> dist.all_reduce (nccl or gloo);
    * numpy array -> staged through torch (gloo);
    * single process -> returned unchanged.
    Integer sums make the result independent of the number and order of ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return votes
    if isinstance(votes, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(votes))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.numpy()
    if votes.is_cuda and ctx is not None and getattr(ctx, "_nccl_ready", False):
        ctx.votes_allreduce(votes)
        return votes
    dist.all_reduce(votes, op=dist.ReduceOp.SUM, group=group)
    return votes


def gather_view_slabs(local, n_views_total, group=None):
    """Concatenate per-rank view slabs [n_local, ...] into [n_views_total, ...] on every rank
    (only needed when one process wants the whole tensor; shards are written independently otherwise)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    is_np = isinstance(local, np.ndarray)
    t = torch.from_numpy(local) if is_np else local
    sizes = [e - b for b, e in (shard_range(n_views_total, r, ws) for r in range(ws))]
    pad = max(sizes)  # all_gather wants equal shapes: pad the short shards by at most one view
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[:t.shape[0]] = t
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf, group=group)
    out = torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
    return out.numpy() if is_np else out
