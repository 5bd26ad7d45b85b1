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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_cpus(device_index):
    """(numa node, set of CPUs) the GPU hangs off, from sysfs; (None, empty set) where that is not exposed."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None, set()
        return node, _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None, set()


class near_gpu:
    """Context manager: run the body on the CPUs of the GPU's NUMA node, then restore the affinity.  Pinned
    staging buffers allocated inside are first-touched on that node, so that the copy-out of every rank of a
    multi-GPU batch extraction stays on its own socket (one process per GPU, by-mesh sharding).

        with near_gpu(local_rank):
            host = torch.empty(shape).pin_memory()
    """

    def __init__(self, device_index):
        self.node, self.cpus = gpu_numa_cpus(device_index)
        self.saved = None

    def __enter__(self):
        import os
        try:
            cur = os.sched_getaffinity(0)
            want = cur & self.cpus
            if want and want != cur:
                os.sched_setaffinity(0, want)
                self.saved = cur
        except (AttributeError, OSError):
            self.saved = None
        return self

    def __exit__(self, *exc):
        import os
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except OSError:
                pass
        return False


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
