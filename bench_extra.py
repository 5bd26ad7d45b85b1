"""Extra records of the bench line (bench.py imports this; the headline keys are untouched).

by_view          BASELINE configs[4]: 2M-triangle mesh (Lm = 316), icosahedron subdivision 4 (162 views), res 1024,
                 sharded BY VIEWPOINT: every rank holds mesh + LBVH and renders shard_range(162) -- strong scaling,
                 time = max over ranks, an additive device-side digest summed over ranks must equal the digest of
                 all 162 views rendered by rank 0 alone.
votes_allreduce  BASELINE configs[3]: fly-by normals + depth views (dental-crown mesh, 200 000 cells, subdivision 2,
                 res 224) -> torchvision vgg19(weights=None) features + 1x1-conv head + upsample, K = 34 -> argmax ->
                 integer vote scatter -> ncclAllReduce of the [200 000, 34] int32 table on the context stream
                 (library-side communicator) -> argmax per cell.  Views are sharded over the ranks; the summed table's
                 digest must equal the table rank 0 computes from all views alone.
plane_scale_2    the headline workload with planeScaleFactor 2.0 (whole surface in view) beside the reference default 1.0.
tolerance_mode   the headline workload with FLYBY_SHADE_TOLERANCE: id mismatches (must be 0) and the measured errors.
copy_ceiling     what N concurrent plain pinned cudaMemcpyAsync D2H copies of the e2e payload reach on this node.
"""
import os
import time

import numpy as np


def _digest(torch, ids, feat):
    """additive over views / ranks: int64 sums of the ids and of the feature bit patterns"""
    return int(ids.to(torch.int64).sum()), int(feat.view(torch.int32).to(torch.int64).sum())


def _allreduce_max(torch, dist, world, dev, vals):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def _allreduce_sum_i64(torch, dist, world, dev, vals):
    t = torch.tensor(vals, dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t]


def by_view(torch, dist, _cabi, synth, fdist, world, rank, local, dev, stream, barrier, steps=3):
    Lm, L, res, radius = 316, 4, 1024, 2.0
    P, F = synth.bumpy_sphere(Lm, seed=5)
    ctx = _cabi.Context(local, stream=stream.cuda_stream)
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    torch.cuda.synchronize()
    ctx.set_mesh(tp, tf)
    ctx.build()
    n = ctx.views_icosahedron(L, radius)
    out = {}
    for ps in (2.0, 1.0):
        opts = ctx.render_opts(use_z=1, split_z=1, plane_scale=ps)
        b, e = fdist.shard_range(n, rank, world)
        feat = torch.empty((e - b, res, res, 4), dtype=torch.float32, device=dev)
        ids = torch.empty((e - b, res, res), dtype=torch.int32, device=dev)
        ctx.render_into(res, opts, feat, ids, view_begin=b, view_end=e)   # warm-up (allocations)
        ctx.synchronize()
        barrier()
        ms = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.render_into(res, opts, feat, ids, view_begin=b, view_end=e)
            e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        st = ctx.stats()
        barrier()
        ms_rank = sorted(ms)[len(ms) // 2]
        (ms_max,) = _allreduce_max(torch, dist, world, dev, [ms_rank])
        d_ids, d_feat = _allreduce_sum_i64(torch, dist, world, dev, list(_digest(torch, ids, feat)))
        hits = _allreduce_sum_i64(torch, dist, world, dev, [int((ids >= 0).sum())])[0]
        rec = {"plane_scale": ps, "ms_render_max_over_ranks": ms_max, "views_per_s": n / (ms_max * 1e-3),
               "mrays_per_s": n * res * res / (ms_max * 1e-3) / 1e6, "hit_fraction": hits / float(n * res * res),
               "views_this_rank": e - b, "digest": "%x.%x" % (d_ids & (2 ** 64 - 1), d_feat & (2 ** 64 - 1)),
               "search_ms_rank0": st["ms_trace"], "resolve_ms_rank0": st["ms_resolve"]}
        del feat, ids
        if world > 1:   # rank 0 alone renders all views (not timed): the sharded digests must add up to its digest
            ok = [1]
            if rank == 0:
                feat = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
                ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
                ctx.render_into(res, opts, feat, ids)
                ctx.synchronize()
                a, bb = _digest(torch, ids, feat)
                ok = [1 if (a == d_ids and bb == d_feat) else 0]
                del feat, ids
            ok = _allreduce_sum_i64(torch, dist, world, dev, ok if rank == 0 else [0])
            rec["digest_equals_single_rank"] = bool(ok[0] == 1)
        else:
            rec["digest_equals_single_rank"] = True
        out["plane_scale_%g" % ps] = rec
    ctx.close()
    torch.cuda.empty_cache()
    return {"workload": "BASELINE configs[4]: %d triangles, icosahedron subdivision %d (%d views), res %d, radius %g, C = 4; "
                        "every rank holds mesh + LBVH and renders shard_range(views): strong scaling, no collective on the "
                        "data path" % (len(F), L, n, res, radius),
            "scaling": "strong", "triangles": int(len(F)), "views": n, "resolution": res, **out}


class _FlyByVGG(object):
    """vgg19(weights=None).features on the 4-channel fly-by views (first conv widened to 4 inputs), 1x1-conv head with K
    outputs, bilinear upsample back to res x res: per-pixel logits [n, K, H, W].  Random init, torch.manual_seed(0)."""

    def __init__(self, torch, dev, K, channels=4):
        import torchvision
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True
        torch.manual_seed(0)
        vgg = torchvision.models.vgg19(weights=None)
        feats = vgg.features
        first = feats[0]
        feats[0] = torch.nn.Conv2d(channels, first.out_channels, kernel_size=3, padding=1)
        self.torch = torch
        self.features = feats.to(dev).eval()
        self.head = torch.nn.Conv2d(512, K, kernel_size=1).to(dev).eval()

    def __call__(self, x):
        torch = self.torch
        with torch.no_grad():
            y = self.head(self.features(x))
            return torch.nn.functional.interpolate(y, size=x.shape[-2:], mode="bilinear", align_corners=False)

    def eval(self):
        return self


def votes_allreduce(torch, dist, _cabi, synth, fdist, world, rank, local, dev, stream, barrier, steps=5):
    from flyby_b200 import labeling
    K, res, L, radius = 34, 224, 2, 2.0
    P, F = synth.dental_crown(100, seed=1)
    ctx = _cabi.Context(local, stream=stream.cuda_stream)
    if world > 1:
        fdist.init_nccl(ctx)
    else:
        ctx.nccl_init(ctx.nccl_unique_id(), 0, 1)
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    torch.cuda.synchronize()
    ctx.set_mesh(tp, tf)
    ctx.build()
    n = ctx.views_icosahedron(L, radius)
    opts = ctx.render_opts(use_z=1, split_z=1, plane_scale=2.0)
    model = _FlyByVGG(torch, dev, K)
    b, e = fdist.shard_range(n, rank, world)
    feat = torch.empty((max(e - b, 1), res, res, 4), dtype=torch.float32, device=dev)
    ids = torch.empty((max(e - b, 1), res, res), dtype=torch.int32, device=dev)
    votes = torch.zeros((len(F), K), dtype=torch.int32, device=dev)

    def local_votes(vb, ve, table):
        """render [vb, ve) -> VGG19 -> argmax -> scatter into `table`; returns (render, cnn, scatter) ms"""
        nv = ve - vb
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record(stream)
        if nv > 0:
            ctx.render_into(res, opts, feat[:nv], ids[:nv], view_begin=vb, view_end=ve)
        ev[1].record(stream)
        lab = None
        if nv > 0:
            # one view per forward pass: the logits of a view then do not depend on how many views this rank holds
            # (cuDNN picks its algorithms per batch shape), so the vote table is identical for every rank count
            with torch.cuda.stream(stream):
                lab = torch.empty((nv, res, res), dtype=torch.int32, device=dev)
                x = feat[:nv].permute(0, 3, 1, 2)
                for i in range(nv):
                    lab[i] = model(x[i:i + 1].contiguous()).argmax(dim=1)[0].to(torch.int32)
        ev[2].record(stream)
        if nv > 0:
            ctx.backproject(ids[:nv], lab, K, votes=table)
        ev[3].record(stream)
        ev[3].synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]

    local_votes(b, e, votes)                       # warm-up (cuDNN autotune, allocations)
    ctx.votes_allreduce(votes)
    ctx.synchronize()
    rows = []
    for _ in range(steps):
        votes.zero_()
        torch.cuda.synchronize()
        barrier()
        t_r, t_c, t_s = local_votes(b, e, votes)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.votes_allreduce(votes)                 # ncclAllReduce(int32, sum) on the context stream
        e1.record(stream)
        e1.synchronize()
        rows.append((t_r, t_c, t_s, e0.elapsed_time(e1)))
    med = [sorted(r[i] for r in rows)[len(rows) // 2] for i in range(4)]
    med = _allreduce_max(torch, dist, world, dev, med)
    cell_labels = torch.empty(len(F), dtype=torch.int32, device=dev)
    ctx.votes_argmax(votes, default=33, out=cell_labels)
    ctx.synchronize()
    d_sum = int(votes.to(torch.int64).sum())
    w = torch.arange(votes.numel(), device=dev, dtype=torch.int64).reshape(votes.shape) % 1000003
    d_weighted = int((votes.to(torch.int64) * w).sum())
    equal = True
    if world > 1:      # rank 0 recomputes the table from ALL views alone: the all-reduced table must equal it
        ok = [0]
        if rank == 0:
            full = torch.zeros_like(votes)
            feat_all = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
            ids_all = torch.empty((n, res, res), dtype=torch.int32, device=dev)
            feat, ids = feat_all, ids_all
            local_votes(0, n, full)
            ok = [1 if bool(torch.equal(full, votes)) else 0]
        equal = bool(_allreduce_sum_i64(torch, dist, world, dev, ok)[0] == 1)
    nbytes = votes.numel() * 4
    bus = (2.0 * (world - 1) / world) * nbytes / (med[3] * 1e-3) / 1e9 if world > 1 else 0.0
    ctx.close()
    torch.cuda.empty_cache()
    return {"workload": "BASELINE configs[3]: dental-crown mesh (%d cells), icosahedron subdivision %d (%d views), res %d, "
                        "plane scale 2.0 -> torchvision vgg19(weights=None).features (4 input channels) + 1x1 head + "
                        "bilinear upsample, K = %d, one view per forward pass -> argmax -> vote scatter -> ncclAllReduce(int32 sum) of the [%d, %d] "
                        "table on the context stream; views sharded over the ranks" % (len(F), L, n, res, K, len(F), K),
            "views": n, "views_this_rank": e - b, "classes": K, "table_bytes": nbytes,
            "render_ms": med[0], "cnn_ms": med[1], "vote_scatter_ms": med[2], "allreduce_ms": med[3],
            "allreduce_bus_gb_s": bus, "n_ranks": world,
            "table_digest": "%x.%x" % (d_sum, d_weighted & (2 ** 64 - 1)), "votes_total": d_sum,
            "cells_labelled": int((cell_labels != 33).sum()),
            "table_equals_single_rank": equal}


def render_variant(torch, ctx, stream, flush, dev_mesh, res, opts, feat, ids, steps=8):
    """median render time of one resident mesh under `opts` (single context, per-render events, L2 flushed)"""
    P, F = dev_mesh
    ctx.set_mesh(P, F)
    ctx.build()
    ms, tr = [], []
    for k in range(steps + 2):
        flush.zero_()
        torch.cuda.synchronize()
        ctx.render_into(res, opts, feat, ids)
        ctx.synchronize()
        st = ctx.stats()
        if k >= 2:
            ms.append(st["ms_render"])
            tr.append(st["ms_trace"])
    ms.sort()
    tr.sort()
    st = ctx.stats()
    return ms[len(ms) // 2], tr[len(tr) // 2], st


def tolerance_errors(torch, feat_exact, ids_exact, feat_tol, ids_tol):
    """id mismatches and the largest errors of the tolerance render against the exact one (same mesh, same options)"""
    mism = int((ids_exact != ids_tol).sum())
    hit = ids_exact >= 0
    fe, ft = feat_exact[hit].to(torch.float64), feat_tol[hit].to(torch.float64)
    depth_rel = float(((ft[:, 3] - fe[:, 3]).abs() / fe[:, 3].abs().clamp_min(1e-30)).max()) if hit.any() else 0.0
    normal_abs = float((ft[:, :3] - fe[:, :3]).abs().max()) if hit.any() else 0.0
    miss_equal = bool(torch.equal(feat_exact[~hit], feat_tol[~hit]))
    return {"id_mismatches": mism, "max_depth_rel_err": depth_rel, "max_normal_abs_err": normal_abs,
            "background_equal": miss_equal, "bar": "ids bit-exact; depth 1e-5 relative; normal channels 1e-5 (unit range)"}


def copy_ceiling(torch, dist, world, dev, nbytes, host_buf, reps=6):
    """N ranks copy `nbytes` device -> pinned host at the same time, one plain cudaMemcpyAsync each: the node's D2H
    ceiling for the e2e payload.  Returns (GB/s of the slowest rank, aggregate GB/s)."""
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = host_buf.view(torch.uint8).reshape(-1)[:nbytes]
    s = torch.cuda.Stream(device=dev)
    best = 0.0
    for k in range(reps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            dst.copy_(src, non_blocking=True)
        s.synchronize()
        dt = time.perf_counter() - t0
        if k > 0:
            best = max(best, nbytes / dt / 1e9)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    tmin, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    del src
    return float(tmin), float(tsum)
