"""Measure BASELINE configs 1, 2, 3 (one mesh of its batch), 4 and 5 on one B200 (device-resident outputs, CUDA events),
each at the reference's default planeScaleFactor 1.0 and at 2.0 (whole surface in view), and check a
sample of every config against the oracle.  Writes gpurun_out/configs.json (copied to profiles/r02_configs.json).
  python profiles/run_configs.py [c1 c2 c3 c4 c5]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from flyby_b200 import _cabi, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CONFIGS = {
    "c1": dict(mesh=lambda: synth.bumpy_sphere(71, 0), views=("ico", 2), radius=2.0, res=512, use_z=1, split_z=1),
    "c2": dict(mesh=lambda: synth.dental_crown(100, 1), views=("spiral", 64, 4), radius=2.0, res=224, use_z=1, split_z=0),
    "c3": dict(mesh=lambda: synth.bumpy_sphere(100, 0), views=("ico", 3), radius=2.0, res=512, use_z=1, split_z=1),
    "c5": dict(mesh=lambda: synth.bumpy_sphere(316, 5), views=("ico", 4), radius=2.0, res=1024, use_z=1, split_z=1),
}


def run(name, cfg, check_views=2, plane_scale=1.0, tolerance=False):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = _cabi.Context(0, stream=stream.cuda_stream)
    P, F = cfg["mesh"]()
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    torch.cuda.synchronize()
    if cfg["views"][0] == "ico":
        n = ctx.views_icosahedron(cfg["views"][1], cfg["radius"])
    else:
        n = ctx.views_spiral(cfg["views"][1], cfg["views"][2], cfg["radius"])
    res = cfg["res"]
    opts = ctx.render_opts(use_z=cfg["use_z"], split_z=cfg["split_z"], plane_scale=plane_scale, shade_tolerance=tolerance)
    C = ctx.channels(opts)
    feat = torch.empty((n, res, res, C), dtype=torch.float32, device=dev)
    ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
    times, builds = [], []
    for it in range(4):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        ctx.set_mesh(tp, tf)
        ctx.build()
        e1.record(stream)
        ctx.render_into(res, opts, feat, ids)
        e2.record(stream)
        e2.synchronize()
        builds.append(e0.elapsed_time(e1))
        times.append(e1.elapsed_time(e2))
    st = ctx.stats()
    out = dict(config=name, plane_scale=plane_scale, shading="tolerance" if tolerance else "exact", triangles=int(len(F)),
               views=n, resolution=res, channels=C, rays=n * res * res,
               build_ms=min(builds[1:]), render_ms=min(times[1:]), trace_ms=st["ms_trace"], resolve_ms=st["ms_resolve"],
               mrays_per_s=n * res * res / (min(times[1:]) * 1e-3) / 1e6, views_per_s=n / (min(times[1:]) * 1e-3),
               hit_fraction=st["n_hits"] / max(1, st["n_rays"]))
    # parity on a sample of views against the oracle (brute-force-equivalent bucket grid)
    uv, nrm, _, _ = ctx.get_mesh()
    views = ctx.get_views()
    sel = sorted(set(np.linspace(0, n - 1, check_views).astype(int).tolist()))
    t0 = time.time()
    fo, io = orc.render(uv, F, nrm, views[sel], cfg["radius"], res, plane_scale=plane_scale, use_z=cfg["use_z"],
                        split_z=cfg["split_z"], grid=True, grid_div=64)
    got_i = ids[sel].cpu().numpy()
    got_f = feat[sel].cpu().numpy()
    out["parity"] = dict(views_checked=sel, rays_checked=int(io.size), id_mismatches=int((io != got_i).sum()),
                         features_bit_equal=bool(np.array_equal(fo, got_f)),
                         max_abs_err=float(np.max(np.abs(fo - got_f))) if io.size else 0.0,
                         oracle_seconds=round(time.time() - t0, 2))
    ctx.close()
    return out


def run_c4():
    """Config 4: the C2 mesh, 42 views at 224^2, normals + depth -> random-init torchvision VGG19 features + 1x1
    head (library code) -> per-pixel argmax over K = 34 labels -> votes [T, 34] -> argmax per cell -> cell label to
    point.  The vote table is compared with the oracle's.  (The NCCL all-reduce of the table over 1/2/4/8 ranks is
    in the bench line: votes_allreduce.)"""
    from flyby_b200 import labeling
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = _cabi.Context(0, stream=stream.cuda_stream)
    P, F = synth.dental_crown(100, 1)
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    torch.cuda.synchronize()
    ctx.set_mesh(tp, tf)
    ctx.build()
    n = ctx.views_icosahedron(2, 2.0)
    res, K = 224, 34
    opts = ctx.render_opts(use_z=1, split_z=1, plane_scale=2.0)
    sys.path.insert(0, ROOT)
    import bench_extra
    model = bench_extra._FlyByVGG(torch, dev, K)   # torchvision vgg19(weights=None).features + 1x1 head + upsample
    feat = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
    ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
    best = None
    for it in range(4):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        voter = labeling.CellVoter(ctx, K, use_torch=True)
        ev[0].record(stream)
        ctx.render_into(res, opts, feat, ids)
        ev[1].record(stream)
        with torch.no_grad():
            logits = model(feat.permute(0, 3, 1, 2).contiguous())
            labels = logits.argmax(dim=1).to(torch.int32).contiguous()
        ev[2].record(stream)
        voter.add(ids, labels)
        ev[3].record(stream)
        cell_labels = voter.cell_labels(default=33)
        point_labels = voter.point_labels(default=33)
        ev[4].record(stream)
        ev[4].synchronize()
        t = [ev[k].elapsed_time(ev[k + 1]) for k in range(4)]
        if it > 0 and (best is None or sum(t) < sum(best)):
            best = t
    votes_ref = orc.backproject(ids.cpu().numpy(), labels.cpu().numpy(), len(F), K)
    lab_ref = orc.votes_argmax(votes_ref, default=33)
    out = dict(config="c4", triangles=int(len(F)), views=n, resolution=res, classes=K,
               render_ms=best[0], cnn_ms=best[1], backproject_ms=best[2], argmax_and_points_ms=best[3],
               parity=dict(votes_equal=bool(np.array_equal(voter.votes.cpu().numpy(), votes_ref)),
                           cell_labels_equal=bool(np.array_equal(cell_labels.cpu().numpy(), lab_ref)),
                           voted_cells=int((votes_ref.sum(1) > 0).sum())))
    ctx.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    results = []
    for k in which:
        if k == "c4":
            results.append(run_c4())
            continue
        for ps in (1.0, 2.0):
            results.append(run(k, CONFIGS[k], plane_scale=ps))
        results.append(run(k, CONFIGS[k], plane_scale=1.0, tolerance=True))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(results, f, indent=1)
    for r in results:
        print(json.dumps(r))
