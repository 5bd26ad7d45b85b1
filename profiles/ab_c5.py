"""A/B on a config-5-like workload (2 M triangles, 12 views at 1024^2): python profiles/ab_c5.py ab/A.so ab/B.so"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))


def one(lib):
    import torch
    from flyby_b200 import _cabi, synth
    _cabi.LIB_PATH = os.path.abspath(lib)
    dev = torch.device("cuda", 0)
    P, F = synth.bumpy_sphere(316, 5)
    ctx = _cabi.Context(0)
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    torch.cuda.synchronize()
    ctx.set_mesh(tp, tf); ctx.build()
    n = ctx.views_icosahedron(1, 2.0)
    res = 1024
    feat = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
    ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
    for ps in (1.0, 2.0):
        opts = ctx.render_opts(use_z=1, split_z=1, plane_scale=ps)
        r, t = [], []
        for it in range(8):
            ctx.render_into(res, opts, feat, ids); ctx.synchronize()
            st = ctx.stats()
            if it >= 2:
                r.append(st["ms_render"]); t.append(st["ms_trace"])
        r.sort(); t.sort()
        print("%-10s ps %.1f render %.4f search %.4f digest %x" % (os.path.basename(lib), ps, r[len(r) // 2], t[len(t) // 2],
              int(ids.to(torch.int64).sum()) ^ int(feat.view(torch.int32).to(torch.int64).sum())), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        for rnd in range(2):
            for lib in sys.argv[1:]:
                subprocess.check_call([sys.executable, __file__, "--one", lib])
