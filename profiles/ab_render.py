"""A/B driver for kernel experiments: the bench workload's render (200 k triangles, 92 views, 512^2) through one or
more builds of libflyby_b200.so, interleaved, with a digest of the outputs so that a variant that changes results is
seen at once.
  python profiles/ab_render.py ab/A.so ab/B.so ...        (each library runs in its own subprocess)
Variants are built beside the in-tree library, which stays untouched:
  make -C fly-by-cnn_b200 -j LIB=../ab/A.so
  touch fly-by-cnn_b200/csrc/raster.cu; make -C fly-by-cnn_b200 -j EXTRA="-DRT_MINB=10" LIB=../ab/B.so
(ab/ is git-ignored and travels to the GPU box with gpurun.)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))


def one(lib, reps):
    import torch
    from flyby_b200 import _cabi, synth
    _cabi.LIB_PATH = os.path.abspath(lib)
    dev = torch.device("cuda", 0)
    P, F = synth.bumpy_sphere(100, seed=0)
    ctx = _cabi.Context(0)
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.set_mesh(tp, tf); ctx.build()
    n = ctx.views_icosahedron(3, 2.0)
    opts = ctx.render_opts(use_z=1, split_z=1, shade_tolerance=os.environ.get("AB_TOLERANCE") == "1",
                           plane_scale=float(os.environ.get("AB_PLANE_SCALE", "1.0")))
    feat = torch.empty((n, 512, 512, 4), dtype=torch.float32, device=dev)
    ids = torch.empty((n, 512, 512), dtype=torch.int32, device=dev)
    r, tr = [], []
    for it in range(reps + 3):
        flush.zero_(); torch.cuda.synchronize()
        ctx.render_into(512, opts, feat, ids); ctx.synchronize()
        st = ctx.stats()
        if it >= 3:
            r.append(st["ms_render"]); tr.append(st["ms_trace"])
    r.sort(); tr.sort()
    dig = int(ids.to(torch.int64).sum()) ^ int(feat.view(torch.int32).to(torch.int64).sum())
    print("%-12s render median %.4f min %.4f | search median %.4f | resolve %.4f | digest %x" %
          (os.path.basename(lib), r[len(r) // 2], r[0], tr[len(tr) // 2], r[len(r) // 2] - tr[len(tr) // 2], dig), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(sys.argv[2], int(sys.argv[3]))
    else:
        for rnd in range(2):
            for lib in sys.argv[1:]:
                subprocess.check_call([sys.executable, __file__, "--one", lib, "15"])
