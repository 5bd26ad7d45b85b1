"""Small driver for ncu: one mesh, build + 3 renders.
  python profiles/prof_render.py [views] [res] [mesh Lm] [ico level]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import torch
from flyby_b200 import _cabi, synth

views = int(sys.argv[1]) if len(sys.argv) > 1 else 92
res = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda", 0)
lm = int(sys.argv[3]) if len(sys.argv) > 3 else 100      # mesh: bumpy sphere, 20 lm^2 triangles
level = int(sys.argv[4]) if len(sys.argv) > 4 else 3     # icosahedron subdivision of the viewpoints
P, F = synth.bumpy_sphere(lm, seed=0)
ctx = _cabi.Context(0)
tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
torch.cuda.synchronize()
ctx.set_mesh(tp, tf)
ctx.build()
n = ctx.views_icosahedron(level, 2.0)
n = min(n, views)
opts = ctx.render_opts(use_z=1, split_z=1, shade_tolerance=os.environ.get("PROF_TOLERANCE") == "1",
                       plane_scale=float(os.environ.get("PROF_PLANE_SCALE", "1.0")))
feat = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
for it in range(2):
    ctx.render_into(res, opts, feat, ids, view_end=n)
    ctx.synchronize()
    print("render ms", ctx.stats()["ms_render"], "hits", ctx.stats()["n_hits"])
ctx.set_count_work(True)
ctx.render_into(res, opts, feat, ids, view_end=n)
st = ctx.stats()
print("work: nodes/ray %.2f tris/ray %.2f" % (st["nodes_visited"] / st["n_rays"], st["tris_tested"] / st["n_rays"]))
