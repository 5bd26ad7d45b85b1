"""Does alternating two contexts (two streams) hide the build of mesh n+1 behind the render of mesh n?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np, torch
from flyby_b200 import _cabi, synth
dev = torch.device("cuda", 0)
meshes = [synth.bumpy_sphere(100, seed=k) for k in range(4)]
dm = [(torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)) for P, F in meshes]
torch.cuda.synchronize()
def make():
    s = torch.cuda.Stream(device=dev)
    c = _cabi.Context(0, stream=s.cuda_stream)
    n = c.views_icosahedron(3, 2.0)
    f = torch.empty((n, 512, 512, 4), dtype=torch.float32, device=dev); i = torch.empty((n, 512, 512), dtype=torch.int32, device=dev)
    return c, s, f, i
ctxs = [make() for _ in range(2)]
opts = ctxs[0][0].render_opts(use_z=1, split_z=1)
def step(k, which):
    c, s, f, i = ctxs[which]
    P, F = dm[k % 4]
    c.set_mesh(P, F); c.build(); c.render_into(512, opts, f, i)
for mode in ("one context", "two contexts"):
    for k in range(6): step(k, k % 2 if mode == "two contexts" else 0)
    torch.cuda.synchronize()
    t = time.perf_counter()
    K = 60
    for k in range(K): step(k, k % 2 if mode == "two contexts" else 0)
    torch.cuda.synchronize()
    print(mode, "%.3f ms/mesh" % ((time.perf_counter() - t) * 1e3 / K))
