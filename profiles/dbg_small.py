"""Warm timings of odd workloads (coarse mesh at high resolution).  python profiles/dbg_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np
from flyby_b200 import _cabi, synth
ctx = _cabi.Context(0)
for name, (P, F) in (("cube", synth.cube()), ("icosphere 320 tris", (synth.icosphere(4)[0].astype(np.float32), synth.icosphere(4)[1])),
                     ("bumpy 2000 tris", synth.bumpy_sphere(10, 0))):
    ctx.set_mesh(P, F); ctx.build()
    ctx.views_icosahedron(1, 2.0)
    for res, ps in ((512, 2.0), (2048, 2.0)):
        opts = ctx.render_opts(plane_scale=ps)
        for _ in range(3):
            feat, ids = ctx.render(res, opts)
        st = ctx.stats()
        print("%-20s res %4d: render %.2f ms for 12 views (search %.2f, resolve %.2f), %.2f Grays/s" % (
            name, res, st["ms_render"], st["ms_trace"], st["ms_resolve"], 12 * res * res / st["ms_render"] / 1e6), flush=True)
