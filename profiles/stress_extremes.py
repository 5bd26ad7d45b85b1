"""Extremes of triangle size against the oracle: many triangles per pixel (2 M triangles at 64^2 / 200^2) and a few
triangles covering a 2048^2 image.   python profiles/stress_extremes.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np
from flyby_b200 import _cabi, synth
from oracle import oracle as orc
ctx = _cabi.Context(0)
bad = 0
for name, mesh, res_list, sel in (("2M triangles", synth.bumpy_sphere(316, 5), (64, 200), [0, 7]),
                                  ("cube", synth.cube(), (2048,), [0, 5]),
                                  ("crown 20k", synth.dental_crown(32, 3), (1536,), [3])):
    P, F = mesh
    ctx.set_mesh(P, F); ctx.build()
    uv, nrm, _, _ = ctx.get_mesh()
    ctx.views_icosahedron(1, 2.0)
    views = ctx.get_views()
    for res in res_list:
        for ps in (1.0, 2.0):
            opts = ctx.render_opts(plane_scale=ps)
            t0 = time.time()
            feat, ids, t = ctx.render(res, opts, want_t=True, view_begin=0, view_end=12)
            st = ctx.stats()
            fo, io, to = orc.render(uv, F, nrm, views[sel], 2.0, res, plane_scale=ps, grid=True, grid_div=64, want_t=True)
            ok = np.array_equal(ids[sel], io) and np.array_equal(t[sel], to) and np.array_equal(feat[sel], fo)
            bad += 0 if ok else 1
            print("%-14s res %4d ps %.1f: render %.2f ms (12 views), hits %d, %s" % (name, res, ps, st["ms_render"], int((io >= 0).sum()), "ok" if ok else "MISMATCH"))
sys.exit(1 if bad else 0)
