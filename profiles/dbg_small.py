import sys
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/fly-by-cnn_b200")
import numpy as np
from flyby_b200 import _cabi, synth
P,F = synth.bumpy_sphere(12,0)
ctx=_cabi.Context(0); ctx.set_mesh(P,F); ctx.build(); ctx.views_icosahedron(1,2.0)
try:
    f,i=ctx.render(64, ctx.render_opts(), view_end=2)
    print("ok hits", (i>=0).sum())
except Exception as e:
    print("ERR", e)
