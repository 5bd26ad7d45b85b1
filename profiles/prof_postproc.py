"""Timing of the label post-processing on one B200 beside the oracle (the reference's own Python loops are
far slower than this C restatement).  python profiles/prof_postproc.py [mesh Lm]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np
import torch
from flyby_b200 import _cabi, synth
from oracle import oracle as orc

lm = int(sys.argv[1]) if len(sys.argv) > 1 else 100
P, F = synth.dental_crown(lm, 2)
V = len(P)
rng = np.random.default_rng(0)
lab = np.argmax(P @ rng.normal(size=(5, 3)).T, axis=1).astype(np.int32)
flip = rng.random(V) < 0.05
lab[flip] = rng.integers(0, 5, flip.sum()).astype(np.int32)
ctx = _cabi.Context(0)
ctx.set_mesh(P, F, skip_normalise=True); ctx.build()
dev = torch.from_numpy(lab).cuda()
torch.cuda.synchronize()


def gpu(fn):
    x = dev.clone(); fn(x); ctx.synchronize()
    x = dev.clone(); torch.cuda.synchronize(); t = time.perf_counter(); fn(x); ctx.synchronize()
    return (time.perf_counter() - t) * 1e3, x.cpu().numpy()


def cpu(fn):
    t = time.perf_counter(); r = fn(); return (time.perf_counter() - t) * 1e3, r


rows = [("remove_islands(label 1, 200)", lambda x: ctx.labels_remove_islands(x, 1, 200, True), lambda: orc.pp_remove_islands(F, V, lab, 1, 200, True)),
        ("connectivity(label 2)", lambda x: ctx.labels_connectivity(x, 2, 10), lambda: orc.pp_connectivity(F, V, lab, 2, 10)[0]),
        ("dilate(label 3, 2 sweeps)", lambda x: ctx.labels_dilate(x, 3, 2), lambda: orc.pp_dilate(F, V, lab, 3, 2)),
        ("erode(label 4, until gone)", lambda x: ctx.labels_erode(x, 4), lambda: orc.pp_erode(F, V, lab, 4))]
print("mesh: %d points, %d cells" % (V, len(F)))
for name, g, c in rows:
    tg, rg = gpu(g); tc, rc = cpu(c)
    print("%-32s gpu %8.3f ms   oracle (C, 1 thread) %9.2f ms   identical %s" % (name, tg, tc, np.array_equal(rg, rc)))
