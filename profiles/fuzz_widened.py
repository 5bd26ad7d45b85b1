"""Randomised parity of the rows added after the hot path: pinhole views, barycentric lookup, label post-processing,
vote back-projection -- against the oracle.   python profiles/fuzz_widened.py [n_cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import importlib.util
import numpy as np
from flyby_b200 import _cabi
from oracle import oracle as orc

spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "profiles", "fuzz_parity.py"))
fz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fz)


def run(n_cases, seed, context=None):
    rng = np.random.default_rng(seed)
    fz.rng = rng
    ctx = context or _cabi.Context(0)
    bad = 0
    t0 = time.time()
    for case in range(n_cases):
        P, F = fz.mesh()
        P = (P / max(1e-6, np.abs(P).max())).astype(np.float32)
        ctx.set_mesh(P, F, skip_normalise=True, consistency="off")
        ctx.build()
        uv, nrm, _, _ = ctx.get_mesh()
        V = len(P)
        ok = True
        # ---- pinhole views
        radius = float(rng.choice([1.5, 2.5, 6.0]))
        nv = ctx.views_icosahedron(1, radius)
        views = ctx.get_views()
        res = int(rng.choice([5, 24, 61]))
        angle = float(rng.choice([10.0, 30.0, 75.0]))
        zn, zf = float(rng.choice([0.01, 0.5])), radius + float(rng.choice([0.2, 2.0]))
        cfg = dict(use_z=int(rng.integers(0, 2)), split_z=int(rng.integers(0, 2)))
        enc = str(rng.choice(["unit01", "uint8", "signed"]))
        opts = ctx.render_opts(normal_encoding=enc, z_scale=float(rng.choice([1.0, 0.3])), **cfg)
        pf, pi, pd = ctx.render_perspective(res, opts, angle, zn, zf)
        of, oi, od = orc.render_perspective(uv, F, nrm, views, res, angle, zn, zf, normal_encoding={"unit01": 0, "uint8": 3, "signed": 2}[enc],
                                            zscale=opts.z_scale, grid=False, **cfg)
        ok = ok and np.array_equal(pi, oi) and np.array_equal(pd, od) and np.array_equal(pf, of, equal_nan=True)
        # ---- barycentric lookup at the hits of an orthographic render
        ps = float(rng.choice([0.5, 2.0]))
        o2 = ctx.render_opts(plane_scale=ps)
        _, ids = ctx.render(res, o2)
        feats = rng.normal(size=(V, int(rng.integers(1, 5)))).astype(np.float32)
        ok = ok and np.array_equal(ctx.interpolate_features(res, o2, ids, feats, zero=-3.0),
                                   orc.interpolate_features(uv, F, views, radius, res, ids, feats, plane_scale=ps, zero=-3.0))
        # ---- votes
        K = int(rng.integers(2, 9))
        labels = rng.integers(-1, K + 1, ids.shape).astype(np.int32)
        votes = ctx.backproject(ids, labels, K)
        ok = ok and np.array_equal(votes, orc.backproject(ids, labels, len(F), K))
        ok = ok and np.array_equal(ctx.votes_argmax(votes, default=-7), orc.votes_argmax(votes, default=-7))
        cl = rng.integers(0, 5, len(F)).astype(np.int32)
        ok = ok and np.array_equal(ctx.cells_to_points(cl, default=9), orc.cells_to_points(cl, F, V, default=9))
        # ---- label post-processing on random (salt-and-pepper over blobs) point labels
        lab = np.argmax(P @ rng.normal(size=(3, 3)).T, axis=1).astype(np.int32)
        flip = rng.random(V) < 0.15
        lab[flip] = rng.integers(-2, 5, flip.sum()).astype(np.int32)
        label = int(rng.integers(-2, 5))
        mc = int(rng.choice([1, 3, 20, 10 ** 6]))
        ok = ok and np.array_equal(ctx.labels_remove_islands(lab.copy(), label, mc, True), orc.pp_remove_islands(F, V, lab, label, mc, True))
        got = lab.copy()
        n = ctx.labels_connectivity(got, label, 7)
        ref, n_ref = orc.pp_connectivity(F, V, lab, label, 7)
        ok = ok and n == n_ref and np.array_equal(got, ref)
        it = int(rng.integers(1, 4))
        tgt = int(rng.integers(-2, 5))
        ok = ok and np.array_equal(ctx.labels_dilate(lab.copy(), label, it, False, None), orc.pp_dilate(F, V, lab, label, it))
        ok = ok and np.array_equal(ctx.labels_dilate(lab.copy(), label, it, True, tgt), orc.pp_dilate(F, V, lab, label, it, True, tgt))
        kw = [dict(), dict(iterations=it), dict(ignore_label=tgt), dict(target=tgt, iterations=it + 7)][int(rng.integers(0, 4))]
        ok = ok and np.array_equal(ctx.labels_erode(lab.copy(), label, **kw), orc.pp_erode(F, V, lab, label, **kw))
        if not ok:
            bad += 1
            print("MISMATCH case", case, "V", V, "T", len(F), "res", res, "angle", angle, cfg, enc, "label", label, "mc", mc, kw)
    print("fuzz (widened rows): %d cases, %d mismatching, %.0f s" % (n_cases, bad, time.time() - t0))
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    sys.exit(1 if run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
