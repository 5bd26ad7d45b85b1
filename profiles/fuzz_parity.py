"""Randomised parity of the render path against the oracle (brute force): meshes with slivers, degenerate and
coincident triangles, random viewpoints / radii / plane scales / spacings / resolutions / tolerances.
  python profiles/fuzz_parity.py [n_cases] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fly-by-cnn_b200"))
import numpy as np
from flyby_b200 import _cabi, synth
from oracle import oracle as orc

rng = None
ctx = None


def soup(n):
    """random triangles: some tiny, some huge, slivers, zero-area, shared vertices, exact duplicates"""
    V = rng.uniform(-1, 1, (n * 3, 3)).astype(np.float32)
    F = np.arange(n * 3, dtype=np.int32).reshape(n, 3)
    k = n // 5
    V[F[:k, 1]] = V[F[:k, 0]] + rng.normal(0, 1e-3, (k, 3)).astype(np.float32)        # slivers / tiny
    V[F[k:k + 3, 2]] = V[F[k:k + 3, 1]]                                                # zero area
    F[-3:] = F[:3]                                                                     # duplicate cells
    g = np.round(rng.uniform(-1, 1, (12, 3)) * 4) / 4                                  # lattice points: rays through vertices
    V[F[k + 3:k + 7].reshape(-1)] = g.astype(np.float32)
    return V, F


def mesh():
    kind = rng.integers(0, 5)
    if kind == 0:
        return synth.bumpy_sphere(int(rng.integers(3, 28)), int(rng.integers(0, 1000)))
    if kind == 1:
        return synth.dental_crown(int(rng.integers(4, 30)), int(rng.integers(0, 1000)))
    if kind == 2:
        return synth.cube()
    if kind == 3:
        P, F = synth.icosphere(int(rng.integers(2, 20)))
        return P.astype(np.float32), F
    return soup(int(rng.integers(8, 400)))


def run(n_cases, seed, context=None):
    """returns the number of mismatching cases"""
    global rng, ctx
    rng = np.random.default_rng(seed)
    ctx = context or _cabi.Context(0)
    bad = 0
    t0 = time.time()
    for case in range(n_cases):
        bad += 0 if one_case(case) else 1
    orc.set_tolerance(1e-10)
    print("fuzz: %d cases, %d mismatching, %.0f s" % (n_cases, bad, time.time() - t0))
    return bad


def one_case(case):
    P, F = mesh()
    skip = bool(rng.integers(0, 2))
    if skip:
        P = (P / max(1e-6, np.abs(P).max()) * rng.choice([0.01, 0.5, 1.0, 30.0])).astype(np.float32)
    ctx.set_mesh(P, F, skip_normalise=skip, consistency="off")  # soups are not orientable surfaces: cells as given
    ctx.build()
    uv, nrm, _, _ = ctx.get_mesh()
    pre_ok = np.array_equal(nrm, orc.point_normals(uv, F), equal_nan=True)      # vtkPolyDataNormals restatement
    if not skip:
        pre_ok = pre_ok and np.array_equal(uv, orc.unit_surf(P)[0])               # GetUnitSurf restatement
    ext = float(np.abs(uv).max())
    radius = float(ext * rng.choice([0.3, 1.2, 2.0, 4.0, 10.0]))
    nv = int(rng.integers(1, 6))
    if rng.integers(0, 2):
        pts = rng.normal(size=(nv, 3))
        pts[rng.integers(0, nv)] = [0, 0, rng.choice([-1.0, 1.0])]      # a pole now and then
        pts = (pts / np.linalg.norm(pts, axis=1, keepdims=True) * radius).astype(np.float32)
        radius = float(np.linalg.norm(pts[0].astype(np.float64)))
        ctx.views_custom(pts, radius)
    else:
        ctx.views_icosahedron(1, radius)
        nv = 12
    views = ctx.get_views()
    res = int(rng.choice([1, 2, 7, 16, 33, 64, 97, 160]))
    ps = float(ext * rng.choice([0.05, 0.5, 1.0, 2.5]))
    sp = float(rng.choice([0.5, 1.0, 1.0, 1.3]))
    tol = float(rng.choice([1e-10, 1e-10, 1e-8, 1e-5]))
    cfg = dict(use_z=int(rng.integers(0, 2)), split_z=int(rng.integers(0, 2)))
    mode = str(rng.choice(["raster", "bvh", "fused"]))     # the engine, through the ABI (flyby_render_opts.mode)
    opts = ctx.render_opts(plane_scale=ps, plane_spacing=sp, tolerance=tol, mode=mode, **cfg)
    # per-point scalars written by the render, in half of the cases (vertex 0 / barycentric)
    S = int(rng.choice([0, 0, 1, 3]))
    smode = str(rng.choice(["vertex0", "barycentric"]))
    scal = rng.normal(size=(len(P), S)).astype(np.float32) if S else None
    ctx.set_point_scalars(scal, mode=smode, zero=-1.5)
    feat, ids, t = ctx.render(res, opts, want_t=True)
    orc.set_tolerance(tol)
    fo, io, to = orc.render(uv, F, nrm, views, radius, res, plane_scale=ps, spacing=sp, grid=False, want_t=True, **cfg)
    Cn = fo.shape[-1]
    ok = pre_ok and np.array_equal(ids, io) and np.array_equal(t, to) and np.array_equal(feat[..., :Cn], fo, equal_nan=True)
    if S:
        want = orc.point_features(io, F, scal, zero=-1.5) if smode == "vertex0" else \
            orc.interpolate_features(uv, F, views, radius, res, io, scal, plane_scale=ps, spacing=sp, zero=-1.5)
        ok = ok and np.array_equal(feat[..., Cn:], want, equal_nan=True)
    ctx.set_point_scalars(None)
    # opt-in tolerance shading: ids bit-exact, values within the bar (finite meshes only; slivers fall back to exact)
    if mode == "raster" and np.isfinite(uv).all():
        tf, ti = ctx.render(res, ctx.render_opts(plane_scale=ps, plane_spacing=sp, tolerance=tol, shade_tolerance=True, **cfg))
        hit = io >= 0
        ok = ok and np.array_equal(ti, io) and np.array_equal(tf[~hit], fo[~hit])
        if hit.any():
            scale = np.maximum(np.abs(fo[hit]).max(), 1.0)
            ok = ok and bool(np.all(np.abs(tf[hit].astype(np.float64) - fo[hit]) <= 1e-5 * scale))
    # arbitrary segments through the same mesh (flyby_cast): the per-ray LBVH traversal with the same tolerance
    nseg = 2000
    a = rng.uniform(-1.5, 1.5, (nseg, 3)) * ext
    b = rng.uniform(-1.5, 1.5, (nseg, 3)) * ext
    b[:50] = a[:50]
    b[50:100, 0] = a[50:100, 0]
    rays = np.concatenate([a, b], axis=1)
    ci, ct, cw = ctx.cast(rays, tolerance=tol)
    oi, ot, ox, ow = orc.cast(uv, F, rays, grid=False)
    hit_c = oi >= 0
    ok = ok and np.array_equal(ci, oi) and np.array_equal(ct, ot) and np.array_equal(cw[hit_c], ow[hit_c])
    if not ok:
        print("MISMATCH case", case, "mode", mode, "S", S, smode, "T", len(F), "skip", skip, "R", radius, "res", res, "ps", ps, "sp", sp, "tol", tol,
              "ids", int((ids != io).sum()), "t", int((t != to).sum()))
    return ok


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    sys.exit(1 if run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
