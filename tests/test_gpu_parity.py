"""GPU parity tests: every call goes through the C-ABI (flyby_b200._cabi) and is
compared with the CPU oracle (oracle/flyby_oracle.c) on the same seeded inputs.

Bar: hit cell ids bit-exact (tie-break smallest t, then lowest cell id; any
disagreement is counted and fails); t / depth / normals within 1e-5 relative
(they are in fact expected to be bit-equal, which is reported).
"""
import numpy as np
import pytest

from flyby_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-5
ROOT_DIR = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))


def _build(ctx, P, F, **kw):
    ctx.set_mesh(P, F, **kw)
    ctx.build()
    return ctx.get_mesh()


def _oracle_mesh(orc, P, F, **kw):
    U, c, s = orc.unit_surf(P, **kw)
    return U, orc.point_normals(U, F), c, s


MESHES = {
    "bumpy12": lambda: synth.bumpy_sphere(12, 0),
    "bumpy40": lambda: synth.bumpy_sphere(40, 3),
    "crown30": lambda: synth.dental_crown(30, 1),
    "cube": lambda: synth.cube(),
}


@pytest.mark.parametrize("name", list(MESHES))
def test_unit_surf_and_normals(gpu_ctx, orc, name):
    P, F = MESHES[name]()
    uv, nrm, c, s = _build(gpu_ctx, P, F)
    U, N, co, so = _oracle_mesh(orc, P, F)
    assert np.array_equal(c, co) and s == so
    assert np.array_equal(uv, U)
    assert np.array_equal(nrm, N), "max diff %g" % np.abs(nrm - N).max()


@pytest.mark.parametrize("kw", [dict(centre_mode=1), dict(scale_mode=1), dict(centre_mode=1, scale_mode=1),
                                dict(translate=[1.0, -2.0, 0.5]), dict(scale_factor=0.03),
                                dict(translate=[0.5, 0.25, -1.0], scale_factor=0.02)])
def test_unit_surf_options(gpu_ctx, orc, kw):
    P, F = synth.bumpy_sphere(20, 4)
    uv, nrm, c, s = _build(gpu_ctx, P, F, **kw)
    U, c2, s2 = orc.unit_surf(P, **kw)
    assert np.allclose(c, c2, rtol=1e-14) and np.isclose(s, s2, rtol=1e-14)
    assert np.array_equal(uv, U)


def test_views_icosahedron(gpu_ctx, orc):
    for L, n in [(1, 12), (2, 42), (3, 92), (4, 162), (5, 252), (10, 1002)]:
        assert gpu_ctx.views_icosahedron(L, 2.75) == n
        assert np.array_equal(gpu_ctx.get_views(), orc.icosahedron(L, 2.75))
    # float32-equality emulation of the reference's point locator (hazard H1)
    for L in (2, 3, 5):
        n = gpu_ctx.views_icosahedron(L, 2.0, dedupe=1)
        ref = orc.icosahedron(L, 2.0, dedupe=1)
        assert n == len(ref) and np.array_equal(gpu_ctx.get_views(), ref)


def test_views_spiral(gpu_ctx, orc):
    for N, turns, R in [(64, 4, 4.0), (64, 4, 2.0), (17, 3, 1.5)]:
        assert gpu_ctx.views_spiral(N, turns, R) == N
        got, ref = gpu_ctx.get_views(), orc.spiral(N, turns, R)
        assert np.array_equal(got, ref)      # libm on the host, as the reference's math.sin / math.cos
        assert np.array_equal(got[0], np.array([0, 0, R], np.float32))


@pytest.mark.parametrize("ps,spacing", [(1.0, 1.0), (2.0, 1.0), (1.0, 0.75)])
def test_rays(gpu_ctx, orc, ps, spacing):
    gpu_ctx.views_icosahedron(2, 2.0)
    views = gpu_ctx.get_views()
    opts = gpu_ctx.render_opts(plane_scale=ps, plane_spacing=spacing)
    res = 33
    rays = gpu_ctx.generate_rays(res, opts)
    for v in range(len(views)):
        ref, _ = orc.view_rays(views[v], 2.0, res, ps, spacing)
        assert np.array_equal(rays[v].reshape(-1, 6), ref), "view %d" % v


@pytest.mark.parametrize("name", list(MESHES))
def test_bvh_invariants(gpu_ctx, name):
    """Walk the LBVH from the root: every triangle slot is covered by exactly one leaf (a leaf is a run
    of consecutive Morton-sorted slots), every stored child box is the exact bound of what hangs below
    it, parent / sibling links are consistent, the top block is a breadth-first prefix."""
    P, F = MESHES[name]()
    uv, _, _, _ = _build(gpu_ctx, P, F)
    nodes, n_top, slots = gpu_ctx.get_bvh()
    T = len(F)
    assert len(nodes) == T - 1 and sorted(slots.tolist()) == list(range(T))
    box = nodes[:, :12].view(np.float32)
    link = nodes[:, 12:].view(np.int32)
    tri = uv[F[slots]]  # [T,3,3] in slot order
    leaf_lo, leaf_hi = tri.min(1), tri.max(1)

    def leaf_range(c):
        v = ~int(c)
        return v & 0x07ffffff, (v >> 27) + 1

    covered = np.zeros(T, np.int32)
    seen = set()
    assert link[0, 2] == -1

    def bounds(i):  # returns (lo, hi) of node i, checking everything below it
        assert i not in seen
        seen.add(i)
        los, his = [], []
        for side in (0, 1):
            c = int(link[i, side])
            b_lo, b_hi = box[i, 6 * side:6 * side + 3], box[i, 6 * side + 3:6 * side + 6]
            if c < 0:
                first, cnt = leaf_range(c)
                assert 1 <= cnt <= 4 or cnt == 1
                covered[first:first + cnt] += 1
                lo, hi = leaf_lo[first:first + cnt].min(0), leaf_hi[first:first + cnt].max(0)
            else:
                assert link[c, 2] == i and link[c, 3] == link[i, 1 - side]
                lo, hi = bounds(c)
            assert np.array_equal(b_lo, lo) and np.array_equal(b_hi, hi), (i, side)
            los.append(lo)
            his.append(hi)
        return np.minimum(los[0], los[1]), np.maximum(his[0], his[1])

    import sys
    sys.setrecursionlimit(10000)
    bounds(0)
    assert np.all(covered == 1)
    # top block: a breadth-first prefix (every parent precedes its children)
    assert 1 <= n_top <= min(1024, T - 1)
    par = link[1:n_top, 2]
    assert np.all(par < np.arange(1, n_top))


CONFIGS = [
    dict(use_z=1, split_z=1),
    dict(use_z=1, split_z=0),
    dict(use_z=0, split_z=0),
    dict(use_z=1, split_z=1, normal_encoding="byte255"),
    dict(use_z=1, split_z=1, normal_encoding="signed", z_scale=0.25),
    dict(use_z=1, split_z=1, plane_scale=2.0),
    dict(use_z=1, split_z=1, plane_spacing=0.8),
]
ENC = {"unit01": 0, "byte255": 1, "signed": 2}


def _compare_render(ctx, orc, P, F, radius, res, cfg, views=None, grid_div=32):
    uv, nrm, _, _ = _build(ctx, P, F)
    if views is None:
        views = ctx.get_views()
    opts = ctx.render_opts(**cfg)
    feat, ids, t = ctx.render(res, opts, want_t=True)
    fo, io, to = orc.render(uv, F, nrm, views, radius, res, plane_scale=cfg.get("plane_scale", 1.0),
                            spacing=cfg.get("plane_spacing", 1.0), use_z=cfg["use_z"], split_z=cfg["split_z"],
                            normal_encoding=ENC[cfg.get("normal_encoding", "unit01")],
                            zscale=cfg.get("z_scale", 1.0), grid=True, grid_div=grid_div, want_t=True)
    mism = int((ids != io).sum())
    assert mism == 0, "%d / %d hit cell ids differ (grazing-edge disagreements are not tolerated)" % (mism, ids.size)
    hit = io >= 0
    assert np.array_equal(t[~hit], to[~hit])
    assert np.allclose(t[hit], to[hit], rtol=RTOL, atol=0)
    assert np.allclose(feat, fo, rtol=RTOL, atol=1e-7)
    return dict(rays=ids.size, hits=int(hit.sum()), t_bit_equal=bool(np.array_equal(t, to)),
                feat_bit_equal=bool(np.array_equal(feat, fo)))


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("name", ["bumpy12", "crown30"])
def test_render_parity(gpu_ctx, orc, name, cfg):
    P, F = MESHES[name]()
    gpu_ctx.views_icosahedron(2, 2.0)
    r = _compare_render(gpu_ctx, orc, P, F, 2.0, 96, cfg)
    assert r["hits"] > 0 and r["t_bit_equal"] and r["feat_bit_equal"]


def test_render_spiral_and_odd_resolution(gpu_ctx, orc):
    P, F = synth.dental_crown(40, 2)
    gpu_ctx.views_spiral(24, 4, 4.0)
    for res in (1, 7, 61, 224):
        r = _compare_render(gpu_ctx, orc, P, F, 4.0, res, dict(use_z=1, split_z=0))
    assert r["hits"] > 0


def test_render_view_ranges_concatenate(gpu_ctx, orc):
    """Sharding by view (config 5): slabs of a view range equal the full render bit for bit."""
    P, F = synth.bumpy_sphere(30, 7)
    _build(gpu_ctx, P, F)
    n = gpu_ctx.views_icosahedron(2, 2.0)
    opts = gpu_ctx.render_opts()
    full_f, full_i = gpu_ctx.render(64, opts)
    parts = [gpu_ctx.render(64, opts, view_begin=a, view_end=b) for a, b in [(0, 5), (5, 6), (6, 30), (30, n)]]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), full_f)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), full_i)


def test_cube_grazing_tie_break(gpu_ctx, orc):
    """Rays exactly through shared edges / vertices of a cube: both adjacent cells are
    inclusive hits; the documented tie-break (smallest t, then lowest cell id) decides."""
    P, F = synth.cube()
    uv, nrm, _, _ = _build(gpu_ctx, P, F, skip_normalise=True)
    assert np.array_equal(uv, P)
    g = np.linspace(-1.0, 1.0, 9)
    rays = []
    for axis in range(3):
        a, b = [k for k in range(3) if k != axis]
        for u in g:
            for v in g:
                o = np.zeros(3); e = np.zeros(3)
                o[a] = e[a] = u; o[b] = e[b] = v
                o[axis], e[axis] = 3.0, -3.0
                rays.append(np.concatenate([o, e]))
                rays.append(np.concatenate([e, o]))
    rays = np.array(rays)
    ids, t, w = gpu_ctx.cast(rays)
    io, to, xo, wo = orc.cast(uv, F, rays, grid=False)
    assert np.array_equal(ids, io) and np.all(ids >= 0)
    assert np.array_equal(t, to) and np.array_equal(w, wo)
    # a ray through the face diagonal is an inclusive hit of both triangles: lowest id wins
    diag = [i for i in range(len(rays)) if (np.sort(np.abs(rays[i][:3]))[:2] == 0).all()]
    assert len(diag) > 0


def test_cast_random_segments(gpu_ctx, orc):
    rng = np.random.default_rng(11)
    P, F = synth.bumpy_sphere(16, 5)
    uv, _, _, _ = _build(gpu_ctx, P, F)
    n = 20000
    a = rng.normal(size=(n, 3)); a /= np.linalg.norm(a, axis=1, keepdims=True)
    b = rng.normal(size=(n, 3)); b /= np.linalg.norm(b, axis=1, keepdims=True)
    rays = np.concatenate([a * rng.uniform(0.0, 3.0, (n, 1)), b * rng.uniform(0.0, 3.0, (n, 1))], axis=1)
    # axis-parallel and degenerate (zero-length) segments
    rays[:500, 3] = rays[:500, 0]; rays[:500, 4] = rays[:500, 1]
    rays[500:1000, 4] = rays[500:1000, 1]
    rays[1000:1010, 3:] = rays[1000:1010, :3]
    ids, t, w = gpu_ctx.cast(rays)
    io, to, xo, wo = orc.cast(uv, F, rays, grid=False)
    assert int((ids != io).sum()) == 0
    assert np.array_equal(t, to)
    hit = io >= 0
    assert hit.sum() > 1000 and np.array_equal(w[hit], wo[hit])


def test_single_triangle_and_empty(gpu_ctx, orc):
    V = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32)
    F = np.array([[0, 1, 2]], np.int32)
    gpu_ctx.views_custom(np.array([[0, 0, 2.0], [0, 0, -2.0], [2.0, 0, 0]], np.float32), 2.0)
    r = _compare_render(gpu_ctx, orc, V, F, 2.0, 32, dict(use_z=1, split_z=1, plane_scale=2.0),
                        views=gpu_ctx.get_views())
    assert r["hits"] > 0
    # a mesh without cells renders all-miss
    gpu_ctx.set_mesh(V, np.zeros((0, 3), np.int32))
    gpu_ctx.build()
    feat, ids = gpu_ctx.render(16, gpu_ctx.render_opts())
    assert np.all(ids == -1) and np.all(feat == 0)


def test_piled_candidates_spill_list(gpu_ctx, orc):
    """More candidates at a pixel than the per-pixel lists of the scan-conversion path hold: 40 coincident
    copies of a patch (every copy is an inclusive hit at the same t: lowest cell id wins), 40 near-coincident
    layers, and a fan of edge-on triangles seen exactly along their plane (grazing: the float32 screening
    cannot decide any of them)."""
    base = np.array([[-0.9, -0.8, 0.1], [0.9, -0.7, 0.1], [0.1, 0.9, 0.1], [-0.8, 0.7, 0.1]], np.float32)
    quad = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    V, F = [], []
    for k in range(40):  # coincident copies
        F.append(quad + len(V) * 4); V.append(base)
    for k in range(40):  # layers 1e-7 apart: inside the float32 depth margin of each other
        F.append(quad + len(V) * 4); V.append(base + np.array([0, 0, -0.3 - 1e-7 * k], np.float32))
    for k in range(24):  # edge-on fan in the plane x = 0.05, seen along +-y and +-z
        a = 0.2 * k
        F.append(np.array([[0, 1, 2]], np.int32) + len(V) * 4)
        V.append(np.array([[0.05, np.cos(a), np.sin(a)], [0.05, np.cos(a + 2), np.sin(a + 2)],
                           [0.05, np.cos(a + 4), np.sin(a + 4)], [0, 0, 0]], np.float32) * np.float32(0.6)
                 + np.array([0.05 * 0.4, 0, 0], np.float32))
    V = np.concatenate(V).astype(np.float32); F = np.concatenate(F).astype(np.int32)
    F = F[np.random.default_rng(3).permutation(len(F))]  # cell ids not in construction order
    views = np.array([[0, 0, 2.0], [0, 0, -2.0], [0, 2.0, 0], [0, -2.0, 0], [2.0, 0, 0], [1.2, 1.2, 1.0]], np.float32)
    gpu_ctx.views_custom(views, 2.0)
    uv, nrm, _, _ = _build(gpu_ctx, V, F, skip_normalise=True)
    opts = gpu_ctx.render_opts(use_z=1, split_z=1, plane_scale=2.0)
    for res in (33, 128):
        feat, ids, t = gpu_ctx.render(res, opts, want_t=True)
        fo, io, to = orc.render(uv, F, nrm, gpu_ctx.get_views(), 2.0, res, plane_scale=2.0, grid=False, want_t=True)
        assert int((ids != io).sum()) == 0 and np.array_equal(t, to) and np.array_equal(feat, fo)
    assert (io[0] >= 0).sum() > 1000 and (io[4] >= 0).sum() > 0 and (io[2] >= 0).sum() == 0  # edge-on: VTK rejects grazing


def test_context_reuse_equals_fresh_contexts():
    """A context keeps grow-only scratch (front buffer, candidate lists, spill list, queues) between renders:
    renders of changing sizes and meshes on ONE context (piled candidates that go through the spill list, then
    fewer pixels, then more pixels than ever before) must equal the same renders on fresh contexts."""
    from flyby_b200 import _cabi
    base = np.array([[-0.9, -0.8, 0.1], [0.9, -0.7, 0.1], [0.1, 0.9, 0.1], [-0.8, 0.7, 0.1]], np.float32)
    Vp = np.concatenate([base + np.array([0, 0, -1e-7 * k], np.float32) for k in range(30)]).astype(np.float32)
    Fp = np.concatenate([np.array([[0, 1, 2], [0, 2, 3]], np.int32) + 4 * k for k in range(30)])
    Pb, Fb = synth.bumpy_sphere(24, 2)
    Pc, Fc = synth.cube()
    jobs = [("piled", Vp, Fp, 96, 2.0), ("bumpy", Pb, Fb, 40, 1.0), ("cube", Pc, Fc, 64, 4.0), ("piled", Vp, Fp, 33, 2.0),
            ("bumpy", Pb, Fb, 150, 2.0), ("piled", Vp, Fp, 96, 2.0), ("bumpy", Pb, Fb, 40, 1.0)]

    def run(ctx, job):
        _, P, F, res, ps = job
        ctx.set_mesh(P, F)
        ctx.build()
        ctx.views_icosahedron(1, 2.0)
        return ctx.render(res, ctx.render_opts(use_z=1, split_z=1, plane_scale=ps), want_t=True)

    one = _cabi.Context(0)
    got = [run(one, j) for j in jobs]
    one.close()
    for j, g in zip(jobs, got):
        fresh = _cabi.Context(0)
        want = run(fresh, j)
        fresh.close()
        for a, b in zip(g, want):
            assert np.array_equal(a, b), j[0]
    assert (got[0][1] >= 0).sum() > 1000 and (got[4][1] >= 0).sum() > 1000


def test_bad_indices_rejected(gpu_ctx):
    from flyby_b200 import FlyByError
    V = np.zeros((3, 3), np.float32)
    gpu_ctx.set_mesh(V, np.array([[0, 1, 3]], np.int32))
    with pytest.raises(FlyByError):
        gpu_ctx.build()


def test_point_features_and_backprojection(gpu_ctx, orc):
    rng = np.random.default_rng(3)
    P, F = synth.bumpy_sphere(24, 9)
    uv, _, _, _ = _build(gpu_ctx, P, F)
    gpu_ctx.views_icosahedron(2, 2.0)
    _, ids = gpu_ctx.render(128, gpu_ctx.render_opts())
    feats = rng.normal(size=(len(P), 3)).astype(np.float32)
    out = gpu_ctx.point_features(ids, feats, zero=-7.0)
    assert np.array_equal(out, orc.point_features(ids, F, feats, zero=-7.0))
    coords = gpu_ctx.point_features(ids, uv, zero=0.0)  # --point_features coords
    assert np.array_equal(coords, orc.point_features(ids, F, uv, zero=0.0))
    K = 34
    labels = rng.integers(0, K, size=ids.shape).astype(np.int32)
    labels[0, :4, :4] = -1  # invalid labels are ignored
    votes = gpu_ctx.backproject(ids, labels, K)
    ref = orc.backproject(ids, labels, len(F), K)
    assert np.array_equal(votes, ref) and votes.sum() > 0
    # accumulation over two calls == one call (determinism of the integer atomics)
    half = ids.shape[0] // 2
    v2 = gpu_ctx.backproject(ids[:half], labels[:half], K)
    v2 = gpu_ctx.backproject(ids[half:], labels[half:], K, votes=v2)
    assert np.array_equal(v2, ref)
    cell_labels = gpu_ctx.votes_argmax(votes, default=33)
    assert np.array_equal(cell_labels, orc.votes_argmax(ref, default=33))
    pts = gpu_ctx.cells_to_points(cell_labels, default=33)
    assert np.array_equal(pts, orc.cells_to_points(cell_labels, F, len(P), default=33))
    # soft votes: fixed point 2^-20
    probs = rng.random(size=ids.shape + (4,)).astype(np.float32)
    vfx = gpu_ctx.backproject_soft(ids, probs, 4)
    ref_fx = np.zeros((len(F), 4), np.int64)
    hit = ids >= 0
    np.add.at(ref_fx, ids[hit], np.rint(probs[hit].astype(np.float32) * np.float32(1048576.0)).astype(np.int64))
    assert np.array_equal(vfx, ref_fx)


def test_torch_device_buffers(gpu_ctx, orc):
    """Device-resident inputs and outputs (the zero-copy path bench.py times)."""
    torch = pytest.importorskip("torch")
    P, F = synth.bumpy_sphere(24, 2)
    dev = torch.device("cuda:0")
    tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
    torch.cuda.synchronize()  # uploads ran on torch's stream, the context works on its own
    gpu_ctx.set_mesh(tp, tf)
    gpu_ctx.build()
    n = gpu_ctx.views_icosahedron(1, 2.0)
    opts = gpu_ctx.render_opts()
    feat = torch.empty((n, 80, 80, 4), dtype=torch.float32, device=dev)
    ids = torch.empty((n, 80, 80), dtype=torch.int32, device=dev)
    gpu_ctx.render_into(80, opts, feat, ids)
    gpu_ctx.synchronize()
    f2, i2 = gpu_ctx.render(80, opts)
    assert np.array_equal(ids.cpu().numpy(), i2) and np.array_equal(feat.cpu().numpy(), f2)
    labels = (ids % 7).to(torch.int32)
    votes = torch.zeros((len(F), 7), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()  # labels / votes were produced on torch's stream
    gpu_ctx.backproject(ids, labels, 7, votes=votes)
    gpu_ctx.synchronize()
    assert np.array_equal(votes.cpu().numpy(), orc.backproject(i2, labels.cpu().numpy(), len(F), 7))


def test_config1_full_size(gpu_ctx, orc):
    """BASELINE config 1 at full size: 100 820 triangles, icosahedron subdivision 2 (42 views),
    resolution 512, radius 2, use_z 1, split_z 1 -- every one of the 11 M rays against the oracle."""
    P, F = synth.bumpy_sphere(71, 0)
    gpu_ctx.views_icosahedron(2, 2.0)
    r = _compare_render(gpu_ctx, orc, P, F, 2.0, 512, dict(use_z=1, split_z=1), grid_div=64)
    assert r["rays"] == 42 * 512 * 512 and r["hits"] > r["rays"] // 2
    assert r["t_bit_equal"] and r["feat_bit_equal"]
    st = gpu_ctx.stats()
    assert st["n_hits"] == r["hits"]


def test_tiny_radius_and_lost_pairs_fallback(orc, tmp_path):
    """Rarely taken paths, in a fresh process because two of them are switched by environment variables read
    once: (1) a view radius that is tiny against the mesh (view planes inside the surface), (2) a spill list too
    small for the piled candidates, so that pixels are re-traced through the LBVH (raster_resolve3_kernel),
    (3) a queue of big triangles that is full, so that raster_kernel rasterises them inline."""
    import os
    import subprocess
    import sys
    script = tmp_path / "lost.py"
    script.write_text('''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from flyby_b200 import _cabi, synth
from oracle import oracle as orc
ctx = _cabi.Context(0)
# (2) 30 coincident copies of a patch: ~29 extras per covered pixel, far more pairs than FLYBY_PAIR_CAP=64 holds
base = np.array([[-0.9, -0.8, 0.1], [0.9, -0.7, 0.1], [0.1, 0.9, 0.1], [-0.8, 0.7, 0.1]], np.float32)
V = np.concatenate([base] * 30); F = np.concatenate([np.array([[0, 1, 2], [0, 2, 3]], np.int32) + 4 * k for k in range(30)])
ctx.set_mesh(V, F, skip_normalise=True); ctx.build()
ctx.views_custom(np.array([[0, 0, 2.0], [1.0, 1.0, 1.4]], np.float32), 2.0)
opts = ctx.render_opts(plane_scale=2.0)
feat, ids, t = ctx.render(48, opts, want_t=True)
ctx.stats()  # with FLYBY_DEBUG_STATS set this prints the counters of the render to stderr
N = orc.point_normals(V, F)
fo, io, to = orc.render(V, F, N, ctx.get_views(), 2.0, 48, plane_scale=2.0, grid=False, want_t=True)
assert (ids != io).sum() == 0 and np.array_equal(t, to) and np.array_equal(feat, fo) and (io >= 0).sum() > 500
assert set(np.unique(io[io >= 0])) <= {0, 1}
# (3) a queue of big triangles too small for a cube seen from twelve sides (FLYBY_BIG_CAP=3): the rest is
# rasterised inline by the warp that owns the triangle
P, F3 = synth.cube()
ctx.set_mesh(P, F3, skip_normalise=True); ctx.build()
ctx.views_icosahedron(1, 3.0)
opts = ctx.render_opts(plane_scale=4.0)
feat, ids, t = ctx.render(96, opts, want_t=True)
N = orc.point_normals(P, F3)
fo, io, to = orc.render(P, F3, N, ctx.get_views(), 3.0, 96, plane_scale=4.0, grid=False, want_t=True)
assert (ids != io).sum() == 0 and np.array_equal(t, to) and np.array_equal(feat, fo) and (io >= 0).sum() > 10000
# (1) radius 0.02 around a unit mesh
P, F2 = synth.bumpy_sphere(16, 3)
ctx.set_mesh(P, F2); ctx.build()
uv, nrm, _, _ = ctx.get_mesh()
ctx.views_icosahedron(1, 0.02)
opts = ctx.render_opts(plane_scale=3.0)
feat, ids, t = ctx.render(40, opts, want_t=True)
fo, io, to = orc.render(uv, F2, nrm, ctx.get_views(), 0.02, 40, plane_scale=3.0, grid=False, want_t=True)
assert (ids != io).sum() == 0 and np.array_equal(t, to) and np.array_equal(feat, fo)
print("ok", int((io >= 0).sum()))
''' % (ROOT_DIR, os.path.join(ROOT_DIR, "fly-by-cnn_b200")))
    env = dict(os.environ, FLYBY_PAIR_CAP="64", FLYBY_BIG_CAP="3", FLYBY_DEBUG_STATS="1")
    env.pop("FLYBY_MODE", None)  # this test is about the default (scan-conversion) path
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout + out.stderr
    first = [ln for ln in out.stderr.splitlines() if "re-traced exactly" in ln][0]
    assert "re-traced exactly 0;" not in first, first  # the LBVH fallback really ran


def test_non_finite_vertices_are_never_hit(gpu_ctx, orc):
    """A NaN coordinate poisons only the triangles that use it (the exact test rejects them, as the oracle does)."""
    P, F = synth.bumpy_sphere(10, 6)
    U, _, _ = orc.unit_surf(P)
    U = U.copy()
    U[[5, 77, 300]] = np.nan
    uv, nrm, _, _ = _build(gpu_ctx, U, F, skip_normalise=True)
    gpu_ctx.views_icosahedron(1, 2.0)
    opts = gpu_ctx.render_opts(plane_scale=2.0)
    feat, ids, t = gpu_ctx.render(64, opts, want_t=True)
    N = orc.point_normals(U, F)
    fo, io, to = orc.render(U, F, N, gpu_ctx.get_views(), 2.0, 64, plane_scale=2.0, grid=False, want_t=True)
    assert int((ids != io).sum()) == 0 and np.array_equal(t, to)
    bad = np.isin(F, [5, 77, 300]).any(axis=1)
    assert not np.isin(io[io >= 0], np.nonzero(bad)[0]).any() and (io >= 0).sum() > 1000
    hit = io >= 0
    assert np.array_equal(feat[hit], fo[hit], equal_nan=True)  # points next to a poisoned triangle get NaN normals on both sides


def test_independent_algorithms_agree_at_bench_size(gpu_ctx, tmp_path):
    """Size-independent cross-check at the full bench workload (24.1 M rays, 200 000 triangles): the scan-conversion
    pipeline and the LBVH packet traversal share only the exact triangle test; their hit ids, hit parameters and
    feature tensors must have identical digests (the LBVH mode runs in a fresh process: the mode is read once)."""
    import hashlib
    import os
    import subprocess
    import sys
    P, F = synth.bumpy_sphere(100, 0)
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    gpu_ctx.views_icosahedron(3, 2.0)
    feat, ids, t = gpu_ctx.render(512, gpu_ctx.render_opts(use_z=1, split_z=1), want_t=True)
    mine = [hashlib.sha1(a.tobytes()).hexdigest() for a in (ids, t, feat)]
    assert (ids >= 0).sum() > 20_000_000
    script = tmp_path / "bvh_digest.py"
    script.write_text('''
import sys, hashlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
from flyby_b200 import _cabi, synth
ctx = _cabi.Context(0)
P, F = synth.bumpy_sphere(100, 0)
ctx.set_mesh(P, F); ctx.build(); ctx.views_icosahedron(3, 2.0)
feat, ids, t = ctx.render(512, ctx.render_opts(use_z=1, split_z=1), want_t=True)
print(" ".join(hashlib.sha1(a.tobytes()).hexdigest() for a in (ids, t, feat)))
''' % (ROOT_DIR, os.path.join(ROOT_DIR, "fly-by-cnn_b200")))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, FLYBY_MODE="bvh"))
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == mine


_C5_DIGEST = '''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from flyby_b200 import _cabi, synth
dev = torch.device("cuda", 0)
ctx = _cabi.Context(0)
P, F = synth.bumpy_sphere(316, 5)                      # 1 997 120 triangles
tp, tf = torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)
torch.cuda.synchronize()
ctx.set_mesh(tp, tf); ctx.build()
n = ctx.views_icosahedron(4, 2.0)                      # 162 views
res = 1024
feat = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
t = torch.empty((n, res, res), dtype=torch.float64, device=dev)
ctx.render_into(res, ctx.render_opts(use_z=1, split_z=1), feat, ids, t)
ctx.synchronize()
def digest(x):                                         # position-weighted sum of the raw bits, wrapping in int64
    b = x.reshape(-1).view(torch.int64 if x.element_size() == 8 else torch.int32).to(torch.int64)
    w = (torch.arange(b.numel(), device=dev, dtype=torch.int64) %% 1000003) + 1
    return int((b * w).sum())
print(int((ids >= 0).sum()), digest(ids), digest(t), digest(feat[..., :2].contiguous()), digest(feat[..., 2:].contiguous()))
'''


def test_config5_full_size_cross_check(tmp_path):
    """BASELINE config 5 in full (2 M triangles, 162 views at 1024^2, 170 M rays): the two independent searches
    agree on every id, hit parameter and feature (position-weighted digests computed on the device)."""
    import os
    import subprocess
    import sys
    script = tmp_path / "c5.py"
    script.write_text(_C5_DIGEST % (ROOT_DIR, os.path.join(ROOT_DIR, "fly-by-cnn_b200")))
    outs = []
    for mode in ("raster", "bvh"):
        env = dict(os.environ)
        env.pop("FLYBY_MODE", None)
        if mode != "raster":
            env["FLYBY_MODE"] = mode
        out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(out.stdout.split())
    assert outs[0] == outs[1] and int(outs[0][0]) > 100_000_000


def test_randomised_parity(gpu_ctx, orc):
    """A short run of profiles/fuzz_parity.py: random meshes (slivers, zero-area and duplicate cells, lattice
    vertices), viewpoints, radii from inside the mesh to far away, plane scales, spacings, resolutions, tolerances
    up to 1e-5, plus random segments through flyby_cast -- everything bit-equal to the brute-force oracle."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT_DIR, "profiles", "fuzz_parity.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    assert fuzz.run(60, 11, context=gpu_ctx) == 0


# ---------------------------------------------------------------- round 2: engines through the ABI, scalars, winding
@pytest.mark.parametrize("mode", ["raster", "bvh", "fused"])
def test_render_engines_selected_through_the_abi(gpu_ctx, orc, mode):
    """flyby_render_opts.mode picks the search engine in-process; all three equal the oracle bit for bit
    (reference: the ray loop of fly_by_features.cxx:742-807).  The fused engine also runs instrumented
    (count_work) so that nodes_visited / tris_tested are exercised."""
    P, F = synth.bumpy_sphere(30, 5)
    gpu_ctx.views_icosahedron(2, 2.0)
    for cfg in (dict(use_z=1, split_z=1), dict(use_z=1, split_z=0, plane_scale=2.0)):
        r = _compare_render(gpu_ctx, orc, P, F, 2.0, 96, dict(cfg, mode=mode))
        assert r["hits"] > 0 and r["t_bit_equal"] and r["feat_bit_equal"]
    st = gpu_ctx.stats()
    assert st["ms_trace"] > 0 and st["ms_render"] >= st["ms_trace"]
    if mode != "raster":
        gpu_ctx.set_count_work(True)
        try:
            opts = gpu_ctx.render_opts(mode=mode)
            f1, i1 = gpu_ctx.render(64, opts)
            st = gpu_ctx.stats()
            assert st["nodes_visited"] > 0 and st["tris_tested"] >= st["n_hits"] > 0
        finally:
            gpu_ctx.set_count_work(False)
        f0, i0 = gpu_ctx.render(64, gpu_ctx.render_opts(mode="raster"))
        assert np.array_equal(i0, i1) and np.array_equal(f0, f1)


def test_unknown_mode_is_rejected(gpu_ctx):
    from flyby_b200 import FlyByError
    P, F = synth.bumpy_sphere(6, 1)
    _build(gpu_ctx, P, F)
    gpu_ctx.views_icosahedron(1, 2.0)
    with pytest.raises(FlyByError):
        gpu_ctx.render(8, gpu_ctx.render_opts(mode=9))


@pytest.mark.parametrize("mode", ["raster", "bvh", "fused"])
@pytest.mark.parametrize("smode", ["vertex0", "barycentric"])
def test_point_scalars_written_by_the_render(gpu_ctx, orc, mode, smode):
    """--point_features ... --point_features_concat 1 (src/py/fly_by_features.py:334-402) in one pass: the S scalar
    channels behind the C image channels equal the reference's id-map lookup (vertex 0; utils.py:604-684) /
    the C++ tool's barycentric interpolation (fly_by_features.cxx:766-793), background = zero."""
    rng = np.random.default_rng(8)
    P, F = synth.dental_crown(24, 3)
    uv, nrm, _, _ = _build(gpu_ctx, P, F)
    views = gpu_ctx.views_icosahedron(1, 2.0) and gpu_ctx.get_views()
    for S, cfg in ((1, dict(use_z=1, split_z=1)), (3, dict(use_z=1, split_z=0)), (5, dict(use_z=0, split_z=0, plane_scale=2.0)),
                   (1, dict(use_z=1, split_z=0))):   # 3 + 1: a pixel stride of 4 that is NOT four image channels
        feats = rng.normal(size=(len(P), S)).astype(np.float32)
        gpu_ctx.set_point_scalars(feats, mode=smode, zero=-3.5)
        opts = gpu_ctx.render_opts(mode=mode, **cfg)
        C = gpu_ctx.channels(opts)
        out, ids = gpu_ctx.render(48, opts)
        assert out.shape[-1] == C + S
        fo, io = orc.render(uv, F, nrm, views, 2.0, 48, plane_scale=cfg.get("plane_scale", 1.0), use_z=cfg["use_z"],
                            split_z=cfg["split_z"])
        assert np.array_equal(ids, io) and np.array_equal(out[..., :C], fo)
        if smode == "vertex0":
            want = orc.point_features(io, F, feats, zero=-3.5)
        else:
            want = orc.interpolate_features(uv, F, views, 2.0, 48, io, feats, plane_scale=cfg.get("plane_scale", 1.0), zero=-3.5)
        assert np.array_equal(out[..., C:], want)
        assert (io < 0).any() and (io >= 0).any()
    # view-range slabs and removal
    part, _ = gpu_ctx.render(48, opts, view_begin=3, view_end=7)
    assert np.array_equal(part, out[3:7])
    gpu_ctx.set_point_scalars(None)
    plain, _ = gpu_ctx.render(48, opts)
    assert plain.shape[-1] == C and np.array_equal(plain, out[..., :C])


def _scramble_winding(F, seed, frac=0.3):
    rng = np.random.default_rng(seed)
    G = F.copy()
    flip = rng.random(len(G)) < frac
    G[flip] = G[flip][:, ::-1]
    return np.ascontiguousarray(G), flip


def test_winding_consistency_pass(gpu_ctx, orc):
    """vtkPolyDataNormals' consistency pass (utils.py:493-501, fly_by_features.cxx:472-478): a mesh with reversed
    cells is refused by default, reoriented like the oracle's restatement of VTK's wave traversal with
    consistency="fix" (every component follows its lowest cell id), used as given with "off"."""
    from flyby_b200 import FlyByError, _cabi
    P, F = synth.bumpy_sphere(14, 6)
    # two components: a second, shifted copy whose FIRST cell is reversed (the whole copy must follow it)
    P2 = np.concatenate([P, P + np.float32(3.0)]).astype(np.float32)
    F2 = np.concatenate([F, F[:, ::-1] + len(P)]).astype(np.int32)
    G, flip = _scramble_winding(F2, 4)
    assert flip.any() and not flip.all()
    gpu_ctx.set_mesh(P2, G)                                   # default: check
    with pytest.raises(FlyByError) as e:
        gpu_ctx.build()
    assert e.value.status == _cabi.ERR_UNSUPPORTED and "consisten" in str(e.value)
    want_tris, want_flip, n = orc.orient_cells(G, len(P2))
    assert n > 0
    gpu_ctx.set_mesh(P2, G, consistency="fix")
    gpu_ctx.build()
    n_gpu, got_flip = gpu_ctx.get_winding()
    assert n_gpu == n and np.array_equal(got_flip, want_flip)
    uv, nrm, _, _ = gpu_ctx.get_mesh()
    U, _, _ = orc.unit_surf(P2)
    assert np.array_equal(uv, U) and np.array_equal(nrm, orc.point_normals(U, want_tris))
    gpu_ctx.views_icosahedron(1, 2.0)
    views = gpu_ctx.get_views()
    feats = np.random.default_rng(2).normal(size=(len(P2), 2)).astype(np.float32)
    gpu_ctx.set_point_scalars(feats, mode="vertex0", zero=9.0)
    opts = gpu_ctx.render_opts(plane_scale=2.0)
    out, ids, t = gpu_ctx.render(64, opts, want_t=True)
    fo, io, to = orc.render(U, want_tris, nrm, views, 2.0, 64, plane_scale=2.0, want_t=True)
    assert np.array_equal(ids, io) and np.array_equal(t, to) and np.array_equal(out[..., :4], fo) and (io >= 0).any()
    # "vertex 0" stays the caller's vertex 0 (the reference builds its point-id map from the original surface)
    assert np.array_equal(out[..., 4:], orc.point_features(io, G, feats, zero=9.0))
    assert np.array_equal(gpu_ctx.point_ids(ids), np.where(io >= 0, G[np.maximum(io, 0), 0], -1))
    lab = (np.arange(len(G)) % 5).astype(np.int32)
    assert np.array_equal(gpu_ctx.cells_to_points(lab, default=-1), orc.cells_to_points(lab, G, len(P2), default=-1))
    gpu_ctx.set_point_scalars(None)
    # a consistent mesh: nothing flips in any mode, and "fix" equals "check"
    for mode in ("check", "fix", "off"):
        gpu_ctx.set_mesh(P, F, consistency=mode)
        gpu_ctx.build()
        assert gpu_ctx.get_winding()[0] == 0
    # off: the scrambled cells are used as given
    gpu_ctx.set_mesh(P2, G, consistency="off")
    gpu_ctx.build()
    assert np.array_equal(gpu_ctx.get_mesh()[1], orc.point_normals(U, G))
    # a Moebius band has no consistent orientation: refused in "fix" mode, like the oracle (-1)
    n_seg = 24
    a = np.linspace(0, 2 * np.pi, n_seg, endpoint=False)
    ring = []
    for s in (-0.3, 0.3):
        ring.append(np.stack([(1 + s * np.cos(a / 2)) * np.cos(a), (1 + s * np.cos(a / 2)) * np.sin(a), s * np.sin(a / 2)], 1))
    Pm = np.concatenate(ring).astype(np.float32)
    Fm = []
    for i in range(n_seg):
        j = (i + 1) % n_seg
        lo_i, hi_i = i, n_seg + i
        lo_j, hi_j = (j, n_seg + j) if j else (n_seg, 0)      # the half twist swaps the rims at the seam
        Fm += [[lo_i, lo_j, hi_i], [lo_j, hi_j, hi_i]]
    Fm = np.asarray(Fm, np.int32)
    assert orc.orient_cells(Fm, len(Pm))[2] == -1
    gpu_ctx.set_mesh(Pm, Fm, consistency="fix")
    with pytest.raises(FlyByError) as e:
        gpu_ctx.build()
    assert e.value.status == _cabi.ERR_UNSUPPORTED


def test_entry_points_validate_indices_without_a_build(gpu_ctx):
    """cells_to_points / point_features / point_ids / backproject_points read the triangles after
    flyby_set_mesh alone -- out-of-range indices must be refused there too, and ids beyond T are misses."""
    from flyby_b200 import FlyByError
    V = np.zeros((3, 3), np.float32)
    bad = np.array([[0, 1, 2], [7, 1, 2]], np.int32)
    ids = np.array([0, 1, -1, 5], np.int32)
    for call in (lambda: gpu_ctx.cells_to_points(np.zeros(2, np.int32)),
                 lambda: gpu_ctx.point_features(ids, np.zeros((3, 1), np.float32)),
                 lambda: gpu_ctx.point_ids(ids),
                 lambda: gpu_ctx.backproject_points(ids, np.zeros(4, np.int32), 2)):
        gpu_ctx.set_mesh(V, bad)
        with pytest.raises(FlyByError):
            call()
    good = np.array([[0, 1, 2], [2, 1, 0]], np.int32)
    gpu_ctx.set_mesh(V, good)
    assert np.array_equal(gpu_ctx.point_ids(ids), np.array([0, 2, -1, -1], np.int32))
    # mode 1 (decoded point ids) never reads the triangles: usable even with bad cells
    gpu_ctx.set_mesh(V, bad)
    f = np.arange(3, dtype=np.float32).reshape(3, 1)
    assert np.array_equal(gpu_ctx.point_features(np.array([2, 0, 9, -1], np.int32), f, zero=-1.0, mode=1).reshape(-1),
                          np.array([2, 0, -1, -1], np.float32))


def test_interpolate_ignores_foreign_cell_ids(gpu_ctx, orc):
    """a cell id >= T (ids from another mesh) is a miss, never a gather."""
    P, F = synth.bumpy_sphere(8, 1)
    _build(gpu_ctx, P, F)
    gpu_ctx.views_icosahedron(1, 2.0)
    opts = gpu_ctx.render_opts()
    _, ids = gpu_ctx.render(16, opts)
    feats = np.ones((len(P), 2), np.float32)
    want = gpu_ctx.interpolate_features(16, opts, ids, feats, zero=-1.0)
    spoiled = ids.copy()
    spoiled[0, :4] = len(F) + np.arange(4 * 16).reshape(4, 16) * 1000 + 1
    got = gpu_ctx.interpolate_features(16, opts, spoiled, feats, zero=-1.0)
    assert np.all(got[0, :4] == -1.0) and np.array_equal(got[0, 4:], want[0, 4:]) and np.array_equal(got[1:], want[1:])


@pytest.mark.parametrize("cfg", [dict(use_z=1, split_z=1), dict(use_z=1, split_z=0, plane_scale=2.0), dict(use_z=0, split_z=0),
                                 dict(use_z=1, split_z=1, normal_encoding="byte255"),
                                 dict(use_z=1, split_z=1, normal_encoding="signed", z_scale=0.25)])
def test_tolerance_shading_keeps_ids_exact(gpu_ctx, orc, cfg):
    """FLYBY_SHADE_TOLERANCE (opt-in): hit cell ids bit-exact, depth and normals within the task's 1e-5 relative of the
    exact chain (bar written here: depth 1e-5 relative, normal channels 1e-5 of their range; measured ~1e-7)."""
    P, F = synth.bumpy_sphere(40, 3)
    uv, nrm, _, _ = _build(gpu_ctx, P, F)
    gpu_ctx.views_icosahedron(2, 2.0)
    views = gpu_ctx.get_views()
    exact_f, exact_i = gpu_ctx.render(160, gpu_ctx.render_opts(**cfg))
    fast_f, fast_i = gpu_ctx.render(160, gpu_ctx.render_opts(shade_tolerance=True, **cfg))
    fo, io = orc.render(uv, F, nrm, views, 2.0, 160, plane_scale=cfg.get("plane_scale", 1.0), use_z=cfg["use_z"],
                        split_z=cfg["split_z"], normal_encoding=ENC[cfg.get("normal_encoding", "unit01")],
                        zscale=cfg.get("z_scale", 1.0))
    assert np.array_equal(exact_i, io) and np.array_equal(exact_f, fo)
    assert np.array_equal(fast_i, io), "%d ids differ in tolerance mode" % int((fast_i != io).sum())
    rng_ = 255.0 if cfg.get("normal_encoding") == "byte255" else 1.0
    hit = io >= 0
    assert np.array_equal(fast_f[~hit], fo[~hit])
    if cfg["use_z"] and cfg["split_z"]:
        d_ref = fo[..., 3][hit].astype(np.float64)
        assert np.max(np.abs(fast_f[..., 3][hit] - d_ref) / d_ref) <= 1e-5
        assert np.max(np.abs(fast_f[..., :3][hit].astype(np.float64) - fo[..., :3][hit])) <= 1e-5 * rng_
    elif not cfg["use_z"]:
        assert np.max(np.abs(fast_f[hit].astype(np.float64) - fo[hit])) <= 1e-5 * rng_
    else:   # channels multiplied by the depth: relative to the depth factor
        _, _, t = gpu_ctx.render(160, gpu_ctx.render_opts(**cfg), want_t=True)
        z = (t[hit] * 4.0 * cfg.get("z_scale", 1.0))[:, None]
        assert np.max(np.abs(fast_f[hit].astype(np.float64) - fo[hit]) / (z * rng_)) <= 1e-5
    # reproducible: every hit pixel is the same function of (pixel, winning triangle) whichever kernel finishes it,
    # so repeated renders and view-range slabs give identical bits although the atomics land in another order
    for _ in range(3):
        again_f, again_i = gpu_ctx.render(160, gpu_ctx.render_opts(shade_tolerance=True, **cfg))
        assert np.array_equal(again_f, fast_f) and np.array_equal(again_i, fast_i)
    part_f, _ = gpu_ctx.render(160, gpu_ctx.render_opts(shade_tolerance=True, **cfg), view_begin=5, view_end=9)
    assert np.array_equal(part_f, fast_f[5:9])
    # an exact hit parameter request falls back to the exact chain
    f3, i3, t3 = gpu_ctx.render(160, gpu_ctx.render_opts(shade_tolerance=True, **cfg), want_t=True)
    assert np.array_equal(f3, fo) and np.array_equal(i3, io)


def test_deferred_build_reports_at_the_next_synchronising_call(gpu_ctx, orc):
    """FLYBY_BUILD_DEFERRED: flyby_build does not wait for the device (streams of meshes); the index / winding checks
    still run, out-of-range indices are neutralised on the device and the verdict arrives with the next
    synchronising call.  Results of a good mesh are bit-identical to a synchronous build."""
    torch = pytest.importorskip("torch")
    from flyby_b200 import FlyByError, _cabi
    P, F = synth.bumpy_sphere(20, 4)
    gpu_ctx.views_icosahedron(1, 2.0)
    opts = gpu_ctx.render_opts()
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    want_f, want_i = gpu_ctx.render(48, opts)
    gpu_ctx.set_mesh(P, F, deferred=True)
    gpu_ctx.build()
    got_f, got_i = gpu_ctx.render(48, opts)
    assert np.array_equal(got_i, want_i) and np.array_equal(got_f, want_f)
    # out-of-range indices: nothing faults, the error surfaces at synchronize()
    bad = F.copy()
    bad[5, 1] = len(P) + 17
    bad[9, 0] = -3
    dev = torch.device("cuda:0")
    ids = torch.empty((12, 48, 48), dtype=torch.int32, device=dev)
    gpu_ctx.set_mesh(P, bad, deferred=True)
    gpu_ctx.build()
    gpu_ctx.render_into(48, opts, None, ids)      # device output: no synchronisation yet
    with pytest.raises(FlyByError) as e:
        gpu_ctx.synchronize()
    assert e.value.status == _cabi.ERR_INVALID and "2 triangle indices" in str(e.value)
    # reversed cells are flagged the same way
    G = F.copy()
    G[::7] = G[::7][:, ::-1]
    gpu_ctx.set_mesh(P, np.ascontiguousarray(G), deferred=True)
    gpu_ctx.build()
    with pytest.raises(FlyByError) as e:
        gpu_ctx.render(16, opts)                   # host outputs synchronise
    assert e.value.status == _cabi.ERR_UNSUPPORTED
    # and the context is usable afterwards
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    again_f, again_i = gpu_ctx.render(48, opts)
    assert np.array_equal(again_i, want_i) and np.array_equal(again_f, want_f)


def test_graph_replay_of_a_step(gpu_ctx, orc):
    """flyby_capture_begin / _end / flyby_graph_launch: set_mesh + build + render of a fixed-shape mesh recorded once and
    replayed with one launch; a replay re-reads the device buffers the captured set_mesh named, so new vertex data in the
    same buffers gives that mesh's result -- bit-equal to the ordinary calls, and a bad mesh is still reported."""
    torch = pytest.importorskip("torch")
    from flyby_b200 import FlyByError, _cabi
    dev = torch.device("cuda:0")
    ctx = _cabi.Context(0)
    try:
        meshes = [synth.bumpy_sphere(24, s) for s in (1, 2, 3)]
        n = ctx.views_icosahedron(2, 2.0)
        opts = ctx.render_opts()
        res = 96
        want = []
        for P, F in meshes:
            ctx.set_mesh(P, F)
            ctx.build()
            want.append(ctx.render(res, opts))
        tp = torch.from_numpy(meshes[0][0]).to(dev)
        tf = torch.from_numpy(meshes[0][1]).to(dev)
        feat = torch.empty((n, res, res, 4), dtype=torch.float32, device=dev)
        ids = torch.empty((n, res, res), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()

        def step():
            ctx.set_mesh(tp, tf, deferred=True)
            ctx.build()
            ctx.render_into(res, opts, feat, ids)

        step()                      # warm-up: every buffer gets its size
        ctx.synchronize()
        l0 = ctx.launch_count()
        with ctx.capture() as cap:
            step()
        assert ctx.launch_count() == l0          # recorded, not launched
        for k in (1, 2, 0, 1):
            tp.copy_(torch.from_numpy(meshes[k][0]))
            torch.cuda.synchronize()
            feat.zero_(); ids.zero_()
            torch.cuda.synchronize()
            before = ctx.launch_count()
            ctx.graph_launch(cap.graph)
            ctx.synchronize()
            assert ctx.launch_count() - before > 20  # the replay accounts for the kernels it ran
            assert np.array_equal(ids.cpu().numpy(), want[k][1]) and np.array_equal(feat.cpu().numpy(), want[k][0]), k
        # ordinary calls still work after replays (front buffer, tree events)
        ctx.set_mesh(*meshes[2])
        ctx.build()
        f, i = ctx.render(res, opts)
        assert np.array_equal(i, want[2][1]) and np.array_equal(f, want[2][0])
        # a replay over corrupted cells reports at the next synchronising call
        ctx.graph_launch(cap.graph)
        ctx.synchronize()
        bad = meshes[0][1].copy()
        bad[3, 2] = 10 ** 6
        tf.copy_(torch.from_numpy(bad))
        torch.cuda.synchronize()
        ctx.graph_launch(cap.graph)
        with pytest.raises(FlyByError) as e:
            ctx.synchronize()
        assert e.value.status == _cabi.ERR_INVALID
        ctx.graph_destroy(cap.graph)
        # what a captured step must not do
        with pytest.raises(FlyByError):
            with ctx.capture():
                ctx.set_mesh(tp, tf)           # not deferred
                ctx.build()
    finally:
        ctx.close()


def test_winding_fix_on_a_large_mesh(gpu_ctx, orc):
    """The union-find with parity of the consistency pass on 180 000 cells with a third of them reversed, against the
    oracle's literal wave traversal; and the cheap detection leaves a consistent mesh of that size alone."""
    P, F = synth.bumpy_sphere(95, 8)
    G, flip = _scramble_winding(F, 17, frac=0.33)
    want_tris, want_flip, n = orc.orient_cells(G, len(P))
    assert n > 50000
    gpu_ctx.set_mesh(P, G, consistency="fix")
    gpu_ctx.build()
    n_gpu, got_flip = gpu_ctx.get_winding()
    assert n_gpu == n and np.array_equal(got_flip, want_flip)
    U, _, _ = orc.unit_surf(P)
    assert np.array_equal(gpu_ctx.get_mesh()[1], orc.point_normals(U, want_tris))
    # cell 0 is never reversed (VTK's seed); if cell 0 itself was scrambled the whole component follows it
    assert got_flip[0] == 0
    gpu_ctx.set_mesh(P, F, consistency="fix")
    gpu_ctx.build()
    assert gpu_ctx.get_winding()[0] == 0
