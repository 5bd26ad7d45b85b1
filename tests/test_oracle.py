"""CPU tests of the oracle (oracle/flyby_oracle.c): analytic known answers, the shape /
count facts the reference documents, brute force == bucket grid, and the committed
golden fixtures.  The reference has no tests and no golden vectors for this path
(SURVEY.md section 4), so these are the anchors there are."""
import os

import numpy as np
import pytest

from flyby_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flyby_golden.npz")


# ---------------------------------------------------------------- viewpoints (a3, a4)
def test_icosahedron_counts_and_order(orc):
    # 10 L^2 + 2 lattice points; subdivision 1 = the 12 vertices (train_useg_flyby.py:153,174: [12,512,512,4])
    for L, n in [(1, 12), (2, 42), (3, 92), (4, 162), (10, 1002)]:
        v = orc.icosahedron(L, 2.75)
        assert v.shape == (n, 3) and v.dtype == np.float32
        assert np.allclose(np.linalg.norm(v.astype(np.float64), axis=1), 2.75, rtol=1e-6)
        assert len(np.unique(np.round(v, 4), axis=0)) == n
    v = orc.icosahedron(2, 2.0)
    # first inserted point = first lattice point of face (0,5,4) = icosahedron vertex 0 = south pole
    assert np.array_equal(v[0], np.array([0, 0, -2], np.float32))
    assert (np.abs(v[:, 2] - 2.0) < 1e-6).sum() == 1  # the north pole appears exactly once


def test_icosahedron_float_dedupe_leaks_duplicates(orc):
    """Hazard H1: the reference de-duplicates by float equality and leaks duplicate viewpoints.  The counts are what
    the reference's own LinearSubdivisionFilter produces under tests/refshim (tests/golden/ref_views.npz): 94 at
    subdivision 3, none leaked at 1, 2, 4, 5, 10.  (Its readme quotes 268 / 1026 ids for subdivisions 5 / 10, which
    neither solid-scale assumption nor an octree-leaf model of the point locator reproduces: DESIGN.md section 2.)"""
    assert len(orc.icosahedron(2, 2.0, dedupe=1)) == 42
    assert len(orc.icosahedron(4, 2.0, dedupe=1)) == 162
    assert len(orc.icosahedron(3, 2.0, dedupe=1)) == 94
    assert len(orc.icosahedron(5, 2.0, dedupe=1)) == 252


def test_spiral(orc):
    v = orc.spiral(64, 4, 4.0)  # predict_LU.py:39-72 setting
    assert v.shape == (64, 3)
    assert np.array_equal(v[0], np.array([0, 0, 4], np.float32))  # i = 0 is the north pole
    i = np.arange(64)
    a = i * np.pi / 64
    ref = 4.0 * np.stack([np.sin(a) * np.cos(8 * a), np.sin(a) * np.sin(8 * a), np.cos(a)], 1)
    assert np.allclose(v, ref, atol=1e-6)
    assert np.all(np.diff(v[:, 2]) < 0)  # monotone descent from the pole


# ---------------------------------------------------------------- plane and rays (a5)
def test_view_basis_and_ray_order(orc):
    rays, B = orc.view_rays(np.array([0, 0, 2], np.float32), 2.0, 3)         # north pole special case
    assert np.array_equal(B, [[0, 0, 1], [1, 0, 0], [0, 1, 0]])
    rays_s, Bs = orc.view_rays(np.array([0, 0, -2], np.float32), 2.0, 3)     # south pole special case
    assert np.array_equal(Bs, [[0, 0, -1], [-1, 0, 0], [0, -1, 0]])
    # 3x3 plane of side 1 centred on the viewpoint, i (x axis) fastest
    o = rays[:, :3]
    assert np.allclose(o[0], [-0.5, -0.5, 2]) and np.allclose(o[1], [0, -0.5, 2]) and np.allclose(o[3], [-0.5, 0, 2])
    assert np.allclose(rays[:, 3:] - o, [0, 0, -4])                          # end = origin - n * 2R
    sp = np.array([1.0, 2.0, 0.5], np.float32)
    rays, B = orc.view_rays(sp, 2.0, 4, plane_scale=2.0)
    n, x, y = B
    assert np.allclose(n, sp / np.linalg.norm(sp))
    assert np.allclose(x, np.cross(n, [0, 0, 1]) / np.linalg.norm(np.cross(n, [0, 0, 1])))
    assert np.allclose(y, np.cross(n, x))
    c = rays[:, :3].mean(0)
    assert np.allclose(c, sp, atol=1e-6)                                     # plane centred on the viewpoint
    assert np.allclose(np.linalg.norm(rays[3, :3] - rays[0, :3]), 2.0, atol=1e-6)
    # plane spacing scales the plane about the viewpoint (cxx:659,746)
    r2, _ = orc.view_rays(sp, 2.0, 4, plane_scale=2.0, spacing=0.5)
    assert np.allclose(r2[:, :3] - sp, 0.5 * (rays[:, :3] - sp), atol=1e-6)


# ---------------------------------------------------------------- triangle test (a7)
def test_triangle_known_answers(orc):
    p0, p1, p2 = [0, 0, 0], [1, 0, 0], [0, 1, 0]
    hit, t, x, w = orc.tri_intersect([0.25, 0.25, 1], [0.25, 0.25, -1], p0, p1, p2)
    assert hit and t == 0.5 and np.allclose(x, [0.25, 0.25, 0]) and np.allclose(w, [0.5, 0.25, 0.25])
    # inclusive edges and vertices (vtkTriangle::EvaluatePosition weights in [0,1])
    assert orc.tri_intersect([0.5, 0.0, 1], [0.5, 0.0, -1], p0, p1, p2)[0]
    assert orc.tri_intersect([0.5, 0.5, 1], [0.5, 0.5, -1], p0, p1, p2)[0]
    assert orc.tri_intersect([0, 0, 1], [0, 0, -1], p0, p1, p2)[0]
    assert orc.tri_intersect([1, 0, 1], [1, 0, -1], p0, p1, p2)[0]
    # outside, beyond the segment, parallel, degenerate
    assert not orc.tri_intersect([0.75, 0.75, 1], [0.75, 0.75, -1], p0, p1, p2)[0]
    assert not orc.tri_intersect([-1e-9, 0.5, 1], [-1e-9, 0.5, -1], p0, p1, p2)[0]
    assert not orc.tri_intersect([0.25, 0.25, 1], [0.25, 0.25, 0.5], p0, p1, p2)[0]
    assert not orc.tri_intersect([0, 0, 1], [1, 1, 1], p0, p1, p2)[0]
    assert not orc.tri_intersect([0.25, 0.25, 1], [0.25, 0.25, -1], p0, p1, p1)[0]
    # the segment end points count (0 <= t <= 1)
    assert orc.tri_intersect([0.25, 0.25, 0], [0.25, 0.25, -1], p0, p1, p2)[1] == 0.0
    assert orc.tri_intersect([0.25, 0.25, 1], [0.25, 0.25, 0], p0, p1, p2)[1] == 1.0


def test_cube_tie_break_lowest_id(orc):
    V, F = synth.cube()
    # straight down the middle of the top face: exactly on the diagonal shared by its two triangles
    rays = np.array([[0, 0, 3, 0, 0, -3]], np.float64)
    ids, t, x, w = orc.cast(V, F, rays, grid=False)
    top = [c for c in range(12) if np.all(V[F[c]][:, 2] == 1)]
    assert ids[0] == min(top) and np.isclose(t[0], 1 / 3)
    ids_g, t_g, _, _ = orc.cast(V, F, rays, grid=True, grid_div=4)
    assert ids_g[0] == ids[0] and t_g[0] == t[0]


# ---------------------------------------------------------------- whole path
def test_sphere_analytic_depth_and_normals(orc):
    """An icosphere seen from the +x axis: depth = R - sqrt(r^2 - d^2) and the interpolated normal is
    radial, up to the tessellation error."""
    V, F = synth.icosphere(24)
    U, c, s = orc.unit_surf(V.astype(np.float32))
    r = float(np.linalg.norm(U.max(0)))  # unit surf: bbox corner on the unit sphere
    assert np.isclose(r, 1.0, atol=1e-6)
    rad = np.linalg.norm(U.astype(np.float64), axis=1).mean()
    N = orc.point_normals(U, F)
    assert np.allclose(N, U / np.linalg.norm(U, axis=1, keepdims=True), atol=2e-3)
    view = np.array([[2.0, 0, 0]], np.float32)
    res = 65
    feat, ids, t = orc.render(U, F, N, view, 2.0, res, plane_scale=1.0, want_t=True, grid=False)
    rays, _ = orc.view_rays(view[0], 2.0, res)
    d2 = (rays[:, 1] ** 2 + rays[:, 2] ** 2).reshape(res, res)
    inside = d2 < (0.97 * rad) ** 2
    outside = d2 > (1.001 * rad) ** 2
    assert np.all(ids[0][inside] >= 0) and np.all(ids[0][outside] == -1)
    depth = feat[0, :, :, 3]
    expect = 2.0 - np.sqrt(np.maximum(rad ** 2 - d2, 0))
    assert np.allclose(depth[inside], expect[inside], atol=3e-3)
    assert np.all(depth[ids[0] < 0] == 0) and np.all(feat[0][ids[0] < 0] == 0)
    # centre pixel: normal (1,0,0) -> encoded (1, .5, .5)
    assert np.allclose(feat[0, res // 2, res // 2, :3], [1.0, 0.5, 0.5], atol=2e-3)
    assert np.allclose(t[0][inside] * 4.0, depth[inside], rtol=1e-6)  # depth = t * 2R along the ray


def test_grid_equals_brute_force(orc):
    rng = np.random.default_rng(1)
    for P, F in (synth.bumpy_sphere(10, 1), synth.dental_crown(12, 3)):
        U, _, _ = orc.unit_surf(P)
        n = 4000
        a = rng.normal(size=(n, 3)); a /= np.linalg.norm(a, axis=1, keepdims=True)
        b = rng.normal(size=(n, 3)); b /= np.linalg.norm(b, axis=1, keepdims=True)
        rays = np.concatenate([a * 2, -a * 2 + b * 0.7], axis=1)
        ib, tb, xb, wb = orc.cast(U, F, rays, grid=False)
        for div in (0, 8, 32):
            ig, tg, xg, wg = orc.cast(U, F, rays, grid=True, grid_div=div)
            assert np.array_equal(ib, ig) and np.array_equal(tb, tg) and np.array_equal(wb, wg)
        assert (ib >= 0).sum() > n // 4


def test_channel_assembly(orc):
    P, F = synth.bumpy_sphere(8, 2)
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.icosahedron(1, 2.0)[:3]
    f4, ids = orc.render(U, F, N, views, 2.0, 32, use_z=1, split_z=1, grid=False)
    f3m, _ = orc.render(U, F, N, views, 2.0, 32, use_z=1, split_z=0, grid=False)
    f3n, _ = orc.render(U, F, N, views, 2.0, 32, use_z=0, split_z=0, grid=False)
    fs, _ = orc.render(U, F, N, views, 2.0, 32, use_z=1, split_z=1, normal_encoding=2, zscale=0.25, grid=False)
    fb, _ = orc.render(U, F, N, views, 2.0, 32, use_z=0, split_z=0, normal_encoding=1, grid=False)
    assert f4.shape[-1] == 4 and f3m.shape[-1] == 3 and f3n.shape[-1] == 3
    assert np.array_equal(f4[..., :3], f3n)                                        # normals channels
    assert np.allclose(f3m, f4[..., :3] * f4[..., 3:4], rtol=1e-6)                 # cxx:802-804
    hit = ids >= 0
    assert np.allclose(fs[..., :3][hit], (f4[..., :3][hit] - 0.5) * 2, atol=1e-6)  # signed normals
    assert np.allclose(fs[..., 3], f4[..., 3] * 0.25, rtol=1e-6)                   # z scale
    assert np.allclose(fb, f3n * 255.0, rtol=1e-6)                                 # 8-bit range


def test_unit_surf_modes(orc):
    P, F = synth.bumpy_sphere(8, 5)
    U, c, s = orc.unit_surf(P)
    lo, hi = P.min(0).astype(np.float64), P.max(0).astype(np.float64)
    assert np.allclose(c, (lo + hi) / 2) and np.isclose(s, 1 / np.linalg.norm(hi - c))
    assert np.isclose(np.linalg.norm(U.max(0)), 1.0, atol=1e-6)
    U2, c2, s2 = orc.unit_surf(P, centre_mode=1, scale_mode=1)
    assert np.allclose(c2, P.astype(np.float64).mean(0)) and np.isclose(np.linalg.norm(U2, axis=1).max(), 1.0, atol=1e-6)
    U3, c3, s3 = orc.unit_surf(P, translate=[1, 2, 3], scale_factor=0.1)
    assert np.allclose(U3, (P.astype(np.float64) - [1, 2, 3]) * 0.1, atol=1e-6)


def test_backprojection_semantics(orc):
    ids = np.array([[0, 0, 1, -1], [2, 2, 2, 1]], np.int32)
    labels = np.array([[3, 3, 1, 2], [0, 4, 4, 1]], np.int32)
    votes = orc.backproject(ids, labels, 4, 5)
    assert votes.sum() == 7                                   # the background pixel does not vote
    assert votes[0, 3] == 2 and votes[1, 1] == 2 and votes[2, 4] == 2 and votes[2, 0] == 1 and votes[3].sum() == 0
    lab = orc.votes_argmax(votes, default=33)
    assert lab.tolist() == [3, 1, 4, 33]
    tie = np.array([[2, 5, 5, 0]], np.int32)
    assert orc.votes_argmax(tie, default=-1).tolist() == [1]   # numpy argmax: first maximum
    tris = np.array([[0, 1, 2], [0, 2, 3], [3, 1, 0], [4, 0, 1]], np.int32)
    pts = orc.cells_to_points(lab, tris, 6, default=33)
    assert pts.tolist() == [1, 33, 33, 4, 33, 33]             # vertex 0 of each cell, later cells overwrite


def test_golden_fixtures(orc):
    g = np.load(GOLDEN)
    U, c, s = orc.unit_surf(g["P"])
    assert np.array_equal(U, g["U"]) and np.array_equal(c, g["centre"]) and s == g["scale"]
    N = orc.point_normals(U, g["F"])
    assert np.array_equal(N, g["N"])
    assert np.array_equal(orc.icosahedron(2, 2.0), g["ico"])
    assert np.allclose(orc.spiral(8, 4, 4.0), g["spiral"], atol=0)
    for grid in (False, True):
        f4, ids, t = orc.render(U, g["F"], N, g["ico"], 2.0, 24, grid=grid, want_t=True)
        assert np.array_equal(ids, g["ids"]) and np.array_equal(t, g["t"]) and np.array_equal(f4, g["feat4"])
    f3, ids_s = orc.render(U, g["F"], N, g["spiral"], 4.0, 24, plane_scale=2.0, use_z=1, split_z=0, grid=True)
    assert np.array_equal(ids_s, g["ids_spiral"]) and np.array_equal(f3, g["feat3_spiral"])
    rays, basis = orc.view_rays(g["ico"][5], 2.0, 5)
    assert np.array_equal(rays, g["rays_view5"]) and np.array_equal(basis, g["basis_view5"])
    votes = orc.backproject(g["ids"], g["labels"], len(g["F"]), 5)
    assert np.array_equal(votes, g["votes"])
    assert np.array_equal(orc.votes_argmax(votes, default=33), g["cell_labels"])


GOLDEN_EXT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flyby_golden_ext.npz")


def test_golden_fixtures_widened_rows(orc):
    g, e = np.load(GOLDEN), np.load(GOLDEN_EXT)
    F, V, lab = g["F"], len(g["U"]), e["labels_in"]
    assert np.array_equal(orc.pp_remove_islands(F, V, lab, 1, 30, True), e["islands"])
    conn, n = orc.pp_connectivity(F, V, lab, 2, 10)
    assert n == int(e["n_regions"]) and np.array_equal(conn, e["connectivity"])
    assert np.array_equal(orc.pp_dilate(F, V, lab, 3, 2), e["dilate"])
    assert np.array_equal(orc.pp_erode(F, V, lab, 0, ignore_label=1), e["erode"])
    assert np.array_equal(orc.pp_erode(F, V, lab, 2, iterations=2, target=3), e["erode2"])
    assert np.array_equal(orc.interpolate_features(g["U"], F, g["ico"][:8], 2.0, 24, g["ids"][:8], e["feats"], zero=-1.0), e["interp"])
    pf, pi, pd = orc.render_perspective(g["U"], F, g["N"], g["ico"][:6], 20, 30.0, 0.01, 3.0, normal_encoding=3, grid=False)
    assert np.array_equal(pf, e["persp_feat"]) and np.array_equal(pi, e["persp_ids"]) and np.array_equal(pd, e["persp_depth"])


@pytest.mark.gpu
def test_gpu_matches_golden_widened_rows(gpu_ctx):
    g, e = np.load(GOLDEN), np.load(GOLDEN_EXT)
    gpu_ctx.set_mesh(g["P"], g["F"])
    gpu_ctx.build()
    lab = e["labels_in"]
    assert np.array_equal(gpu_ctx.labels_remove_islands(lab.copy(), 1, 30, True), e["islands"])
    got = lab.copy()
    assert gpu_ctx.labels_connectivity(got, 2, 10) == int(e["n_regions"]) and np.array_equal(got, e["connectivity"])
    assert np.array_equal(gpu_ctx.labels_dilate(lab.copy(), 3, 2), e["dilate"])
    assert np.array_equal(gpu_ctx.labels_erode(lab.copy(), 0, ignore_label=1), e["erode"])
    assert np.array_equal(gpu_ctx.labels_erode(lab.copy(), 2, iterations=2, target=3), e["erode2"])
    gpu_ctx.views_icosahedron(2, 2.0)
    opts = gpu_ctx.render_opts(use_z=1, split_z=1)
    assert np.array_equal(gpu_ctx.interpolate_features(24, opts, g["ids"][:8], e["feats"], zero=-1.0, view_end=8), e["interp"])
    pf, pi, pd = gpu_ctx.render_perspective(20, gpu_ctx.render_opts(normal_encoding="uint8"), 30.0, 0.01, 3.0, view_end=6)
    assert np.array_equal(pf, e["persp_feat"]) and np.array_equal(pi, e["persp_ids"]) and np.array_equal(pd, e["persp_depth"])


@pytest.mark.gpu
def test_gpu_matches_golden(gpu_ctx):
    g = np.load(GOLDEN)
    gpu_ctx.set_mesh(g["P"], g["F"])
    gpu_ctx.build()
    uv, nrm, c, s = gpu_ctx.get_mesh()
    assert np.array_equal(uv, g["U"]) and np.array_equal(nrm, g["N"])
    assert gpu_ctx.views_icosahedron(2, 2.0) == 42 and np.array_equal(gpu_ctx.get_views(), g["ico"])
    f4, ids, t = gpu_ctx.render(24, gpu_ctx.render_opts(use_z=1, split_z=1), want_t=True)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(t, g["t"]) and np.array_equal(f4, g["feat4"])
    gpu_ctx.views_custom(g["spiral"], 4.0)
    f3, ids_s = gpu_ctx.render(24, gpu_ctx.render_opts(use_z=1, split_z=0, plane_scale=2.0))
    assert np.array_equal(ids_s, g["ids_spiral"]) and np.array_equal(f3, g["feat3_spiral"])
    votes = gpu_ctx.backproject(ids, g["labels"], 5)
    assert np.array_equal(votes, g["votes"])
    assert np.array_equal(gpu_ctx.votes_argmax(votes, default=33), g["cell_labels"])


def test_perspective_camera_known_answers(orc):
    """Pinhole views of a sphere of radius r from distance R: the centre pixels see depth R - r with the normal
    facing the camera, the silhouette subtends asin(r/R), rows run bottom to top with view-up (0,0,-1)."""
    V, F = synth.icosphere(40)
    V = (V * 0.5).astype(np.float32)
    N = orc.point_normals(V, F)
    R, res = 2.0, 201
    views = np.array([[R, 0, 0], [0, 0, R], [0, 0, -R], [1.6, -1.0, 1.2]], np.float32)
    feat, ids, depth = orc.render_perspective(V, F, N, views, res, view_angle=30.0, z_near=0.01, z_far=R + 1,
                                              grid=True, grid_div=16)
    c = res // 2
    for v in range(4):
        Rv = float(np.linalg.norm(views[v]))
        assert abs(depth[v, c, c] - (Rv - 0.5)) < 2e-3
        nrm = feat[v, c, c, :3] * 2 - 1
        assert np.allclose(nrm, views[v] / Rv, atol=0.05) and abs(feat[v, c, c, 3] - depth[v, c, c]) < 1e-6
        # silhouette: pixels whose ray makes an angle below asin(r/R) with the axis are hits
        frac = (ids[v] >= 0).mean()
        tan_s = np.tan(np.arcsin(0.5 / Rv)) / np.tan(np.radians(15.0))
        assert abs(frac - np.pi * tan_s ** 2 / 4) < 0.01
        assert (depth[v][ids[v] < 0] == -1).all() and (feat[v][ids[v] < 0] == 0).all()
    # view-up (0,0,-1): for the camera on +x the top rows (large j) look towards -z
    top, bottom = feat[0, res - 30, c, :3] * 2 - 1, feat[0, 30, c, :3] * 2 - 1
    assert top[2] < -0.3 and bottom[2] > 0.3
    # uint8 colours: integers in 0..255
    f8, _, _ = orc.render_perspective(V, F, N, views[:1], 64, z_far=R + 1, normal_encoding=3)
    assert np.array_equal(f8[..., :3], np.round(f8[..., :3])) and f8[..., :3].max() <= 255 and f8[..., :3].min() >= 0


def test_oracle_agrees_with_an_independent_ray_caster(orc):
    """The oracle restates VTK's plane-hit + dominant-axis projection.  A different algorithm -- Moeller-Trumbore in
    numpy float64, brute force over all triangles -- must find the same nearest cell and the same segment parameter on
    every ray that does not graze an edge (those are counted: VTK's inclusive weights and tolerance band decide them,
    Moeller-Trumbore's do not)."""
    from flyby_b200 import synth
    P, F = synth.bumpy_sphere(10, 7)          # 2000 triangles
    U, _, _ = orc.unit_surf(P)
    rng = np.random.default_rng(5)
    n = 1500
    o = rng.normal(size=(n, 3)); o = 2.5 * o / np.linalg.norm(o, axis=1, keepdims=True)
    target = rng.uniform(-0.6, 0.6, (n, 3))
    e = o + 2.0 * (target - o)                # segments through the mesh, some missing it
    e[:100] = o[:100] + 0.2 * (target[:100] - o[:100])   # too short to reach it
    rays = np.concatenate([o, e], axis=1)
    ids, t, x, w = orc.cast(U, F, rays, grid=True)
    ids_b, t_b, _, _ = orc.cast(U, F, rays, grid=False)
    assert np.array_equal(ids, ids_b) and np.array_equal(t, t_b)

    V0, V1, V2 = (U[F[:, k]].astype(np.float64) for k in range(3))
    e1, e2 = V1 - V0, V2 - V0
    best_t = np.full(n, np.inf); best_id = np.full(n, -1); margin = np.full(n, np.inf)
    for k in range(n):
        d = e[k] - o[k]
        pv = np.cross(d, e2)
        det = np.einsum("ij,ij->i", e1, pv)
        ok = np.abs(det) > 1e-300
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o[k] - V0
        u = np.einsum("ij,ij->i", tv, pv) * inv
        qv = np.cross(tv, e1)
        v = (qv @ d) * inv
        tt = np.einsum("ij,ij->i", e2, qv) * inv
        edge = np.minimum(np.minimum(u, v), 1.0 - u - v)          # > 0 inside
        hit = ok & (edge >= 0) & (tt >= 0) & (tt <= 1)
        near = ok & (np.abs(edge) < 1e-9) & (tt >= -1e-9) & (tt <= 1 + 1e-9)
        if near.any():
            margin[k] = 0.0                                          # a grazing candidate exists on this ray
        if hit.any():
            tmin = tt[hit].min()
            cand = np.nonzero(hit & (tt == tmin))[0]
            best_t[k], best_id[k] = tmin, cand.min()
    clear = margin > 0
    assert clear.sum() > n - 10                                      # grazing rays are rare in a random sample
    assert np.array_equal(ids[clear], best_id[clear])
    h = clear & (ids >= 0)
    assert h.sum() > 800 and (ids[clear] < 0).sum() > 100
    assert np.allclose(t[h], best_t[h], rtol=1e-11, atol=0)
    # the hit point and the weights the oracle returns describe the same point
    recon = w[h, 0:1] * U[F[ids[h], 0]] + w[h, 1:2] * U[F[ids[h], 1]] + w[h, 2:3] * U[F[ids[h], 2]]
    assert np.abs(recon - x[h]).max() < 1e-9


def test_oracle_pinned_against_vtk_when_available():
    """Runs oracle/pin_vtk.py (the reference's own VTK calls against the oracle) where vtk is importable; skipped in
    the build environment of this repository, where it is not -- the reason DESIGN.md says "parity unpinned"."""
    import pytest
    pytest.importorskip("vtk")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "oracle", "pin_vtk.py")], capture_output=True, text=True,
                         timeout=1800)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
