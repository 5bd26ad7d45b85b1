"""The CUDA path (through the C-ABI) against vectors computed by the REFERENCE's own Python code
(tests/golden/ref_*.npz, written by tests/golden/make_ref_golden.py from /root/reference/src/py
under the container-only vtk stand-in).  CPU twin: tests/test_ref_golden.py (oracle vs the same
vectors).  Nothing here reads /root/reference.  Bar: bit-exact.
"""
import numpy as np
import pytest

from test_ref_golden import gold, view_cases, idc_inputs, pp_expected, predict_v3_sequence, vote_inputs, PP_CASES

pytestmark = pytest.mark.gpu


def test_icosahedron_views_equal_reference(gpu_ctx):
    g, ico, _ = view_cases()
    for key, L, r in ico:
        n = gpu_ctx.views_icosahedron(L, r, dedupe=1)
        assert n == len(g[key]) and np.array_equal(gpu_ctx.get_views(), g[key]), key
        if L != 3:   # the reference leaks duplicates only at L = 3 among these
            assert gpu_ctx.views_icosahedron(L, r, dedupe=0) == len(g[key])
            assert np.array_equal(gpu_ctx.get_views(), g[key]), key
    assert gpu_ctx.views_icosahedron(3, 2.0, dedupe=0) == 92 and len(g["ico_L3_r2"]) == 94


def test_spiral_views_equal_reference(gpu_ctx):
    g, _, spi = view_cases()
    for key, R, N, turns in spi:
        assert gpu_ctx.views_spiral(N, turns, R) == N
        assert np.array_equal(gpu_ctx.get_views(), g[key]), key


@pytest.mark.parametrize("name", ["bumpy", "crown", "shifted", "aniso"])
def test_unit_surf_equals_reference_scalesurf(gpu_ctx, name):
    g = gold("unit_surf")
    P, F = g["us_%s_P" % name], g["us_%s_F" % name]
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    uv, _, centre, scale = gpu_ctx.get_mesh()
    assert np.array_equal(centre, g["us_%s_mean" % name]) and scale == float(g["us_%s_scale" % name])
    assert np.array_equal(uv, g["us_%s_unit" % name].astype(np.float32))
    gpu_ctx.set_mesh(P, F, translate=g["us_given_mean"], scale_factor=float(g["us_given_scale"]))
    gpu_ctx.build()
    assert np.array_equal(gpu_ctx.get_mesh()[0], g["us_%s_unit_given" % name].astype(np.float32))


def test_point_features_equal_reference_codec(gpu_ctx):
    g, P, F, feats3, feats1 = idc_inputs()
    gpu_ctx.set_mesh(P, F)            # no build: the id-map lookup only needs the cells
    cells = g["idc_cells"]
    assert np.array_equal(gpu_ctx.point_features(cells, feats3, zero=-2.0).astype(np.float64), g["idc_out3"])
    assert np.array_equal(gpu_ctx.point_features(cells, feats1.reshape(-1, 1).copy(), zero=0.5).astype(np.float64), g["idc_out1"])
    assert np.array_equal(gpu_ctx.point_features(g["idc_cells_full"], P, zero=0.0).astype(np.float64), g["idc_out_coords"])
    # the decoded id map itself: point id = b*255^2 + g*255 + r - 1 of the reference's colours (utils.py:652)
    rgb = g["idc_rgb"].astype(np.int64)
    pid_ref = rgb[..., 2] * 255 * 255 + rgb[..., 1] * 255 + rgb[..., 0] - 1
    assert np.array_equal(gpu_ctx.point_ids(cells), pid_ref.astype(np.int32))
    # mode 1 of point_features takes those decoded point ids
    assert np.array_equal(gpu_ctx.point_features(pid_ref.astype(np.int32), feats3, zero=-2.0, mode=1).astype(np.float64),
                          g["idc_out3"])


class _GpuPP(object):
    """The pp_* call shapes of the oracle wrapper on top of the C-ABI (labels edited in place on a copy)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def pp_remove_islands(self, F, V, x, label, min_count, ignore_neg1=False):
        return self.ctx.labels_remove_islands(x.copy(), label, min_count, ignore_neg1)

    def pp_connectivity(self, F, V, x, label, start):
        y = x.copy()
        n = self.ctx.labels_connectivity(y, label, start)
        return y, n

    def pp_dilate(self, F, V, x, label, iterations=2, over_target=False, target=None):
        return self.ctx.labels_dilate(x.copy(), label, iterations, over_target, target)

    def pp_erode(self, F, V, x, label, ignore_label=None, iterations=None, target=None):
        return self.ctx.labels_erode(x.copy(), label, ignore_label, iterations, target)

    def pp_relabel(self, x, label, relabel):
        return self.ctx.labels_relabel(x.copy(), label, relabel)


@pytest.mark.parametrize("name", PP_CASES)
def test_post_process_equals_reference(gpu_ctx, name):
    from flyby_b200 import synth
    g = gold("post_process")
    F, V, lab = g["pp_%s_F" % name], int(g["pp_%s_V" % name]), g["pp_%s_in" % name]
    # the operations only read the connectivity: any point coordinates do
    P = {"s6": lambda: synth.bumpy_sphere(6, seed=42), "s9": lambda: synth.bumpy_sphere(9, seed=4),
         "crown7": lambda: synth.dental_crown(7, seed=1)}[name]()[0]
    assert len(P) == V
    gpu_ctx.set_mesh(np.asarray(P, np.float32), F)
    gpu_ctx.build()
    api = _GpuPP(gpu_ctx)
    for key, fn in pp_expected(g, name):
        got, ref = fn(api, F, V, lab), g["pp_%s_%s" % (name, key)]
        assert np.array_equal(got, ref), (name, key, int((got != ref).sum()))
    assert np.array_equal(predict_v3_sequence(api, F, V, lab), g["pp_%s_predict_v3_sequence" % name])


def test_votes_equal_reference_loops(gpu_ctx):
    from flyby_b200 import synth
    g, F, V, cells, pid = vote_inputs()
    K, pred = int(g["vote_K"]), g["vote_pred"]
    P = np.asarray(synth.bumpy_sphere(6, seed=42)[0], np.float32)
    gpu_ctx.set_mesh(P, F)
    assert np.array_equal(gpu_ctx.point_ids(cells), pid)
    miss = cells < 0
    votes = gpu_ctx.backproject_points(cells, pred, K)          # predict_v3.py:59-69
    ref = g["v3_counts"]
    assert np.array_equal(votes[:-1], ref[:-1])
    assert np.array_equal(votes[-1] + np.bincount(pred[miss], minlength=K), ref[-1])     # the reference's id -1 row
    assert np.array_equal(gpu_ctx.votes_argmax(votes, default=0)[:-1], g["v3_labels"][:-1])
    votes4 = gpu_ctx.backproject_points(cells, g["p1_pred"], 4)   # predict.py:154-169
    assert np.array_equal(votes4, g["p1_counts"])
    assert np.array_equal(gpu_ctx.votes_argmax(votes4, default=-1), g["p1_labels"])
    # dental_model_seg.py:192-211: soft votes per cell (fixed point 2^-20 here), argmax, gum 33, cell -> vertex 0
    probs = g["dms_probs"]                                    # [views, K, res, res]
    Ks = probs.shape[1]
    fx = gpu_ctx.backproject_soft(cells, np.ascontiguousarray(probs.transpose(0, 2, 3, 1)), Ks)
    soft_ref = g["dms_soft"].T.copy()                         # [T, K]; misses accumulated into the last cell there
    soft_ref[-1] -= probs.transpose(0, 2, 3, 1)[miss].astype(np.float64).sum(axis=0)
    n_votes = np.bincount(cells[~miss], minlength=len(F))
    assert np.all(np.abs(fx / 1048576.0 - soft_ref) <= (n_votes[:, None] + 1) * 2.0 ** -20)
    assert np.array_equal(gpu_ctx.cells_to_points(g["dms_faces"].astype(np.int32), default=33), g["dms_points"])


def test_rays_equal_predict_py(gpu_ctx):
    g = gold("predict_rays")
    views, res, R = g["pr_views"], int(g["pr_res"]), float(g["pr_radius"])
    gpu_ctx.set_mesh(g["pr_U"], g["pr_F"], skip_normalise=True)
    gpu_ctx.build()
    assert gpu_ctx.views_icosahedron(1, R, dedupe=1) == len(views)
    assert np.array_equal(gpu_ctx.get_views(), views)
    rays = gpu_ctx.generate_rays(res, gpu_ctx.render_opts()).reshape(len(views), res * res, 6)
    for v in range(1, len(views)):        # view 0 = south pole: the C++ tool's mirrored basis (see the CPU twin)
        assert np.array_equal(rays[v, :, :3], g["pr_starts"][v]) and np.array_equal(rays[v, :, 3:], g["pr_ends"][v]), v
    ids, t, w = gpu_ctx.cast(rays[1:].reshape(-1, 6), tolerance=1e-8)
    assert np.array_equal(gpu_ctx.point_ids(ids).reshape(len(views) - 1, -1), g["pr_pointids"][1:])
