"""GPU parity of the raster-compatible pinhole views (flyby_render_perspective) against the oracle."""
import numpy as np
import pytest

from flyby_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["bumpy", "crown"])
def test_perspective_matches_oracle(gpu_ctx, orc, name):
    P, F = synth.bumpy_sphere(24, 2) if name == "bumpy" else synth.dental_crown(30, 3)
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    uv, nrm, _, _ = gpu_ctx.get_mesh()
    gpu_ctx.views_icosahedron(1, 2.5)           # includes both poles (view-up switches to +-x there)
    views = gpu_ctx.get_views()
    for cfg in (dict(use_z=1, split_z=1, normal_encoding="uint8"), dict(use_z=1, split_z=0, normal_encoding="unit01", z_scale=0.25),
                dict(use_z=0, split_z=0, normal_encoding="signed")):
        opts = gpu_ctx.render_opts(**cfg)
        for res, angle, zn, zf in ((64, 30.0, 0.01, 4.0), (37, 50.0, 1.0, 3.2)):
            feat, ids, depth = gpu_ctx.render_perspective(res, opts, angle, zn, zf)
            fo, io, do = orc.render_perspective(uv, F, nrm, views, res, angle, zn, zf, use_z=cfg["use_z"], split_z=cfg["split_z"],
                                                normal_encoding={"unit01": 0, "signed": 2, "uint8": 3}[cfg["normal_encoding"]],
                                                zscale=cfg.get("z_scale", 1.0), grid=True, grid_div=24)
            assert int((ids != io).sum()) == 0 and np.array_equal(depth, do) and np.array_equal(feat, fo)
            assert (io >= 0).sum() > 500
    # a view range
    opts = gpu_ctx.render_opts()
    f2, i2, d2 = gpu_ctx.render_perspective(48, opts, 30.0, 0.01, 4.0, view_begin=4, view_end=9)
    _, io, do = orc.render_perspective(uv, F, nrm, views[4:9], 48, 30.0, 0.01, 4.0, grid=True, grid_div=24)
    assert np.array_equal(i2, io) and np.array_equal(d2, do)


def test_flybygenerator_perspective_mode(orc):
    from flyby_b200 import fly_by_features as fbf
    P, F = synth.bumpy_sphere(16, 4)
    surf = fbf.GetUnitSurf(fbf.Surf(P, F, {}, {}))
    sphere = fbf.CreateIcosahedron(radius=4.0, sl=1)
    gen = fbf.FlyByGenerator(sphere, resolution=64, use_z=True, split_z=True, projection="perspective")
    gen.addActor(fbf.GetNormalsActor(surf))
    img = gen.getFlyBy()
    assert img.shape == (12, 64, 64, 4) and img.dtype == np.float32
    rgb = img[..., :3]
    assert np.array_equal(rgb, np.round(rgb)) and rgb.max() <= 255 and rgb.min() >= 0       # uint8 colours
    hit = gen.last_cell_ids >= 0
    assert hit.any() and (~hit).any() and (img[~hit] == 0).all()                            # background is zero
    assert np.all(img[..., 3][hit] > 2.9) and np.all(img[..., 3][hit] < 5.1)                # depth along the view axis
    # against the oracle with the same clipping range
    lo, hi = surf.points.min(0), surf.points.max(0)
    zn, zf = fbf.clipping_range(sphere.points[3], (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]))
    N = orc.point_normals(surf.points, F)
    fo, io, _ = orc.render_perspective(surf.points, F, N, sphere.points[3:4], 64, 30.0, zn, zf, normal_encoding=3, grid=False)
    assert np.array_equal(img[3], fo[0]) and np.array_equal(gen.last_cell_ids[3], io[0])
    # id map actor in the same mode
    gen2 = fbf.FlyByGenerator(sphere, resolution=32, projection="perspective")
    gen2.addActor(fbf.GetPointIdMapActor(surf))
    rgb_ids = gen2.getFlyBy()
    pid = fbf.DecodeIdsRGB(rgb_ids)
    assert np.array_equal(pid >= 0, gen2.last_cell_ids >= 0)
