"""GPU tests of the barycentric feature lookup (flyby_interpolate_features) and of the torch op that gives the
fly-by renderer the signature of the reference's in-model renderer (challenge-teeth/teeth_nets.py:121-153)."""
import numpy as np
import pytest

from flyby_b200 import synth

pytestmark = pytest.mark.gpu


def test_interpolate_features_matches_oracle(gpu_ctx, orc):
    P, F = synth.bumpy_sphere(24, 3)
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    uv, nrm, _, _ = gpu_ctx.get_mesh()
    gpu_ctx.views_icosahedron(1, 2.0)
    views = gpu_ctx.get_views()
    rng = np.random.default_rng(0)
    feats = rng.normal(size=(len(P), 5)).astype(np.float32)
    for ps, res in ((1.0, 64), (2.0, 47)):
        opts = gpu_ctx.render_opts(use_z=1, split_z=1, plane_scale=ps)
        _, ids = gpu_ctx.render(res, opts)
        got = gpu_ctx.interpolate_features(res, opts, ids, feats, zero=-7.0)
        ref = orc.interpolate_features(uv, F, views, 2.0, res, ids, feats, plane_scale=ps, zero=-7.0)
        assert np.array_equal(got, ref)
        assert (got[ids < 0] == -7.0).all() and (ids >= 0).sum() > 100
        # interpolating the point normals reproduces the render's own normal channels (unit01 encoding: n/2 + 1/2)
        feat, _ = gpu_ctx.render(res, opts)
        nn = gpu_ctx.interpolate_features(res, opts, ids, nrm)
        hit = ids >= 0
        assert np.allclose(feat[..., :3][hit], nn[hit] * 0.5 + 0.5, atol=1e-6)
    # a view range
    opts = gpu_ctx.render_opts(plane_scale=2.0)
    _, ids = gpu_ctx.render(32, opts, view_begin=3, view_end=7)
    got = gpu_ctx.interpolate_features(32, opts, ids, feats, view_begin=3, view_end=7)
    assert np.array_equal(got, orc.interpolate_features(uv, F, views[3:7], 2.0, 32, ids, feats, plane_scale=2.0))


def test_torch_renderer_signature_and_values(orc):
    torch = pytest.importorskip("torch")
    from flyby_b200 import torch_ops
    meshes = [synth.bumpy_sphere(12, s) for s in (1, 2)]
    U = [orc.unit_surf(P)[0] for P, _ in meshes]
    V = torch.from_numpy(np.stack(U)).cuda()
    F = torch.from_numpy(np.stack([f for _, f in meshes]).astype(np.int64)).cuda()
    CN = torch.from_numpy(np.stack([orc.point_normals(u, f) for u, (_, f) in zip(U, meshes)])).cuda()
    r = torch_ops.FlyByRenderer(radius=1.35, subdivision_level=1, image_size=48, plane_scale=2.0)
    X, PF = r.render(V, F, CN)
    torch.cuda.synchronize()
    T = 12
    assert X.shape == (2, T, 4, 48, 48) and PF.shape == (2, T, 1, 48, 48)
    assert X.dtype == torch.float32 and PF.dtype == torch.int64
    views = orc.icosahedron(1, 1.35)
    nf = F.shape[1]
    for b in range(2):
        Fb = meshes[b][1]
        N = orc.point_normals(U[b], Fb)
        _, io, to = orc.render(U[b], Fb, N, views, 1.35, 48, plane_scale=2.0, grid=False, want_t=True)
        col = orc.interpolate_features(U[b], Fb, views, 1.35, 48, io, N, plane_scale=2.0, zero=1.0)
        pf = PF[b, :, 0].cpu().numpy()
        assert np.array_equal(pf, np.where(io >= 0, io.astype(np.int64) + b * nf, -1))
        assert np.array_equal(X[b, :, :3].permute(0, 2, 3, 1).cpu().numpy(), col)
        depth = np.where(io >= 0, (to * 2.7).astype(np.float32), np.float32(-1))
        assert np.array_equal(X[b, :, 3].cpu().numpy(), depth)
    # the label gather of teeth_nets.py:166
    YF = torch.arange(2 * nf, device="cuda").reshape(2, nf) % 9
    y = torch_ops.gather_face_labels(YF, PF)
    assert y.shape == PF.shape and int((y[PF < 0] != 0).sum()) == 0
    assert torch.equal(y[PF >= 0], (PF[PF >= 0] % 9))
    r.close()
