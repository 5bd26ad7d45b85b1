"""GPU tests of the reference-facing host layer: the fly_by_features command line, the
FlyByGenerator class, the id-map / point-feature path and the labeling driver, each
checked against the oracle."""
import numpy as np
import pytest

from flyby_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fbf():
    from flyby_b200 import fly_by_features
    return fly_by_features


def _mesh_file(tmp_path, P, F, **pd):
    from flyby_b200 import readers
    path = str(tmp_path / "surf.vtk")
    readers.write_vtk(path, P, F, point_data=pd, binary=True)
    return path


def test_free_functions_match_oracle(fbf, orc):
    P, F = synth.bumpy_sphere(16, 6)
    surf = fbf.Surf(P, F)
    unit, mean, scale = fbf.ScaleSurf(surf)
    U, c, s = orc.unit_surf(P)
    assert np.array_equal(unit.points, U) and np.array_equal(mean, c) and scale == s and unit.is_unit
    assert np.array_equal(fbf.GetUnitSurf(surf, mean_arr=[1, 2, 3], scale_factor=0.05).points,
                          orc.unit_surf(P, translate=[1, 2, 3], scale_factor=0.05)[0])
    assert np.array_equal(fbf.ComputeNormals(unit).point_data["Normals"], orc.point_normals(U, F))
    assert np.array_equal(fbf.CreateIcosahedron(2.75, 3).points, orc.icosahedron(3, 2.75, dedupe=1))
    assert np.array_equal(fbf.CreateIcosahedron(2.75, 3, dedupe=0).points, orc.icosahedron(3, 2.75))
    sp = fbf.CreateSpiral(4, 64, 4)
    assert sp.GetNumberOfPoints() == 64 and np.array_equal(sp.points, orc.spiral(64, 4, 4.0))


def test_cli_config1_small(fbf, orc, tmp_path):
    """BASELINE config 1 option line at a reduced resolution, through main() to an NRRD file."""
    from flyby_b200 import nrrd
    P, F = synth.bumpy_sphere(20, 0)
    surf = _mesh_file(tmp_path, P, F)
    out = str(tmp_path / "out.nrrd")
    args = fbf.build_parser().parse_args(("--surf %s --subdivision 2 --resolution 64 --radius 2 --use_z 1 --split_z 1 "
                                          "--normal_encoding unit01 --out %s" % (surf, out)).split())
    fbf.main(args)
    img, hdr = nrrd.read(out)
    assert img.shape == (42, 64, 64, 4) and hdr["sizes"] == "4 64 64 42"
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    ref, ids = orc.render(U, F, N, orc.icosahedron(2, 2.0), 2.0, 64, use_z=1, split_z=1, grid=True, grid_div=16)
    assert np.allclose(img, ref, rtol=1e-5, atol=1e-7) and (ids >= 0).mean() > 0.5


def test_cli_cxx_tool_output_contract(fbf, orc, tmp_path):
    """--flyByCompose / --planeSpacing / --planeScaleFactor: the C++ tool's output (fly_by_features.cxx:802-807,
    1047-1110): itk::Image<RGBPixel<double>,3>, pixel = (n*0.5+0.5)*depth, spacing (planeSpacing, planeSpacing, 1)."""
    import os
    from flyby_b200 import nrrd
    P, F = synth.bumpy_sphere(16, 3)
    surf = _mesh_file(tmp_path, P, F)
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.icosahedron(1, 3.0, dedupe=1)
    ref, ids = orc.render(U, F, N, views, 3.0, 48, plane_scale=2.0, spacing=0.75, use_z=1, split_z=0, normal_encoding=0,
                          zscale=1.0, grid=False)
    out = str(tmp_path / "vol.nrrd")
    base = "--surf %s --subdivision 1 --resolution 48 --radius 3 --planeSpacing 0.75 --planeScaleFactor 2 " % surf
    fbf.main(fbf.build_parser().parse_args((base + "--flyByCompose 1 --out " + out).split()))
    img, hdr = nrrd.read(out)
    assert img.dtype == np.float64 and hdr["type"] == "double" and hdr["sizes"] == "3 48 48 12"
    assert hdr["kinds"].split()[0] == "RGB-color" and "(0.75,0,0) (0,0.75,0) (0,0,1)" in hdr["space directions"]
    assert np.array_equal(img, ref.astype(np.float64)) and (ids >= 0).any() and (ids < 0).any()
    outdir = str(tmp_path / "per_view")
    fbf.main(fbf.build_parser().parse_args((base + "--flyByCompose 0 --out " + outdir).split()))
    assert sorted(os.listdir(outdir), key=lambda s: int(s.split(".")[0])) == ["%d.nrrd" % i for i in range(12)]
    one, hdr = nrrd.read(os.path.join(outdir, "5.nrrd"))
    assert one.dtype == np.float64 and hdr["sizes"] == "3 48 48" and np.array_equal(one, ref[5].astype(np.float64))


def test_cli_spiral_point_features_and_per_view_files(fbf, orc, tmp_path):
    import os
    from flyby_b200 import nrrd
    P, F = synth.dental_crown(16, 1)
    curv = np.linspace(-1, 1, len(P)).astype(np.float32)
    surf = _mesh_file(tmp_path, P, F, Curv=curv)
    outdir = str(tmp_path / "views")
    args = fbf.build_parser().parse_args(("--surf %s --spiral 6 --turns 4 --resolution 40 --radius 4 --use_z 1 --split_z 0 "
                                          "--point_features Curv coords --point_features_concat 1 --zero -5 "
                                          "--concatenate 0 --out %s" % (surf, outdir)).split())
    fbf.main(args)
    files = sorted(os.listdir(outdir), key=lambda s: int(s.split(".")[0]))
    assert files == ["%d.nrrd" % i for i in range(6)]
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.spiral(6, 4, 4.0)
    ref, ids = orc.render(U, F, N, views, 4.0, 40, use_z=1, split_z=0, normal_encoding=1, zscale=1.0 / 5.0, grid=False)
    for i in range(6):
        img, hdr = nrrd.read(os.path.join(outdir, files[i]))
        assert img.shape == (40, 40, 7) and hdr["space dimension"] == "2"
        assert np.allclose(img[..., :3], ref[i], rtol=1e-5, atol=1e-6)
        v0 = F[np.maximum(ids[i], 0), 0]
        hit = ids[i] >= 0
        assert np.array_equal(img[..., 3][hit], curv[v0][hit]) and np.all(img[..., 3][~hit] == -5)
        assert np.array_equal(img[..., 4:][hit], U[v0][hit]) and np.all(img[..., 4:][~hit] == -5)


def test_flybygenerator_predict_v3_flow(fbf, orc):
    """The call sequence of the reference's predict_v3.py:23-80 with a stand-in for the network."""
    P, F = synth.bumpy_sphere(18, 3)
    surf = fbf.Surf(P, F)
    unit_surf = fbf.GetUnitSurf(surf)
    sphere = fbf.CreateIcosahedron(radius=2.75, sl=2)
    flyby = fbf.FlyByGenerator(sphere, resolution=48, visualize=False, use_z=True, split_z=True)
    flyby.addActor(fbf.GetNormalsActor(unit_surf))
    img_np = flyby.getFlyBy()
    flyby.removeActors()
    assert img_np.shape == (42, 48, 48, 4) and img_np.dtype == np.float32
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    ref, ids = orc.render(U, F, N, sphere.points, 2.75, 48, use_z=1, split_z=1, normal_encoding=1, grid=False)
    assert np.allclose(img_np, ref, rtol=1e-5, atol=1e-5)       # byte255 range, like the Python reference
    assert img_np[..., :3].max() <= 255.0 and img_np[..., :3].max() > 200
    flyby_features = fbf.FlyByGenerator(sphere, 48, visualize=False)
    flyby_features.addActor(fbf.GetPointIdMapActor(unit_surf))
    rgb = flyby_features.getFlyBy()
    assert rgb.dtype == np.uint8 and rgb.shape == (42, 48, 48, 3)
    assert np.array_equal(flyby_features.last_cell_ids, ids)
    point_ids = fbf.DecodeIdsRGB(rgb)                           # predict_v3.py:61-67
    assert np.array_equal(point_ids, np.where(ids >= 0, F[np.maximum(ids, 0), 0], -1))
    # per-point votes exactly as predict_v3.py:59-80, with miss pixels skipped
    K = 4
    pred = ((np.arange(ids.size).reshape(ids.shape) * 3) % K).astype(np.int32)
    count = np.zeros((len(P), K))
    np.add.at(count, (point_ids[point_ids >= 0], pred[point_ids >= 0]), 1)
    feats = fbf.ExtractPointFeatures(unit_surf, rgb, "coords", zero=0)
    assert np.array_equal(feats[ids >= 0], U[point_ids[ids >= 0]]) and np.all(feats[ids < 0] == 0)
    with pytest.raises(NotImplementedError):
        flyby.getFlyBy(view_up_points=np.zeros((42, 3)))


def test_labeling_driver_single_process(orc):
    """Config 4 in miniature: views -> a random-init torch network -> per-cell votes -> argmax."""
    torch = pytest.importorskip("torch")
    from flyby_b200 import _cabi, labeling
    P, F = synth.dental_crown(20, 2)
    ctx = _cabi.Context(0)
    ctx.set_mesh(P, F)
    ctx.build()
    ctx.views_icosahedron(1, 2.0)
    opts = ctx.render_opts(use_z=1, split_z=1)
    torch.manual_seed(0)
    K = 6
    model = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, K, 1)).cuda()
    cell_labels, votes = labeling.label_mesh_cells(ctx, model, K, 56, opts, default=33, batch_views=5)
    feat, ids = ctx.render(56, opts)
    with torch.no_grad():
        logits = model(torch.from_numpy(feat).cuda().permute(0, 3, 1, 2).contiguous())
    pred = logits.argmax(1).cpu().numpy().astype(np.int32)
    ref = orc.backproject(ids, pred, len(F), K)
    got = votes.cpu().numpy()
    # cuDNN may pick different algorithms for different batch sizes: allow a handful of argmax flips
    assert np.abs(got - ref).sum() <= 0.001 * ref.sum()
    assert got.sum() == (ids >= 0).sum()
    lab = cell_labels.cpu().numpy()
    assert np.array_equal(lab, orc.votes_argmax(got, default=33))
    voter = labeling.CellVoter(ctx, K, use_torch=False)
    voter.add(ids, pred)
    assert np.array_equal(voter.votes, ref)
    assert np.array_equal(voter.point_labels(default=33), orc.cells_to_points(orc.votes_argmax(ref, 33), F, len(P), 33))
    ctx.close()


def test_render_async_streams_meshes(gpu_ctx):
    """flyby_render_async: consecutive meshes through the two staging sets give what the blocking call gives."""
    torch = pytest.importorskip("torch")
    from flyby_b200 import synth
    gpu_ctx.views_icosahedron(2, 2.0)
    opts = gpu_ctx.render_opts()
    res = 96
    meshes = [synth.bumpy_sphere(20 + 4 * k, k) for k in range(5)]
    want = []
    for P, F in meshes:
        gpu_ctx.set_mesh(P, F)
        gpu_ctx.build()
        want.append(gpu_ctx.render(res, opts, want_t=True))
    n = gpu_ctx.num_views()
    bufs = [(torch.empty((n, res, res, 4), dtype=torch.float32).pin_memory(), torch.empty((n, res, res), dtype=torch.int32).pin_memory(),
             torch.empty((n, res, res), dtype=torch.float64).pin_memory()) for _ in range(2)]
    for k, (P, F) in enumerate(meshes):
        gpu_ctx.set_mesh(P, F)
        gpu_ctx.build()
        f, i, t = bufs[k % 2]
        gpu_ctx.render_into(res, opts, f.numpy(), i.numpy(), t.numpy(), wait=False)
        if k >= 1:  # the previous mesh's buffers: complete once everything queued so far has drained
            gpu_ctx.synchronize()
            pf, pi, pt = bufs[(k - 1) % 2]
            assert np.array_equal(pf.numpy(), want[k - 1][0]) and np.array_equal(pi.numpy(), want[k - 1][1])
            assert np.array_equal(pt.numpy(), want[k - 1][2])
    gpu_ctx.synchronize()
    f, i, t = bufs[(len(meshes) - 1) % 2]
    assert np.array_equal(f.numpy(), want[-1][0]) and np.array_equal(i.numpy(), want[-1][1])


@pytest.mark.parametrize("res,split_z", [(96, 1), (33, 1), (64, 0)])
def test_render_packed_equals_the_float_tensor(gpu_ctx, res, split_z):
    """flyby_render_packed (uint8 colours + float32 depth, 7 B per pixel) holds exactly what flyby_render writes for
    normal_encoding "uint8": streamed through pinned host buffers (two meshes, two staging sets), into device tensors,
    at a resolution whose pixel count is not a multiple of 4 (scalar packing path), and without a depth channel."""
    torch = pytest.importorskip("torch")
    from flyby_b200 import synth, _cabi
    gpu_ctx.views_icosahedron(2, 2.0)
    use_z = 1 if split_z else 0
    opts = gpu_ctx.render_opts(use_z=use_z, split_z=split_z, normal_encoding="uint8")
    meshes = [synth.bumpy_sphere(16 + 6 * k, k) for k in range(3)]
    n = gpu_ctx.num_views()
    C = 4 if split_z else 3
    bufs = [(torch.empty((n, res, res, 3), dtype=torch.uint8).pin_memory(), torch.empty((n, res, res), dtype=torch.float32).pin_memory(),
             torch.empty((n, res, res), dtype=torch.int32).pin_memory()) for _ in range(2)]
    want, got = [], []
    for P, F in meshes:
        gpu_ctx.set_mesh(P, F)
        gpu_ctx.build()
        want.append(gpu_ctx.render(res, opts))
    for k, (P, F) in enumerate(meshes):
        gpu_ctx.set_mesh(P, F)
        gpu_ctx.build()
        rgb, z, ids = bufs[k % 2]
        gpu_ctx.render_packed_into(res, opts, rgb.numpy(), z.numpy() if split_z else None, ids.numpy())
        gpu_ctx.synchronize()
        got.append((rgb.numpy().copy(), z.numpy().copy(), ids.numpy().copy()))
    for (feat, ids_w), (rgb, z, ids) in zip(want, got):
        assert feat.shape[-1] == C and np.array_equal(ids_w, ids)
        assert np.array_equal(feat[..., :3], rgb.astype(np.float32))  # the colours are integers 0..255
        assert rgb.max() > 200 and (rgb[ids_w < 0] == 0).all()
        if split_z:
            assert np.array_equal(feat[..., 3].view(np.int32), z.view(np.int32))
    # device outputs, in stream order; two back-to-back streamed calls without a synchronisation in between
    dev = torch.device("cuda", 0)
    rgb_d = torch.empty((n, res, res, 3), dtype=torch.uint8, device=dev)
    z_d = torch.empty((n, res, res), dtype=torch.float32, device=dev) if split_z else None
    torch.cuda.synchronize()
    gpu_ctx.render_packed_into(res, opts, rgb_d, z_d, None)
    gpu_ctx.render_packed_into(res, opts, bufs[0][0].numpy(), bufs[0][1].numpy() if split_z else None, None)
    gpu_ctx.render_packed_into(res, opts, bufs[1][0].numpy(), bufs[1][1].numpy() if split_z else None, None)
    gpu_ctx.synchronize()
    assert np.array_equal(rgb_d.cpu().numpy(), got[-1][0]) and np.array_equal(bufs[0][0].numpy(), got[-1][0])
    assert np.array_equal(bufs[1][0].numpy(), got[-1][0])
    if split_z:
        assert np.array_equal(z_d.cpu().numpy(), got[-1][1]) and np.array_equal(bufs[1][1].numpy(), got[-1][1])
    # contracts the packed form cannot carry are refused
    with pytest.raises(_cabi.FlyByError):
        gpu_ctx.render_packed_into(res, gpu_ctx.render_opts(normal_encoding="unit01"), rgb_d, None, None)
    with pytest.raises(_cabi.FlyByError):
        gpu_ctx.render_packed_into(res, gpu_ctx.render_opts(use_z=1, split_z=0, normal_encoding="uint8"), rgb_d, None, None)


def test_share_gpu_flag_changes_no_result(gpu_ctx):
    """FLYBY_SHARE_GPU only changes how the resolve kernels are launched (four persistent CTAs per SM): same tensors,
    exact and tolerance shading, also for a render of a single view (fewer pixels than the persistent grid has threads)."""
    from flyby_b200 import synth
    P, F = synth.dental_crown(40, 2)
    gpu_ctx.set_mesh(P, F)
    gpu_ctx.build()
    gpu_ctx.views_icosahedron(1, 2.0)
    for kw in (dict(), dict(shade_tolerance=True), dict(use_z=1, split_z=0, plane_scale=2.0)):
        for res, ve in ((160, None), (24, 1)):
            f0, i0 = gpu_ctx.render(res, gpu_ctx.render_opts(**kw), view_end=ve)
            f1, i1 = gpu_ctx.render(res, gpu_ctx.render_opts(share_gpu=True, **kw), view_end=ve)
            assert np.array_equal(i0, i1) and np.array_equal(f0.view(np.int32), f1.view(np.int32))
            assert (i0 >= 0).any()


def test_two_contexts_interleaved_on_one_gpu(gpu_ctx):
    """Independent contexts on their own streams may overlap on the device: results equal the sequential ones."""
    torch = pytest.importorskip("torch")
    from flyby_b200 import _cabi, synth
    dev = torch.device("cuda", 0)
    meshes = [synth.bumpy_sphere(30 + 3 * k, k) for k in range(6)]
    gpu_ctx.views_icosahedron(2, 2.0)
    opts = gpu_ctx.render_opts()
    want = []
    for P, F in meshes:
        gpu_ctx.set_mesh(P, F)
        gpu_ctx.build()
        want.append(gpu_ctx.render(128, opts))
    pair = []
    for _ in range(2):
        s = torch.cuda.Stream(device=dev)
        c = _cabi.Context(0, stream=s.cuda_stream)
        n = c.views_icosahedron(2, 2.0)
        pair.append((c, s, [torch.empty((n, 128, 128, 4), dtype=torch.float32, device=dev) for _ in range(3)],
                     [torch.empty((n, 128, 128), dtype=torch.int32, device=dev) for _ in range(3)]))
    dm = [(torch.from_numpy(P).to(dev), torch.from_numpy(F).to(dev)) for P, F in meshes]
    torch.cuda.synchronize()  # the uploads ran on torch's stream; the contexts work on their own streams
    for k, (P, F) in enumerate(dm):          # nothing waits between the calls: the two streams run side by side
        c, s, fs, is_ = pair[k % 2]
        c.set_mesh(P, F)
        c.build()
        c.render_into(128, opts, fs[k // 2], is_[k // 2])
    torch.cuda.synchronize()
    for k in range(len(meshes)):
        c, s, fs, is_ = pair[k % 2]
        assert np.array_equal(fs[k // 2].cpu().numpy(), want[k][0]) and np.array_equal(is_[k // 2].cpu().numpy(), want[k][1])
    for c, _, _, _ in pair:
        c.close()
