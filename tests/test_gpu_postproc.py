"""GPU parity of the mesh-label post-processing (csrc/postproc.cu through the C-ABI) against the oracle's
restatement of the reference's src/py/post_process.py.  Integer work: every label must be identical."""
import numpy as np
import pytest

from flyby_b200 import synth

pytestmark = pytest.mark.gpu


def _bind(ctx, P, F):
    ctx.set_mesh(P, F, skip_normalise=True)
    ctx.build()


def _noisy_labels(P, F, seed, k=5, noise=0.08):
    """Smooth label regions (by direction) with salt-and-pepper noise: big regions plus many islands."""
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=(k, 3))
    lab = np.argmax(P @ dirs.T, axis=1).astype(np.int32)
    flip = rng.random(len(P)) < noise
    lab[flip] = rng.integers(-1, k + 1, flip.sum()).astype(np.int32)
    return lab


MESHES = {"bumpy30": lambda: synth.bumpy_sphere(30, 1), "crown40": lambda: synth.dental_crown(40, 2),
          "bumpy100": lambda: synth.bumpy_sphere(100, 0)}


@pytest.mark.parametrize("name", list(MESHES))
def test_each_operation_matches_oracle(gpu_ctx, orc, name):
    P, F = MESHES[name]()
    V = len(P)
    _bind(gpu_ctx, P, F)
    lab = _noisy_labels(P, F, 3)
    for label in (-1, 0, 2, 5):
        for min_count, flag in ((5, True), (200, True), (10 ** 9, True), (50, False)):
            got = gpu_ctx.labels_remove_islands(lab.copy(), label, min_count, flag)
            assert np.array_equal(got, orc.pp_remove_islands(F, V, lab, label, min_count, flag)), (label, min_count)
        got = lab.copy()
        n = gpu_ctx.labels_connectivity(got, label, 7)
        ref, n_ref = orc.pp_connectivity(F, V, lab, label, 7)
        assert n == n_ref and np.array_equal(got, ref)
        for it in (1, 3):
            assert np.array_equal(gpu_ctx.labels_dilate(lab.copy(), label, it), orc.pp_dilate(F, V, lab, label, it))
            assert np.array_equal(gpu_ctx.labels_dilate(lab.copy(), label, it, True, 1),
                                  orc.pp_dilate(F, V, lab, label, it, True, 1))
        for kw in (dict(), dict(iterations=2), dict(ignore_label=0), dict(target=1, iterations=5),
                   dict(ignore_label=1, iterations=9), dict(iterations=8), dict(iterations=16)):
            assert np.array_equal(gpu_ctx.labels_erode(lab.copy(), label, **kw), orc.pp_erode(F, V, lab, label, **kw)), kw
        assert np.array_equal(gpu_ctx.labels_relabel(lab.copy(), label, 40), orc.pp_relabel(lab, label, 40))


def test_predictor_sequence_matches_oracle(gpu_ctx, orc):
    """The order predict_v3.py:93-133 applies them: dilate the gum, remove islands per label, relabel the gum,
    number the teeth by connectivity, erode the gum away -- on device-resident labels when torch is present."""
    P, F = synth.dental_crown(60, 4)
    V = len(P)
    _bind(gpu_ctx, P, F)
    lab = _noisy_labels(P, F, 9, k=4, noise=0.05)
    ref = orc.pp_dilate(F, V, lab, 3, 2)
    for label in range(int(ref.min()), int(ref.max()) + 1):
        ref = orc.pp_remove_islands(F, V, ref, label, 200, True)
    ref = orc.pp_relabel(ref, 3, -1)
    teeth = int(np.bincount(ref[ref >= 0]).argmax())  # the label whose regions get numbered
    ref, n_ref = orc.pp_connectivity(F, V, ref, teeth, 10)
    ref = orc.pp_erode(F, V, ref, -1, ignore_label=0)

    torch = pytest.importorskip("torch")
    got = torch.from_numpy(lab.copy()).cuda()
    torch.cuda.synchronize()  # the upload ran on torch's stream, the context works on its own
    gpu_ctx.labels_dilate(got, 3, 2)
    lo, hi = int(got.min()), int(got.max())
    for label in range(lo, hi + 1):
        gpu_ctx.labels_remove_islands(got, label, 200, True)
    gpu_ctx.labels_relabel(got, 3, -1)
    n = gpu_ctx.labels_connectivity(got, teeth, 10)
    gpu_ctx.labels_erode(got, -1, ignore_label=0)
    gpu_ctx.synchronize()
    assert n == n_ref and np.array_equal(got.cpu().numpy(), ref)
    assert n_ref >= 1 and (ref != lab).sum() > 100


def test_host_mirror_keeps_reference_names(gpu_ctx, orc):
    from flyby_b200 import post_process, utils
    P, F = synth.bumpy_sphere(20, 5)
    V = len(P)
    surf = utils.Surf(P, F, {}, {})
    lab = _noisy_labels(P, F, 1)
    work = lab.copy()
    post_process.DilateLabel(surf, work, 2, iterations=2)
    post_process.RemoveIslands(surf, work, 0, 200, ignore_neg1=True)
    post_process.ConnectivityLabeling(surf, work, 2, 10)
    post_process.ErodeLabel(surf, work, 1, iterations=2, target=None)
    ref = orc.pp_dilate(F, V, lab, 2, 2)
    ref = orc.pp_remove_islands(F, V, ref, 0, 200, True)
    ref, _ = orc.pp_connectivity(F, V, ref, 2, 10)
    ref = orc.pp_erode(F, V, ref, 1, iterations=2)
    assert np.array_equal(work, ref)
    with pytest.raises(TypeError):
        post_process.DilateLabel(surf, lab.astype(np.int64), 2)


def test_edge_cases(gpu_ctx, orc):
    P, F = synth.cube()
    _bind(gpu_ctx, P, F)
    lab = np.array([0, 0, 1, 1, 2, 2, 2, 0], np.int32)
    for label in (0, 1, 2, 9):
        got = lab.copy()
        assert gpu_ctx.labels_connectivity(got, label, 3) == orc.pp_connectivity(F, 8, lab, label, 3)[1]
        assert np.array_equal(got, orc.pp_connectivity(F, 8, lab, label, 3)[0])
        assert np.array_equal(gpu_ctx.labels_erode(lab.copy(), label), orc.pp_erode(F, 8, lab, label))
        assert np.array_equal(gpu_ctx.labels_remove_islands(lab.copy(), label, 100, True),
                              orc.pp_remove_islands(F, 8, lab, label, 100, True))
    # a point no cell uses keeps its label and forms its own region
    P2 = np.concatenate([P, [[5, 5, 5]]]).astype(np.float32)
    _bind(gpu_ctx, P2, F)
    lab2 = np.append(lab, 2).astype(np.int32)
    got = lab2.copy()
    n = gpu_ctx.labels_connectivity(got, 2, 10)
    ref, n_ref = orc.pp_connectivity(F, 9, lab2, 2, 10)
    assert n == n_ref == 2 and np.array_equal(got, ref)


def test_randomised_widened_rows(gpu_ctx, orc):
    """A short run of profiles/fuzz_widened.py: pinhole views, barycentric lookup, votes and the label
    post-processing on random meshes (incl. triangle soups with isolated points) and random label fields."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("fuzz_widened", os.path.join(root, "profiles", "fuzz_widened.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    assert fuzz.run(30, 4, context=gpu_ctx) == 0
