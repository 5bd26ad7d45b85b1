"""Golden vectors computed by the REFERENCE's own code.

Runs in the build container only (needs /root/reference): the reference's pure-Python
modules are imported UNMODIFIED from /root/reference/src/py with the container-only
`vtk` / `itk` / `monai` stand-ins of tests/refshim on sys.path, script files that cannot
be imported (argparse + tensorflow at module level) are executed as *line slices read
from the reference at run time* (nothing of them is copied into this repository), and
the results are written to tests/golden/ref_*.npz.  tests/test_ref_golden.py (CPU) holds
the oracle to these vectors and tests/test_gpu_ref_golden.py the CUDA path.

    python tests/golden/make_ref_golden.py

What this pins (reference arithmetic executed, not restated):
  a1  ScaleSurf / GetUnitSurf                 utils.py:249-290,341-343
  a3  CreateIcosahedron + LinearSubdivisionFilter.GenerateData + normalize_points
                                              utils.py:22-50, LinearSubdivisionFilter.py:23-100
  a4  CreateSpiral                            utils.py:52-89
  a5  view-plane basis / segment end points   predict.py:83-126 (python twin of cxx:652-748)
  a10 GetPointIdColors + ExtractPointFeaturesClass.__call__   utils.py:604-662
  a12 vote loops + argmax                     predict_v3.py:59-80, predict.py:154-169,
                                              dental_model_seg.py:192-211
  f1  post_process.py:243-380 and the predict_v3.py:93-133 sequence
  f3  OFFReader                               readers.py:3-63
What it cannot pin (arithmetic inside VTK: vtkPlatonicSolidSource's table and solid scale,
vtkPlaneSource, vtkPolyDataNormals, vtkCellLocator / vtkTriangle) stays with oracle/pin_vtk.py.
"""
import hashlib
import io
import os
import sys
import textwrap
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src/py"
sys.path[:0] = [os.path.join(ROOT, "tests", "refshim"), REF, ROOT, os.path.join(ROOT, "fly-by-cnn_b200")]

import vtk  # noqa: E402  (tests/refshim/vtk)
from vtk.util.numpy_support import vtk_to_numpy, numpy_to_vtk  # noqa: E402
import utils as ref_utils  # noqa: E402  (/root/reference/src/py/utils.py, unmodified)
import post_process as ref_pp  # noqa: E402
import readers as ref_readers  # noqa: E402
from flyby_b200 import synth  # noqa: E402  (input meshes only)
from oracle import oracle as orc  # noqa: E402  (pixel -> cell maps used as INPUTS of the id / vote code)


def polydata(P, F):
    pts = vtk.vtkPoints()
    pts.arr.data = np.ascontiguousarray(P, dtype=np.float32).copy()
    polys = vtk.vtkCellArray()
    polys.cells = [list(map(int, f)) for f in F]
    s = vtk.vtkPolyData()
    s.SetPoints(pts)
    s.SetPolys(polys)
    return s


def ref_slice(fname, first, last, starts_with):
    """Lines first..last (1-based, inclusive) of a reference script, dedented and compiled."""
    lines = open(os.path.join(REF, fname)).read().split("\n")[first - 1:last]
    assert lines[0].strip().startswith(starts_with), (fname, first, lines[0])
    return compile(textwrap.dedent("\n".join(lines).expandtabs(4)), "%s:%d-%d" % (fname, first, last), "exec")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def int_labels(arr):
    out = vtk.vtkIntArray()
    out.SetNumberOfComponents(1)
    out.SetNumberOfTuples(len(arr))
    out.data[:, 0] = arr
    out.SetName("RegionId")
    return out


# --------------------------------------------------------------------------- a3 / a4
def gen_views(out):
    for L, r in [(1, 2.0), (2, 2.0), (3, 2.0), (4, 2.0), (5, 2.0), (10, 2.0), (3, 2.75), (2, 4.0), (4, 1.1)]:
        ico = ref_utils.CreateIcosahedron(r, L)
        out["ico_L%d_r%g" % (L, r)] = vtk_to_numpy(ico.GetPoints().GetData()).copy()
        if L <= 3:
            out["ico_cells_L%d_r%g" % (L, r)] = np.asarray(ico.GetPolys().cells, dtype=np.int32)
    for R, N, turns in [(4, 64, 4), (2.0, 64, 4), (4, 16, 2), (1.5, 100, 7), (2.0, 1, 4)]:
        sp = ref_utils.CreateSpiral(R, N, turns)
        out["spiral_R%g_N%d_t%d" % (R, N, turns)] = vtk_to_numpy(sp.GetPoints().GetData()).copy()


# --------------------------------------------------------------------------- a1
def gen_unit_surf(out):
    meshes = {
        "bumpy": synth.bumpy_sphere(6, seed=42),
        "crown": synth.dental_crown(8, seed=1),
    }
    rng = np.random.default_rng(3)
    P, F = synth.bumpy_sphere(5, seed=9)
    meshes["shifted"] = ((P * np.float32(37.5) + np.array([120.0, -40.0, 7.25], np.float32)).astype(np.float32), F)
    meshes["aniso"] = ((P * np.array([0.01, 3.0, 250.0], np.float32) + rng.normal(size=3).astype(np.float32)), F)
    for name, (P, F) in meshes.items():
        P = np.ascontiguousarray(P, dtype=np.float32)
        out["us_%s_P" % name], out["us_%s_F" % name] = P, F.astype(np.int32)
        surf, mean, scale = ref_utils.ScaleSurf(polydata(P, F))
        out["us_%s_unit" % name] = vtk_to_numpy(surf.GetPoints().GetData()).copy()     # float64 in the reference
        out["us_%s_mean" % name], out["us_%s_scale" % name] = np.asarray(mean, np.float64), np.float64(scale)
        # GetUnitSurf(surf, mean_arr, scale_factor): --translate / --scale_factor (fly_by_features.py:236-240)
        m2, s2 = np.array([0.125, -0.5, 2.0]), 0.0375
        surf2 = ref_utils.GetUnitSurf(polydata(P, F), mean_arr=m2, scale_factor=s2)
        out["us_%s_unit_given" % name] = vtk_to_numpy(surf2.GetPoints().GetData()).copy()
        out["us_given_mean"], out["us_given_scale"] = m2, np.float64(s2)


# --------------------------------------------------------------------------- a10
def gen_id_codec(out):
    # enough points for all three colour channels: 255^2 = 65 025 < V
    P, F = synth.bumpy_sphere(82, seed=2)          # V = 67 242, T = 134 480
    F = F.astype(np.int32)
    surf = polydata(P, F)
    colours = vtk_to_numpy(ref_utils.GetPointIdColors(surf)).copy()      # uint8 [T,3], per CELL
    rng = np.random.default_rng(11)
    cells = rng.integers(0, len(F), size=(3, 40, 40)).astype(np.int32)
    cells[rng.random(cells.shape) < 0.3] = -1                            # background pixels
    pick = np.concatenate([[0, 1, 254, 255, 256, len(F) - 1], np.argsort(F[:, 0])[[0, 1, -1, -2]]])
    cells.reshape(-1)[:len(pick)] = pick                                 # edges of the code range
    # what the RGB read-back of GetPointIdMapActor holds for these pixels (background = black);
    # int64, not uint8: the reference ran on numpy 1.x where uint8 * 255 promotes to a Python-sized int
    rgb = np.where((cells >= 0)[..., None], colours[np.maximum(cells, 0)], 0).astype(np.int64)
    feats3 = rng.normal(size=(len(P), 3)).astype(np.float32)
    feats1 = rng.normal(size=(len(P),)).astype(np.float32)
    ex3 = ref_utils.ExtractPointFeaturesClass(feats3, np.zeros(3) + -2.0)
    # hits come back as [row], misses as the bare `zero` vector (utils.py:657-660): flatten each element
    f3 = np.array([np.asarray(e, np.float64).reshape(3) for e in ex3(rgb)]).reshape(cells.shape + (3,))
    arr = numpy_to_vtk(feats1)
    arr.SetName("curv")
    surf.GetPointData().AddArray(arr)
    f1 = ref_utils.ExtractPointFeatures(surf, rgb, "curv", zero=0.5, use_multi=False)
    # 3-component features through ExtractPointFeatures itself only work without background pixels
    # (its np.array(feat) is ragged otherwise): first view rendered "full"
    cells_full = np.where(cells[:1] >= 0, cells[:1], cells[:1].max())
    rgb_full = colours[cells_full].astype(np.int64)
    fc = ref_utils.ExtractPointFeatures(surf, rgb_full, "coords", zero=0, use_multi=False)
    # the mesh and the feature arrays are regenerated by the tests (synth.bumpy_sphere(82, seed=2), default_rng(11)
    # in the order used above); their digests guard against drift of the generators
    out.update(idc_digest=np.frombuffer(hashlib.sha256(np.asarray(P, np.float32).tobytes() + F.tobytes() + feats3.tobytes()
                                                       + feats1.tobytes()).digest(), dtype=np.uint8),
               idc_colours_digest=np.frombuffer(hashlib.sha256(colours.tobytes()).digest(), dtype=np.uint8),
               idc_colours_head=colours[:1024].copy(), idc_colours_tail=colours[-1024:].copy(),
               idc_cells=cells, idc_rgb=rgb.astype(np.uint8), idc_out3=f3.astype(np.float64),
               idc_out1=np.asarray(f1, np.float64), idc_out_coords=np.asarray(fc, np.float64), idc_cells_full=cells_full)


# --------------------------------------------------------------------------- f1
def noisy_labels(U, seed, n_labels=4, flip=0.1, lo=-1, hi=5):
    rng = np.random.default_rng(seed)
    lab = np.argmax(U @ rng.normal(size=(n_labels, 3)).T, axis=1).astype(np.int32)
    f = rng.random(len(U)) < flip
    lab[f] = rng.integers(lo, hi, f.sum()).astype(np.int32)
    return lab


def gen_post_process(out):
    cases = {"s6": synth.bumpy_sphere(6, seed=42), "s9": synth.bumpy_sphere(9, seed=4), "crown7": synth.dental_crown(7, seed=1)}
    for name, (P, F) in cases.items():
        F = F.astype(np.int32)
        surf = polydata(P, F)
        lab = noisy_labels(np.asarray(P, np.float64), seed=7 + len(P))
        out["pp_%s_F" % name], out["pp_%s_V" % name], out["pp_%s_in" % name] = F, np.int64(len(P)), lab

        def run(fn, *a, **k):
            arr = int_labels(lab)
            fn(surf, arr, *a, **k)
            return arr.data[:, 0].copy()

        for lbl in (0, 1, 2, 3):
            out["pp_%s_islands_%d" % (name, lbl)] = run(ref_pp.RemoveIslands, lbl, 30, True)
            out["pp_%s_conn_%d" % (name, lbl)] = run(ref_pp.ConnectivityLabeling, lbl, 10)
            out["pp_%s_dilate_%d" % (name, lbl)] = run(ref_pp.DilateLabel, lbl, iterations=2)
            out["pp_%s_erode_%d" % (name, lbl)] = run(ref_pp.ErodeLabel, lbl, ignore_label=(lbl + 1) % 4)
        out["pp_%s_islands_noflag" % name] = run(ref_pp.RemoveIslands, 2, 30)                 # :292 quirk: nothing changes
        out["pp_%s_islands_big" % name] = run(ref_pp.RemoveIslands, 1, 10 ** 6, True)
        out["pp_%s_islands_neg1" % name] = run(ref_pp.RemoveIslands, -1, 50, True)
        out["pp_%s_erode_all" % name] = run(ref_pp.ErodeLabel, -1)
        out["pp_%s_erode_it2_t3" % name] = run(ref_pp.ErodeLabel, 2, iterations=2, target=3)
        out["pp_%s_dilate_over" % name] = run(ref_pp.DilateLabel, 3, iterations=3, dilateOverTarget=True, target=0)
        out["pp_%s_relabel" % name] = run(ref_pp.ReLabel, 3, -1)
        # ConnectedRegion / NeighborLabel on their own (first point of label 2)
        arr = int_labels(lab)
        pid0 = int(np.flatnonzero(lab == 2)[0])
        visited = np.zeros(len(lab))
        region = ref_pp.ConnectedRegion(surf, pid0, arr, 2, visited)
        out["pp_%s_region_pid" % name], out["pp_%s_region" % name] = np.int64(pid0), np.asarray(region, np.int64)
        out["pp_%s_region_nlabel" % name] = np.int64(ref_pp.NeighborLabel(surf, arr, 2, region))
        # the sequence predict_v3.py runs after the vote (dilate gum, islands, relabel, connectivity, erode)
        code = ref_slice("predict_v3.py", 93, 133, "if args.dilate:")

        class Args(object):
            dilate, out = 2, "/tmp/ref_golden_unused.vtk"

        ns = dict(args=Args, post_process=ref_pp, surf=surf, real_labels=int_labels(lab), np=np, vtk=vtk, os=os)
        quiet(exec, code, ns)
        out["pp_%s_predict_v3_sequence" % name] = ns["real_labels"].data[:, 0].copy()


# --------------------------------------------------------------------------- a12
def gen_votes(out):
    P, F = synth.bumpy_sphere(6, seed=42)
    F = F.astype(np.int32)
    U, _, _ = orc.unit_surf(P)
    N = orc.point_normals(U, F)
    views = orc.icosahedron(1, 2.0)
    res = 24
    _, cells = orc.render(U, F, N, views, 2.0, res, plane_scale=2.0, grid=False)     # INPUT: a pixel -> cell map
    rng = np.random.default_rng(5)
    K = 6
    pred = rng.integers(0, K, size=cells.shape).astype(np.int32)
    pred.reshape(-1)[0] = K - 1
    surf = polydata(U, F)
    colours = vtk_to_numpy(ref_utils.GetPointIdColors(surf))
    rgb = np.where((cells >= 0)[..., None], colours[np.maximum(cells, 0)], 0).astype(np.int64)
    # predict_v3.py:59-80 -- per-POINT hard votes through the RGB id map, argmax, default 0
    ns = dict(np=np, vtk=vtk, surf=surf, img_point_id_map_np=rgb.reshape(-1, 3), img_predict_np=pred.reshape(-1))
    exec(ref_slice("predict_v3.py", 59, 80, "prediction_array_count = np.zeros"), ns)
    out.update(vote_F=F, vote_V=np.int64(len(U)), vote_cells=cells, vote_pred=pred, vote_K=np.int64(K),
               v3_counts=ns["prediction_array_count"].astype(np.int64),
               v3_labels=ns["real_labels"].data[:, 0].copy())
    # predict.py:154-169 -- per-point votes from ray-cast ids (-1 = miss skipped), default -1
    pointid = np.where(cells >= 0, F[np.maximum(cells, 0), 0], -1).reshape(len(views), -1)
    label_array = np.zeros([len(U), 4])
    pred4 = (pred % 4)
    vote_code = ref_slice("predict.py", 154, 157, "for index in range(planeResolution*planeResolution):")
    for v in range(len(views)):
        exec(vote_code, dict(planeResolution=res, pointid_array=pointid[v], prediction=pred4[v].reshape(-1), label_array=label_array))
    ns = dict(np=np, vtk=vtk, surf=surf, label_array=label_array)
    exec(ref_slice("predict.py", 160, 169, "real_labels = vtk.vtkIntArray()"), ns)
    out.update(p1_counts=label_array.astype(np.int64), p1_labels=ns["real_labels"].data[:, 0].copy(), p1_pred=pred4.astype(np.int32))
    # dental_model_seg.py:192-211 -- per-CELL soft votes via pix_to_face, argmax, gum default 33, cell -> vertex 0
    Ks = 5
    probs = rng.random(size=(len(views), Ks, res, res)).astype(np.float32)
    probs /= probs.sum(axis=1, keepdims=True)
    array_faces = np.zeros((Ks, len(F)))
    loop = ref_slice("dental_model_seg.py", 192, 194, "for x in range(image_size):")
    for v in range(len(views)):
        exec(loop, dict(image_size=res, array_faces=array_faces, pix_to_face=cells[v], outputs_softmax=probs[v]))
    soft_raw = array_faces.copy()

    class Args(object):
        scal = "PredictedID"

    ns = dict(np=np, array_faces=array_faces, SURF=surf, vtk_to_numpy=vtk_to_numpy, numpy_to_vtk=numpy_to_vtk, args=Args)
    exec(ref_slice("dental_model_seg.py", 196, 215, "array_faces[:,-1][0] = 0"), ns)
    out.update(dms_probs=probs, dms_soft=soft_raw, dms_faces=np.asarray(ns["final_faces_array"], np.int64),
               dms_points=np.asarray(ns["id_points"], np.int64))


# --------------------------------------------------------------------------- a5 (python twin)
def gen_predict_rays(out):
    """predict.py:86-139 with vtkPlaneSource / vtkCellLocator / normals ORACLE-BACKED: pins the code around
    those calls (basis incl. the pole rule of :97-99, plane corners, spacing offset, segment end), not VTK."""
    P, F = synth.bumpy_sphere(5, seed=3)
    F = F.astype(np.int32)
    U, _, _ = orc.unit_surf(P)
    base = polydata(U, F)
    nf = vtk.vtkPolyDataNormals()
    nf.SetInputData(base)
    surf = nf.GetOutput()
    tree = vtk.vtkCellLocator()
    tree.SetDataSet(surf)
    tree.BuildLocator()
    lines = open(os.path.join(REF, "predict.py")).read().split("\n")
    assert lines[85].strip().startswith("sphere_point = icosahedron.GetPoint(icoid)"), lines[85]
    assert lines[138].strip().startswith("features[pid][3] = np.linalg.norm"), lines[138]
    body = compile(textwrap.dedent("\n".join(lines[85:139]).expandtabs(4)), "predict.py:86-139", "exec")
    res, radius, spacing = 12, 2.0, 1.0
    ico = ref_utils.CreateIcosahedron(radius, 1)
    feats, ids, starts, ends = [], [], [], []
    for icoid in range(ico.GetNumberOfPoints()):
        rec = []

        class Tree(object):
            def IntersectWithLine(self, p1, p2, *a):
                rec.append((np.array(p1, np.float64), np.array(p2, np.float64)))
                return tree.IntersectWithLine(p1, p2, *a)

        ns = dict(np=np, vtk=vtk, icosahedron=ico, icoid=icoid, planeSpacing=spacing, planeResolution=res,
                  sphereRadius=radius, normalize_vector=ref_utils.normalize_vector, CreatePlane=ref_utils.CreatePlane,
                  tree=Tree(), surf=surf)
        exec(body, ns)
        feats.append(ns["features"].copy())
        ids.append(ns["pointid_array"].copy())
        starts.append(np.array([r[0] for r in rec]))
        ends.append(np.array([r[1] for r in rec]))
    out.update(pr_U=U, pr_F=F, pr_views=vtk_to_numpy(ico.GetPoints().GetData()).copy(), pr_res=np.int64(res),
               pr_radius=np.float64(radius), pr_features=np.array(feats), pr_pointids=np.array(ids, np.int64),
               pr_starts=np.array(starts), pr_ends=np.array(ends))


# --------------------------------------------------------------------------- f3
def gen_off(out):
    P, F = synth.bumpy_sphere(3, seed=1)
    path = "/tmp/ref_golden_mesh.off"
    with open(path, "w") as f:
        f.write("OFF\n%d %d 0\n" % (len(P), len(F)))
        for p in P:
            f.write("%r %r %r\n" % (float(p[0]), float(p[1]), float(p[2])))
        for t in F:
            f.write("3 %d %d %d\n" % tuple(t))
    rd = ref_readers.OFFReader()
    rd.SetFileName(path)
    rd.Update()
    s = rd.GetOutput()
    out.update(off_text=np.frombuffer(open(path, "rb").read(), dtype=np.uint8),
               off_P=vtk_to_numpy(s.GetPoints().GetData()).copy(), off_F=np.asarray(s.GetPolys().cells, np.int32))
    os.remove(path)


def main():
    groups = {"views": gen_views, "unit_surf": gen_unit_surf, "id_codec": gen_id_codec, "post_process": gen_post_process,
              "votes": gen_votes, "predict_rays": gen_predict_rays, "off": gen_off}
    only = sys.argv[1:]
    for name, fn in groups.items():
        if only and name not in only:
            continue
        out = {}
        fn(out)
        path = os.path.join(os.environ.get("REF_GOLDEN_OUT", HERE), "ref_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%-14s %3d arrays  %8d bytes" % (name, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
