#!/usr/bin/env python
"""Pin the oracle against the real VTK (TEST INFRASTRUCTURE, like everything under oracle/).

NOT RUN in the build environment of this repository: vtk is not installed there and cannot be installed (no network),
which is why DESIGN.md says "parity unpinned".  On a machine with the reference's pin (`vtk==9.2.2`,
src/py/challenge-teeth/flyby.yml:93)

    python oracle/pin_vtk.py            # exit code 0: every comparison agrees; writes oracle/_ref/pin_vtk_report.json

runs the reference's own calls on synthetic meshes and compares them, value by value, with oracle/flyby_oracle.c:

  * vtkPlatonicSolidSource icosahedron  (utils.py:36-50)            <-> orc.icosahedron(1, r)   (as a point set)
  * vtkPolyDataNormals, splitting off   (utils.py:493-501)          <-> orc.point_normals
  * vtkPlaneSource points               (fly_by_features.cxx:696-702, predict.py:107) <-> orc.view_rays origins
  * vtkCellLocator::IntersectWithLine   (fly_by_features.cxx:750-758 tol 1e-10; predict.py:114-128 tol 1e-8)
                                                                    <-> orc.cast: cell id, t, x, and the weights of
                                                                        vtkTriangle::EvaluatePosition on the hit cell
Disagreements are counted per category and the first few id mismatches are listed.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fly-by-cnn_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def polydata(vtk, U, F):
    from vtk.util.numpy_support import numpy_to_vtk, numpy_to_vtkIdTypeArray
    pts = vtk.vtkPoints()
    pts.SetData(numpy_to_vtk(np.ascontiguousarray(U, np.float32), deep=True))  # float32 points, as ReadSurf delivers them
    cells = np.empty((len(F), 4), np.int64)
    cells[:, 0] = 3
    cells[:, 1:] = F
    ca = vtk.vtkCellArray()
    ca.SetCells(len(F), numpy_to_vtkIdTypeArray(cells.ravel(), deep=True))
    pd = vtk.vtkPolyData()
    pd.SetPoints(pts)
    pd.SetPolys(ca)
    return pd


def main():
    try:
        import vtk
    except ImportError:
        print("pin_vtk: vtk is not importable here -- nothing pinned (parity stays 'unpinned')")
        return 2
    from flyby_b200 import synth
    from oracle import oracle as orc
    from vtk.util.numpy_support import vtk_to_numpy
    ref = getattr(vtk, "reference", None) or vtk.mutable
    report = {"vtk_version": vtk.vtkVersion.GetVTKVersion(), "cases": []}
    bad_total = 0

    # ---- icosahedron of the viewpoints
    ico = vtk.vtkPlatonicSolidSource()
    ico.SetSolidTypeToIcosahedron()
    ico.Update()
    vp = vtk_to_numpy(ico.GetOutput().GetPoints().GetData()).astype(np.float64)
    vp = vp / np.linalg.norm(vp, axis=1, keepdims=True) * 2.0
    mine = orc.icosahedron(1, 2.0).astype(np.float64)
    d = np.abs(vp[:, None, :] - mine[None, :, :]).max(axis=2)
    ico_err = float(d.min(axis=1).max())
    same_order = bool(len(vp) == len(mine) and np.abs(vp - mine).max() < 1e-6)
    report["icosahedron"] = {"points": int(len(vp)), "max_nearest_distance": ico_err, "same_order": same_order}
    bad_total += int(ico_err > 1e-6)

    meshes = {"bumpy12": synth.bumpy_sphere(12, 0), "crown20": synth.dental_crown(20, 1), "cube": synth.cube()}
    for name, (P, F) in meshes.items():
        U, _, _ = orc.unit_surf(P)
        pd = polydata(vtk, U, F)
        case = {"mesh": name, "triangles": int(len(F))}
        # ---- point normals
        pn = vtk.vtkPolyDataNormals()
        pn.SetInputData(pd)
        pn.ComputeCellNormalsOff()
        pn.SplittingOff()
        pn.Update()
        nv = vtk_to_numpy(pn.GetOutput().GetPointData().GetArray("Normals"))
        no = orc.point_normals(U, F)
        case["normals_bit_equal"] = bool(np.array_equal(nv, no))
        case["normals_max_abs_diff"] = float(np.abs(nv.astype(np.float64) - no).max())
        bad_total += int(case["normals_max_abs_diff"] > 1e-6)
        # ---- rays of three views (south pole, a generic one, north pole) and the locator
        loc = vtk.vtkCellLocator()
        loc.SetDataSet(pd)
        loc.BuildLocator()
        views = orc.icosahedron(1, 2.0)[[0, 3, 11]]
        res = 48
        for tol in (1e-10, 1e-8):
            orc.set_tolerance(tol)
            n_id = n_t = n_w = n_rays = n_plane = 0
            examples = []
            for sp in views:
                rays, basis = orc.view_rays(sp, 2.0, res, plane_scale=2.0)
                # the plane points the reference would cast from (vtkPlaneSource, float32)
                bx, by = basis[1], basis[2]  # orc_view_rays: rows n, x, y
                o0 = sp.astype(np.float64) - bx - by  # sp - (s/2) x - (s/2) y with plane scale s = 2
                ps = vtk.vtkPlaneSource()
                ps.SetOrigin(*o0)
                ps.SetPoint1(*(o0 + 2.0 * bx))
                ps.SetPoint2(*(o0 + 2.0 * by))
                ps.SetResolution(res - 1, res - 1)
                ps.Update()
                pp = vtk_to_numpy(ps.GetOutput().GetPoints().GetData()).astype(np.float64)
                n_plane += int((np.abs(pp - rays[:, :3]) > 0).any(axis=1).sum())
                ids, tt, xx, ww = orc.cast(U, F, rays, grid=False)
                for k in range(len(rays)):
                    t, x, pc, sub, cid = ref(0.0), [0.0, 0.0, 0.0], [0.0, 0.0, 0.0], ref(0), ref(-1)
                    hit = loc.IntersectWithLine(rays[k, :3].tolist(), rays[k, 3:].tolist(), tol, t, x, pc, sub, cid)
                    vid = int(cid.get()) if hit else -1
                    n_rays += 1
                    if vid != int(ids[k]):
                        n_id += 1
                        if len(examples) < 5:
                            examples.append({"ray": k, "vtk": vid, "oracle": int(ids[k]), "t_vtk": float(t.get()),
                                             "t_oracle": float(tt[k])})
                        continue
                    if vid < 0:
                        continue
                    if float(t.get()) != float(tt[k]):
                        n_t += 1
                    # weights: vtkTriangle pcoords (r, s) -> (1 - r - s, r, s) belong to points 0, 1, 2 of the cell
                    wv = np.array([1.0 - pc[0] - pc[1], pc[0], pc[1]])
                    if np.abs(wv - ww[k]).max() > 1e-12:
                        n_w += 1
            case["tol_%g" % tol] = {"rays": n_rays, "id_mismatches": n_id, "t_not_bit_equal": n_t,
                                    "weights_differ": n_w, "plane_points_differ": n_plane, "examples": examples}
            bad_total += n_id + n_plane
        report["cases"].append(case)
    orc.set_tolerance(1e-10)
    out_dir = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "pin_vtk_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))
    print("pin_vtk: %s" % ("oracle PINNED against vtk %s" % report["vtk_version"] if bad_total == 0
                           else "%d disagreements -- see the report" % bad_total))
    return 0 if bad_total == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
