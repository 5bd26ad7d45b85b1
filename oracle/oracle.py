"""ctypes front end of the CPU oracle (oracle/flyby_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(fly-by-cnn_b200/) never imports this module.  Parity: pinned to the reference's own
Python arithmetic through tests/golden/ref_*.npz, unpinned for what lives inside VTK --
see the header of flyby_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "flyby_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_grid_build.restype = C.c_void_p
        _LIB.orc_icosahedron.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
        _LIB.orc_tri_intersect.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def set_tolerance(tol):
    """tol of vtkCellLocator::IntersectWithLine (1e-10 in fly_by_features.cxx:750, 1e-8 in predict.py:114)."""
    lib().orc_set_tolerance(C.c_double(tol))


def set_threads(n):
    """OpenMP threads of the ray loops (1 = the reference's serial loop)."""
    lib().orc_set_threads(C.c_int(int(n)))


def num_threads():
    return lib().orc_num_threads()


def unit_surf(verts, centre_mode=0, scale_mode=0, translate=None, scale_factor=None):
    v = _f32(verts)
    out = np.empty_like(v)
    cs = np.zeros(4, dtype=np.float64)
    tr = np.ascontiguousarray(translate, dtype=np.float64) if translate is not None else None
    lib().orc_unit_surf(_p(v), C.c_int64(len(v)), C.c_int(centre_mode), C.c_int(scale_mode),
                        C.c_int(tr is not None), _p(tr), C.c_int(scale_factor is not None),
                        C.c_double(scale_factor or 0.0), _p(out), _p(cs))
    return out, cs[:3].copy(), float(cs[3])


def point_normals(verts, tris):
    v, t = _f32(verts), _i32(tris)
    out = np.empty_like(v)
    lib().orc_point_normals(_p(v), C.c_int64(len(v)), _p(t), C.c_int64(len(t)), _p(out))
    return out


def orient_cells(tris, n_points):
    """vtkPolyDataNormals' consistency pass: returns (tris reoriented, flipped uint8 [T], n_flipped or -1)."""
    t = _i32(tris).copy()
    flipped = np.zeros(len(t), dtype=np.uint8)
    f = lib().orc_orient_cells
    f.restype = C.c_int64
    n = f(_p(t), C.c_int64(len(t)), C.c_int64(int(n_points)), _p(flipped))
    return t, flipped, int(n)


def icosahedron(L, radius, dedupe=0):
    out = np.empty((20 * (L + 1) * (L + 2) // 2, 3), dtype=np.float32)
    n = lib().orc_icosahedron(C.c_int(L), C.c_double(radius), C.c_int(dedupe), _p(out))
    return out[:n].copy()


def spiral(N, turns, radius):
    out = np.empty((N, 3), dtype=np.float32)
    lib().orc_spiral(C.c_int(N), C.c_double(turns), C.c_double(radius), _p(out))
    return out


def view_rays(sp, radius, res, plane_scale=1.0, spacing=1.0):
    sp = _f32(sp)
    rays = np.empty((res * res, 6), dtype=np.float64)
    basis = np.empty((3, 3), dtype=np.float64)
    lib().orc_view_rays(_p(sp), C.c_double(radius), C.c_double(plane_scale), C.c_double(spacing),
                        C.c_int(res), _p(rays), _p(basis))
    return rays, basis


def tri_intersect(o, e, p0, p1, p2):
    o = np.ascontiguousarray(o, dtype=np.float64)
    e = np.ascontiguousarray(e, dtype=np.float64)
    t = C.c_double(0)
    x = np.zeros(3)
    w = np.zeros(3)
    hit = lib().orc_tri_intersect(_p(o), _p(e), _p(_f32(p0)), _p(_f32(p1)), _p(_f32(p2)),
                                  C.byref(t), _p(x), _p(w))
    return bool(hit), t.value, x, w


def cast(verts, tris, rays, grid=True, grid_div=0):
    v, t = _f32(verts), _i32(tris)
    r = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    n = len(r)
    ids = np.empty(n, dtype=np.int32)
    tt = np.empty(n, dtype=np.float64)
    xx = np.empty((n, 3), dtype=np.float64)
    ww = np.empty((n, 3), dtype=np.float64)
    g = None
    if grid:
        g = C.c_void_p(lib().orc_grid_build(_p(v), C.c_int64(len(v)), _p(t), C.c_int64(len(t)),
                                            C.c_int(grid_div)))
    lib().orc_cast(g, _p(v), _p(t), C.c_int64(len(t)), _p(r), C.c_int64(n), _p(ids), _p(tt),
                   _p(xx), _p(ww))
    if g is not None:
        lib().orc_grid_free(g)
    return ids, tt, xx, ww


def channels(use_z, split_z):
    return 4 if (use_z and split_z) else 3


def render(unit_verts, tris, normals, views, radius, res, plane_scale=1.0, spacing=1.0, use_z=1,
           split_z=1, normal_encoding=0, zscale=1.0, grid=True, grid_div=0, want_t=False,
           want_features=True):
    """O5-O8 for a set of float32 viewpoints; returns (features, ids[, t])."""
    v, t, nrm, vw = _f32(unit_verts), _i32(tris), _f32(normals), _f32(views).reshape(-1, 3)
    Cn = channels(use_z, split_z)
    nv = len(vw)
    feats = np.empty((nv, res, res, Cn), dtype=np.float32) if want_features else None
    ids = np.empty((nv, res, res), dtype=np.int32)
    tt = np.empty((nv, res, res), dtype=np.float64) if want_t else None
    lib().orc_render(_p(v), C.c_int64(len(v)), _p(t), C.c_int64(len(t)), _p(nrm), _p(vw),
                     C.c_int(nv), C.c_double(radius), C.c_int(res), C.c_double(plane_scale),
                     C.c_double(spacing), C.c_int(use_z), C.c_int(split_z),
                     C.c_int(normal_encoding), C.c_double(zscale), C.c_int(1 if grid else 0),
                     C.c_int(grid_div), _p(feats), _p(ids), _p(tt))
    return (feats, ids, tt) if want_t else (feats, ids)


def point_features(cell_ids, tris, feats, zero=0.0):
    ids = _i32(cell_ids)
    t = _i32(tris)
    f = _f32(feats).reshape(len(feats), -1)
    out = np.empty(ids.shape + (f.shape[1],), dtype=np.float32)
    lib().orc_point_features(_p(ids), C.c_int64(ids.size), _p(t), _p(f), C.c_int(f.shape[1]),
                             C.c_float(zero), _p(out))
    return out


def backproject(cell_ids, labels, n_cells, K):
    ids, lab = _i32(cell_ids).reshape(-1), _i32(labels).reshape(-1)
    votes = np.zeros((n_cells, K), dtype=np.int32)
    lib().orc_backproject(_p(ids), _p(lab), C.c_int64(ids.size), C.c_int(K), _p(votes))
    return votes


def votes_argmax(votes, default=0):
    v = _i32(votes)
    out = np.empty(len(v), dtype=np.int32)
    lib().orc_votes_argmax(_p(v), C.c_int64(len(v)), C.c_int(v.shape[1]), C.c_int32(default), _p(out))
    return out


def cells_to_points(cell_labels, tris, n_points, default=0):
    cl, t = _i32(cell_labels), _i32(tris)
    out = np.empty(n_points, dtype=np.int32)
    lib().orc_cells_to_points(_p(cl), _p(t), C.c_int64(len(t)), C.c_int64(n_points),
                              C.c_int32(default), _p(out))
    return out


# ---- mesh-label post-processing (oracle/postproc_oracle.c; reference src/py/post_process.py:243-380).
# Each function returns a new label array; the reference edits its vtk array in place.
def pp_remove_islands(tris, n_points, labels, label, min_count, ignore_neg1=False):
    t, out = _i32(tris), _i32(labels).copy()
    lib().orc_pp_remove_islands(_p(t), C.c_int64(len(t)), C.c_int64(n_points), _p(out), C.c_int32(label),
                                C.c_int64(min_count), C.c_int(1 if ignore_neg1 else 0))
    return out


def pp_connectivity(tris, n_points, labels, label, start_label):
    t, out = _i32(tris), _i32(labels).copy()
    f = lib().orc_pp_connectivity
    f.restype = C.c_int64
    n = f(_p(t), C.c_int64(len(t)), C.c_int64(n_points), _p(out), C.c_int32(label), C.c_int32(start_label))
    return out, int(n)


def pp_erode(tris, n_points, labels, label, ignore_label=None, iterations=None, target=None):
    t, out = _i32(tris), _i32(labels).copy()
    f = lib().orc_pp_erode
    f.restype = C.c_int64
    f(_p(t), C.c_int64(len(t)), C.c_int64(n_points), _p(out), C.c_int32(label),
      C.c_int(ignore_label is not None), C.c_int32(ignore_label or 0),
      C.c_int64(-1 if iterations is None else iterations), C.c_int(target is not None), C.c_int32(target or 0))
    return out


def pp_dilate(tris, n_points, labels, label, iterations=2, over_target=False, target=None):
    t, out = _i32(tris), _i32(labels).copy()
    lib().orc_pp_dilate(_p(t), C.c_int64(len(t)), C.c_int64(n_points), _p(out), C.c_int32(label),
                        C.c_int64(iterations), C.c_int(1 if over_target else 0), C.c_int32(target or 0))
    return out


def pp_relabel(labels, label, relabel):
    out = _i32(labels).copy()
    lib().orc_pp_relabel(_p(out), C.c_int64(len(out)), C.c_int32(label), C.c_int32(relabel))
    return out


def interpolate_features(verts, tris, views, radius, res, cell_ids, feats, plane_scale=1.0, spacing=1.0, zero=0.0):
    """Barycentric lookup of per-point features at the hits (fly_by_features.cxx:766-790)."""
    v, t, vw = _f32(verts), _i32(tris), _f32(views).reshape(-1, 3)
    ids = _i32(cell_ids)
    f = _f32(feats).reshape(len(feats), -1)
    out = np.empty(ids.shape + (f.shape[1],), dtype=np.float32)
    lib().orc_interpolate_features(_p(v), _p(t), _p(vw), C.c_int(len(vw)), C.c_double(radius), C.c_int(res),
                                   C.c_double(plane_scale), C.c_double(spacing), _p(ids), _p(f), C.c_int(f.shape[1]),
                                   C.c_float(zero), _p(out))
    return out


def render_perspective(unit_verts, tris, normals, views, res, view_angle=30.0, z_near=0.01, z_far=10.0, use_z=1,
                       split_z=1, normal_encoding=0, zscale=1.0, grid=True, grid_div=0):
    """Pinhole views as FlyByGenerator sets the camera up (fly_by_features.py:85-131); returns (features, ids, depth)."""
    v, t, nrm, vw = _f32(unit_verts), _i32(tris), _f32(normals), _f32(views).reshape(-1, 3)
    nv = len(vw)
    feats = np.empty((nv, res, res, channels(use_z, split_z)), dtype=np.float32)
    ids = np.empty((nv, res, res), dtype=np.int32)
    depth = np.empty((nv, res, res), dtype=np.float64)
    lib().orc_render_perspective(_p(v), C.c_int64(len(v)), _p(t), C.c_int64(len(t)), _p(nrm), _p(vw), C.c_int(nv),
                                 C.c_int(res), C.c_double(view_angle), C.c_double(z_near), C.c_double(z_far),
                                 C.c_int(use_z), C.c_int(split_z), C.c_int(normal_encoding), C.c_double(zscale),
                                 C.c_int(1 if grid else 0), C.c_int(grid_div), _p(feats), _p(ids), _p(depth))
    return feats, ids, depth
