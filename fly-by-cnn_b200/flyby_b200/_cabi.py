"""ctypes binding of libflyby_b200.so (include/flyby_b200.h).

The library is the product: every compute call below runs hand-written sm_100a
CUDA kernels.  There is no CPU fallback -- if the shared library is missing or
no CUDA device is present, creating a :class:`Context` raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libflyby_b200.so")
ABI_VERSION = 2

_lib = None


class FlyByError(RuntimeError):
    """status = the flyby_status the C call returned (0 when the error was raised by this layer)."""

    def __init__(self, msg, status=0):
        RuntimeError.__init__(self, msg)
        self.status = status


ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_UNSUPPORTED = 1, 2, 3, 4


class NormOpts(C.Structure):
    _fields_ = [("centre_mode", C.c_int32), ("scale_mode", C.c_int32), ("has_translate", C.c_int32),
                ("has_scale", C.c_int32), ("translate", C.c_double * 3), ("scale_factor", C.c_double),
                ("skip", C.c_int32), ("consistency", C.c_int32)]


class RenderOpts(C.Structure):
    _fields_ = [("use_z", C.c_int32), ("split_z", C.c_int32), ("normal_encoding", C.c_int32),
                ("mode", C.c_int32), ("plane_scale", C.c_double), ("plane_spacing", C.c_double),
                ("z_scale", C.c_double), ("tolerance", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("ms_normalise", C.c_float), ("ms_normals", C.c_float), ("ms_sort", C.c_float),
                ("ms_tree", C.c_float), ("ms_render", C.c_float), ("reserved", C.c_int32),
                ("n_rays", C.c_int64), ("n_hits", C.c_int64), ("nodes_visited", C.c_int64),
                ("tris_tested", C.c_int64), ("ms_trace", C.c_float), ("ms_resolve", C.c_float)]


# every symbol include/flyby_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "flyby_create", "flyby_destroy", "flyby_last_error", "flyby_abi_version", "flyby_synchronize",
    "flyby_set_mesh", "flyby_build", "flyby_get_mesh", "flyby_get_bvh", "flyby_views_icosahedron",
    "flyby_views_spiral", "flyby_views_custom", "flyby_num_views", "flyby_get_views",
    "flyby_generate_rays", "flyby_render", "flyby_cast", "flyby_point_features", "flyby_backproject",
    "flyby_backproject_soft", "flyby_votes_argmax", "flyby_cells_to_points", "flyby_nccl_unique_id",
    "flyby_nccl_init", "flyby_votes_allreduce", "flyby_get_stats", "flyby_set_count_work",
    "flyby_launch_count", "flyby_labels_remove_islands", "flyby_labels_connectivity", "flyby_labels_erode",
    "flyby_labels_dilate", "flyby_labels_relabel", "flyby_interpolate_features", "flyby_render_perspective",
    "flyby_render_async", "flyby_set_point_scalars", "flyby_get_winding", "flyby_point_ids",
    "flyby_backproject_points", "flyby_render_packed", "flyby_capture_begin", "flyby_capture_end", "flyby_graph_launch", "flyby_graph_destroy",
]

NORMAL_ENCODINGS = {"unit01": 0, "byte255": 1, "signed": 2, "uint8": 3}
RENDER_MODES = {"default": 0, "raster": 1, "bvh": 2, "fused": 3}
CONSISTENCY = {"check": 0, "fix": 1, "off": 2}
SHADE_TOLERANCE = 0x100
SHARE_GPU = 0x200
BUILD_DEFERRED = 0x100


def load():
    """Load the CUDA library; raises FlyByError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlyByError(
            "libflyby_b200.so is missing (%s). Build it with `python __graft_entry__.py` or "
            "`make -C fly-by-cnn_b200`; there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.flyby_last_error.restype = C.c_char_p
    lib.flyby_last_error.argtypes = [C.c_void_p]
    lib.flyby_destroy.restype = None
    lib.flyby_destroy.argtypes = [C.c_void_p]
    if lib.flyby_abi_version() != ABI_VERSION:
        raise FlyByError("libflyby_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _ptr(x):
    """Raw address of a numpy array / torch tensor / int / None."""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise FlyByError("array must be C-contiguous")
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):  # torch tensor (device or host)
        if not x.is_contiguous():
            raise FlyByError("tensor must be contiguous")
        return C.c_void_p(x.data_ptr())
    raise FlyByError("unsupported buffer type %r" % type(x))


def _check_dtype(x, np_dtype, name):
    if x is None or isinstance(x, int):
        return
    if isinstance(x, np.ndarray):
        ok = x.dtype == np.dtype(np_dtype)
    else:
        ok = str(x.dtype).replace("torch.", "") == np.dtype(np_dtype).name
    if not ok:
        raise FlyByError("%s must have dtype %s, got %s" % (name, np.dtype(np_dtype).name, x.dtype))


class _Capture:
    def __init__(self, ctx):
        self.ctx, self.graph = ctx, None

    def __enter__(self):
        self.ctx._ck(self.ctx._lib.flyby_capture_begin(self.ctx._h))
        return self

    def __exit__(self, et, ev, tb):
        g = C.c_void_p()
        rc = self.ctx._lib.flyby_capture_end(self.ctx._h, C.byref(g))
        if et is None:
            self.ctx._ck(rc)
            self.graph = g
        return False


class Context:
    """One GPU context = one device + one stream + the mesh/BVH/viewpoints it holds."""

    def __init__(self, device=0, stream=None):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.flyby_create(C.c_int(int(device)), C.c_void_p(stream or 0), C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise FlyByError("flyby_create(device=%d) failed with status %d: a CUDA device is "
                             "required, there is no CPU path" % (device, rc))
        self.device = int(device)
        self.n_verts = 0
        self.n_tris = 0
        self.n_scalars = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.flyby_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self._lib.flyby_last_error(self._h)
            raise FlyByError("flyby status %d: %s" % (rc, msg.decode() if msg else ""), status=rc)

    def synchronize(self):
        self._ck(self._lib.flyby_synchronize(self._h))

    # ---- mesh ---------------------------------------------------------------
    def set_mesh(self, verts, tris, centre_mode=0, scale_mode=0, translate=None, scale_factor=None,
                 skip_normalise=False, consistency="check", deferred=False):
        """consistency: what vtkPolyDataNormals' consistency pass means for this mesh -- "check" (build() raises
        FlyByError(status=ERR_UNSUPPORTED) if cells are inconsistently wound), "fix" (reverse cells as VTK does),
        "off" (cells as given).  deferred=True: FLYBY_BUILD_DEFERRED -- build() does not wait for the device; a
        bad mesh is reported by the next synchronising call instead."""
        _check_dtype(verts, np.float32, "verts")
        _check_dtype(tris, np.int32, "tris")
        V = int(verts.shape[0])
        T = int(tris.shape[0])
        o = NormOpts()
        o.centre_mode, o.scale_mode = int(centre_mode), int(scale_mode)
        if translate is not None:
            o.has_translate = 1
            for k in range(3):
                o.translate[k] = float(translate[k])
        if scale_factor is not None:
            o.has_scale = 1
            o.scale_factor = float(scale_factor)
        o.skip = 1 if skip_normalise else 0
        o.consistency = (CONSISTENCY[consistency] if isinstance(consistency, str) else int(consistency)) | (
            BUILD_DEFERRED if deferred else 0)
        self.n_scalars = 0
        self._ck(self._lib.flyby_set_mesh(self._h, _ptr(verts), C.c_int64(V), _ptr(tris), C.c_int64(T),
                                          C.byref(o)))
        self.n_verts, self.n_tris = V, T

    def build(self):
        self._ck(self._lib.flyby_build(self._h))

    def set_point_scalars(self, feats, mode="vertex0", zero=0.0):
        """Per-point scalars [V,S] that render() writes behind the feature channels (None / S = 0 removes them)."""
        if feats is None:
            self._ck(self._lib.flyby_set_point_scalars(self._h, None, C.c_int(0), C.c_int(0), C.c_float(0.0)))
            self.n_scalars = 0
            return
        _check_dtype(feats, np.float32, "feats")
        S = int(feats.shape[1]) if len(feats.shape) > 1 else 1
        if int(feats.shape[0]) != self.n_verts:
            raise FlyByError("point scalars need one row per mesh point (%d), got %d" % (self.n_verts, int(feats.shape[0])))
        m = {"vertex0": 0, "barycentric": 1}[mode] if isinstance(mode, str) else int(mode)
        self._ck(self._lib.flyby_set_point_scalars(self._h, _ptr(feats), C.c_int(S), C.c_int(m), C.c_float(zero)))
        if isinstance(feats, np.ndarray):
            self.synchronize()      # the copy from (possibly pinned) host memory is asynchronous
        self.n_scalars = S

    def get_winding(self):
        """(number of cells the consistency pass reversed, uint8 [T] flags)."""
        n = C.c_int64(0)
        flags = np.zeros(self.n_tris, np.uint8)
        self._ck(self._lib.flyby_get_winding(self._h, C.byref(n), _ptr(flags)))
        return int(n.value), flags

    def get_mesh(self):
        """(unit_verts [V,3] f32, normals [V,3] f32, centre [3], scale) as numpy."""
        uv = np.empty((self.n_verts, 3), np.float32)
        nr = np.empty((self.n_verts, 3), np.float32)
        cs = np.zeros(4, np.float64)
        self._ck(self._lib.flyby_get_mesh(self._h, _ptr(uv), _ptr(nr), _ptr(cs)))
        return uv, nr, cs[:3].copy(), float(cs[3])

    def get_bvh(self):
        """(nodes uint32 [n,16], n_top, slot_cell_ids [T]) -- introspection for tests."""
        n_nodes = C.c_int64(0)
        n_top = C.c_int32(0)
        self._ck(self._lib.flyby_get_bvh(self._h, C.byref(n_nodes), C.byref(n_top), None, None))
        nodes = np.empty((n_nodes.value, 16), np.uint32)
        slots = np.empty(self.n_tris, np.int32)
        self._ck(self._lib.flyby_get_bvh(self._h, C.byref(n_nodes), C.byref(n_top), _ptr(nodes), _ptr(slots)))
        return nodes, int(n_top.value), slots

    # ---- viewpoints ---------------------------------------------------------
    def views_icosahedron(self, subdivision, radius, dedupe=0):
        self._ck(self._lib.flyby_views_icosahedron(self._h, C.c_int(int(subdivision)), C.c_double(radius),
                                                   C.c_int(int(dedupe))))
        return self.num_views()

    def views_spiral(self, n_samples, turns, radius):
        self._ck(self._lib.flyby_views_spiral(self._h, C.c_int(int(n_samples)), C.c_double(float(turns)),
                                              C.c_double(radius)))
        return self.num_views()

    def views_custom(self, points, radius):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        self._ck(self._lib.flyby_views_custom(self._h, _ptr(pts), C.c_int(len(pts)), C.c_double(radius)))
        return len(pts)

    def num_views(self):
        n = C.c_int(0)
        self._ck(self._lib.flyby_num_views(self._h, C.byref(n)))
        return n.value

    def get_views(self):
        out = np.empty((self.num_views(), 3), np.float32)
        self._ck(self._lib.flyby_get_views(self._h, _ptr(out)))
        return out

    # ---- rendering ----------------------------------------------------------
    @staticmethod
    def render_opts(use_z=1, split_z=1, normal_encoding="unit01", plane_scale=1.0, plane_spacing=1.0,
                    z_scale=1.0, tolerance=1e-10, mode="default", shade_tolerance=False, share_gpu=False):
        """mode: "default" | "raster" | "bvh" | "fused" -- the search engine; results are bit-identical.
        shade_tolerance: FLYBY_SHADE_TOLERANCE -- ids stay exact, depth / normals within 1e-5 relative.
        share_gpu: FLYBY_SHARE_GPU -- another context renders on this GPU at the same time (two contexts fed alternately)."""
        o = RenderOpts()
        o.mode = (RENDER_MODES[mode] if isinstance(mode, str) else int(mode)) | (SHADE_TOLERANCE if shade_tolerance else 0) \
            | (SHARE_GPU if share_gpu else 0)
        o.use_z, o.split_z = int(bool(use_z)), int(bool(split_z))
        o.normal_encoding = NORMAL_ENCODINGS[normal_encoding] if isinstance(normal_encoding, str) \
            else int(normal_encoding)
        o.plane_scale, o.plane_spacing, o.z_scale = float(plane_scale), float(plane_spacing), float(z_scale)
        o.tolerance = float(tolerance)
        return o

    @staticmethod
    def channels(opts):
        return 4 if (opts.use_z and opts.split_z) else 3

    def out_channels(self, opts):
        """floats per pixel of flyby_render's feature output: image channels + point scalars."""
        return self.channels(opts) + self.n_scalars

    def generate_rays(self, res, opts, view_begin=0, view_end=None):
        ve = self.num_views() if view_end is None else view_end
        out = np.empty((ve - view_begin, res, res, 6), np.float64)
        self._ck(self._lib.flyby_generate_rays(self._h, C.c_int(view_begin), C.c_int(ve), C.c_int(res),
                                               C.byref(opts), _ptr(out)))
        return out

    def render_into(self, res, opts, features=None, cell_ids=None, t=None, view_begin=0, view_end=None, wait=True):
        """Render into caller buffers (numpy = host staging + sync; torch cuda tensors = in place, async).
        wait=False with host buffers: streaming form (flyby_render_async) -- the buffers are complete after
        synchronize(); alternate between two sets of (pinned) host buffers."""
        ve = self.num_views() if view_end is None else view_end
        _check_dtype(features, np.float32, "features")
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(t, np.float64, "t")
        fn = self._lib.flyby_render if wait else self._lib.flyby_render_async
        self._ck(fn(self._h, C.c_int(view_begin), C.c_int(ve), C.c_int(res), C.byref(opts),
                    _ptr(features), _ptr(cell_ids), _ptr(t)))

    def render_packed_into(self, res, opts, rgb, depth=None, cell_ids=None, view_begin=0, view_end=None):
        """flyby_render_packed: uint8 colours [n,res,res,3] + float32 depth [n,res,res] (normal_encoding "uint8").
        Host buffers: streaming (complete after synchronize()); device tensors: in stream order."""
        ve = self.num_views() if view_end is None else view_end
        _check_dtype(rgb, np.uint8, "rgb")
        _check_dtype(depth, np.float32, "depth")
        _check_dtype(cell_ids, np.int32, "cell_ids")
        self._ck(self._lib.flyby_render_packed(self._h, C.c_int(view_begin), C.c_int(ve), C.c_int(res), C.byref(opts),
                                               _ptr(rgb), _ptr(depth), _ptr(cell_ids)))

    def render(self, res, opts, view_begin=0, view_end=None, want_t=False):
        """Render to new numpy arrays: (features [n,res,res,C], cell_ids [n,res,res][, t])."""
        ve = self.num_views() if view_end is None else view_end
        n = ve - view_begin
        feat = np.empty((n, res, res, self.out_channels(opts)), np.float32)
        ids = np.empty((n, res, res), np.int32)
        tt = np.empty((n, res, res), np.float64) if want_t else None
        self.render_into(res, opts, feat, ids, tt, view_begin, ve)
        return (feat, ids, tt) if want_t else (feat, ids)

    def render_perspective(self, res, opts, view_angle=30.0, z_near=0.01, z_far=None, view_begin=0, view_end=None,
                           features=None, cell_ids=None, depth=None):
        """Pinhole views (raster-compatible): returns (features [n,res,res,C], cell_ids [n,res,res], depth [n,res,res])."""
        ve = self.num_views() if view_end is None else view_end
        n = ve - view_begin
        if features is None:
            features = np.empty((n, res, res, self.channels(opts)), np.float32)
        if cell_ids is None:
            cell_ids = np.empty((n, res, res), np.int32)
        if depth is None:
            depth = np.empty((n, res, res), np.float64)
        _check_dtype(features, np.float32, "features")
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(depth, np.float64, "depth")
        self._ck(self._lib.flyby_render_perspective(self._h, C.c_int(view_begin), C.c_int(ve), C.c_int(res), C.byref(opts),
                                                    C.c_double(view_angle), C.c_double(z_near), C.c_double(z_far),
                                                    _ptr(features), _ptr(cell_ids), _ptr(depth)))
        return features, cell_ids, depth

    def cast(self, rays, tolerance=1e-10):
        r = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        n = len(r)
        ids = np.empty(n, np.int32)
        tt = np.empty(n, np.float64)
        w = np.empty((n, 3), np.float64)
        self._ck(self._lib.flyby_cast(self._h, _ptr(r), C.c_int64(n), C.c_double(tolerance), _ptr(ids), _ptr(tt), _ptr(w)))
        return ids, tt, w

    # ---- features / labels --------------------------------------------------
    def point_features(self, cell_ids, feats, zero=0.0, out=None, mode=0):
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(feats, np.float32, "feats")
        n_pix = int(np.prod(cell_ids.shape))
        F = int(feats.shape[1]) if len(feats.shape) > 1 else 1
        if out is None:
            out = np.empty(tuple(cell_ids.shape) + (F,), np.float32)
        self._ck(self._lib.flyby_point_features(self._h, _ptr(cell_ids), C.c_int64(n_pix), _ptr(feats),
                                                C.c_int(F), C.c_float(zero), C.c_int(int(mode)), _ptr(out)))
        return out

    def point_ids(self, cell_ids, out=None):
        """Vertex 0 of every hit cell (-1 = background): the decoded point-id map of GetPointIdMapActor."""
        _check_dtype(cell_ids, np.int32, "cell_ids")
        if out is None:
            out = np.empty(tuple(cell_ids.shape), np.int32)
        _check_dtype(out, np.int32, "out")
        self._ck(self._lib.flyby_point_ids(self._h, _ptr(cell_ids), C.c_int64(int(np.prod(cell_ids.shape))), _ptr(out)))
        return out

    def backproject_points(self, cell_ids, labels, n_classes, votes=None):
        """Per-point votes [V,K]: every hit pixel votes for vertex 0 of its cell (predict_v3.py:59-69)."""
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(labels, np.int32, "labels")
        if votes is None:
            votes = np.zeros((self.n_verts, n_classes), np.int32)
        _check_dtype(votes, np.int32, "votes")
        self._ck(self._lib.flyby_backproject_points(self._h, _ptr(cell_ids), _ptr(labels),
                                                    C.c_int64(int(np.prod(cell_ids.shape))), C.c_int(n_classes), _ptr(votes)))
        return votes

    def interpolate_features(self, res, opts, cell_ids, feats, zero=0.0, out=None, view_begin=0, view_end=None):
        """Barycentric lookup of per-point features at the hits of render() (same res / opts / view range)."""
        ve = self.num_views() if view_end is None else view_end
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(feats, np.float32, "feats")
        F = int(feats.shape[1]) if len(feats.shape) > 1 else 1
        if out is None:
            out = np.empty(tuple(cell_ids.shape) + (F,), np.float32)
        _check_dtype(out, np.float32, "out")
        self._ck(self._lib.flyby_interpolate_features(self._h, C.c_int(view_begin), C.c_int(ve), C.c_int(res),
                                                      C.byref(opts), _ptr(cell_ids), _ptr(feats), C.c_int(F),
                                                      C.c_float(zero), _ptr(out)))
        return out

    def backproject(self, cell_ids, labels, n_classes, votes=None):
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(labels, np.int32, "labels")
        n_pix = int(np.prod(cell_ids.shape))
        if votes is None:
            votes = np.zeros((self.n_tris, n_classes), np.int32)
        _check_dtype(votes, np.int32, "votes")
        self._ck(self._lib.flyby_backproject(self._h, _ptr(cell_ids), _ptr(labels), C.c_int64(n_pix),
                                             C.c_int(n_classes), _ptr(votes)))
        return votes

    def backproject_soft(self, cell_ids, probs, n_classes, votes_fx=None):
        _check_dtype(cell_ids, np.int32, "cell_ids")
        _check_dtype(probs, np.float32, "probs")
        n_pix = int(np.prod(cell_ids.shape))
        if votes_fx is None:
            votes_fx = np.zeros((self.n_tris, n_classes), np.int64)
        _check_dtype(votes_fx, np.int64, "votes_fx")
        self._ck(self._lib.flyby_backproject_soft(self._h, _ptr(cell_ids), _ptr(probs), C.c_int64(n_pix),
                                                  C.c_int(n_classes), _ptr(votes_fx)))
        return votes_fx

    def votes_argmax(self, votes, default=0, out=None):
        _check_dtype(votes, np.int32, "votes")
        n_cells, K = int(votes.shape[0]), int(votes.shape[1])
        if out is None:
            out = np.empty(n_cells, np.int32)
        self._ck(self._lib.flyby_votes_argmax(self._h, _ptr(votes), C.c_int64(n_cells), C.c_int(K),
                                              C.c_int32(default), _ptr(out)))
        return out

    def cells_to_points(self, cell_labels, default=0, out=None):
        _check_dtype(cell_labels, np.int32, "cell_labels")
        if out is None:
            out = np.empty(self.n_verts, np.int32)
        self._ck(self._lib.flyby_cells_to_points(self._h, _ptr(cell_labels), C.c_int32(default), _ptr(out)))
        return out

    # ---- mesh-label post-processing (labels int32 [V], numpy or torch-cuda, edited in place) --------
    def labels_remove_islands(self, labels, label, min_count, ignore_neg1=False):
        _check_dtype(labels, np.int32, "labels")
        self._ck(self._lib.flyby_labels_remove_islands(self._h, _ptr(labels), C.c_int32(label), C.c_int64(min_count),
                                                       C.c_int(1 if ignore_neg1 else 0)))
        return labels

    def labels_connectivity(self, labels, label, start_label):
        _check_dtype(labels, np.int32, "labels")
        n = C.c_int64(0)
        self._ck(self._lib.flyby_labels_connectivity(self._h, _ptr(labels), C.c_int32(label), C.c_int32(start_label),
                                                     C.byref(n)))
        return int(n.value)

    def labels_erode(self, labels, label, ignore_label=None, iterations=None, target=None):
        _check_dtype(labels, np.int32, "labels")
        self._ck(self._lib.flyby_labels_erode(
            self._h, _ptr(labels), C.c_int32(label), C.c_int(ignore_label is not None), C.c_int32(ignore_label or 0),
            C.c_int64(-1 if iterations is None else int(iterations)), C.c_int(target is not None), C.c_int32(target or 0)))
        return labels

    def labels_dilate(self, labels, label, iterations=2, over_target=False, target=None):
        _check_dtype(labels, np.int32, "labels")
        self._ck(self._lib.flyby_labels_dilate(self._h, _ptr(labels), C.c_int32(label), C.c_int64(iterations),
                                               C.c_int(1 if over_target else 0), C.c_int32(target or 0)))
        return labels

    def labels_relabel(self, labels, label, relabel):
        _check_dtype(labels, np.int32, "labels")
        self._ck(self._lib.flyby_labels_relabel(self._h, _ptr(labels), C.c_int32(label), C.c_int32(relabel)))
        return labels

    # ---- collective ---------------------------------------------------------
    def nccl_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._ck(self._lib.flyby_nccl_unique_id(self._h, buf))
        return bytes(buf)

    def nccl_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(self._lib.flyby_nccl_init(self._h, buf, C.c_int(rank), C.c_int(world)))
        self._nccl_ready = True

    def votes_allreduce(self, votes, root=-1):
        n = int(np.prod(votes.shape))
        self._ck(self._lib.flyby_votes_allreduce(self._h, _ptr(votes), C.c_int64(n), C.c_int(root)))

    # ---- CUDA-graph replay of a step ------------------------------------------
    def capture(self):
        """Context manager: the set_mesh(deferred=True) / build / render_into(device tensors) calls inside are recorded
        into a CUDA graph instead of run; `.graph` is then launched with graph_launch().  Run the identical step once
        before capturing it."""
        return _Capture(self)

    def graph_launch(self, graph):
        self._ck(self._lib.flyby_graph_launch(self._h, graph))

    def graph_destroy(self, graph):
        self._lib.flyby_graph_destroy.restype = None
        self._lib.flyby_graph_destroy(graph)

    # ---- measurement --------------------------------------------------------
    def stats(self):
        s = Stats()
        self._ck(self._lib.flyby_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_ if k != "reserved"}

    def set_count_work(self, enable):
        self._ck(self._lib.flyby_set_count_work(self._h, C.c_int(1 if enable else 0)))

    def launch_count(self):
        n = C.c_int64(0)
        self._ck(self._lib.flyby_launch_count(self._h, C.byref(n)))
        return n.value
