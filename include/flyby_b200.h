/*
 * flyby_b200.h -- C-ABI of the B200-native fly-by feature extractor.
 *
 * The reference (DCBIA-OrthoLab/Fly-by-CNN) has no FFI layer: the path sits
 * behind the Python class FlyByGenerator / CLI of src/py/fly_by_features.py and
 * the C++ tool src/app/fly_by_features.cxx.  Each entry point below names the
 * reference code it replaces (paths relative to the reference repo).  The
 * Python host in fly-by-cnn_b200/flyby_b200/ binds these with ctypes and keeps
 * the reference's class / function / CLI names on top.
 *
 * Conventions
 *  - plain C types only; every call returns a flyby_status (0 = ok) and stores
 *    a message retrievable with flyby_last_error(); nothing throws.
 *  - a context is bound to one CUDA device and one stream; calls on one
 *    context are not thread-safe, different contexts are independent.
 *  - every buffer argument may be HOST or DEVICE memory (detected with
 *    cudaPointerGetAttributes).  Device buffers are used in place and the call
 *    returns without synchronising (work is ordered on the context stream);
 *    host buffers are staged through context-owned device scratch and the call
 *    returns after the copy back completed.  Two internal streams work behind
 *    the context stream and are ordered against it with events: one copies
 *    finished view chunks to the host, one finishes the LBVH of flyby_build
 *    (which the default render path does not read) while the caller goes on;
 *    everything a caller can observe is ordered on the context stream, and
 *    flyby_synchronize waits for all three.
 *  - there is no CPU implementation behind this API: without a CUDA device
 *    flyby_create fails with FLYBY_ERR_CUDA.
 */
#ifndef FLYBY_B200_H
#define FLYBY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLYBY_ABI_VERSION 2

#if defined(__GNUC__)
#define FLYBY_API __attribute__((visibility("default")))
#else
#define FLYBY_API
#endif

typedef enum {
  FLYBY_OK = 0,
  FLYBY_ERR_INVALID = 1, /* bad argument / call order */
  FLYBY_ERR_CUDA = 2,    /* CUDA runtime error (message has the detail) */
  FLYBY_ERR_NCCL = 3,    /* NCCL missing or failed */
  FLYBY_ERR_UNSUPPORTED = 4
} flyby_status;

typedef struct flyby_ctx flyby_ctx;

#define FLYBY_SHADE_TOLERANCE 0x100 /* flag of flyby_render_opts.mode */
#define FLYBY_SHARE_GPU 0x200       /* flag of flyby_render_opts.mode */
#define FLYBY_BUILD_DEFERRED 0x100  /* flag of flyby_norm_opts.consistency */

/* Mesh normalisation: src/py/utils.py:249-290 (ScaleSurf/GetUnitSurf) and
 * src/app/fly_by_features.cxx:393-450. */
typedef struct {
  int32_t centre_mode;   /* 0 bbox mid (utils.py:263-265), 1 centre of mass (cxx:397-403) */
  int32_t scale_mode;    /* 0 1/|bbox_max-centre| (utils.py:281), 1 1/max|p-centre| (cxx:429-435) */
  int32_t has_translate; /* use translate[] as the centre (--translate, utils.py:271) */
  int32_t has_scale;     /* use scale_factor (--scale_factor, utils.py:279) */
  double translate[3];
  double scale_factor;
  int32_t skip;          /* 1 = mesh is already normalised, use the points as given */
  int32_t consistency;   /* winding consistency pass of vtkPolyDataNormals (utils.py:493-501, cxx:472-478):
                            0 check: flyby_build fails with FLYBY_ERR_UNSUPPORTED when two cells traverse a shared
                              edge in the same direction (VTK would reverse cells; nothing is reordered silently);
                            1 fix:   reproduce VTK -- every edge-connected component is oriented like its lowest
                              cell id, reversed cells become (p2,p1,p0) as vtkPolyData::ReverseCell leaves them;
                            2 off:   use the cells as given (ConsistencyOff)
                            | FLYBY_BUILD_DEFERRED: flyby_build does not synchronise with the host.  The checks still
                              run on the device -- out-of-range indices are counted and replaced by 0 before anything
                              gathers through them, duplicate directed edges are flagged -- and their verdict is
                              returned by the next synchronising call on the context (flyby_synchronize, a render or
                              read-back with host buffers, a non-deferred flyby_build): FLYBY_ERR_INVALID /
                              FLYBY_ERR_UNSUPPORTED, after which the mesh counts as not built.  For streams of
                              meshes (batch extraction): the host queues mesh n+1 while mesh n renders. */
} flyby_norm_opts;

/* Render options: src/py/fly_by_features.py:124-150 (channel assembly) and
 * src/app/fly_by_features.xml:62-76 (planeSpacing, planeScaleFactor). */
typedef struct {
  int32_t use_z;           /* --use_z */
  int32_t split_z;         /* --split_z */
  int32_t normal_encoding; /* 0 n*0.5+0.5 (cxx:802-804), 1 *255 (8-bit raster range), 2 signed n (--rescale_features),
                              3 round(255 (n*0.5+0.5)) clamped to 0..255 (the uint8 an RGB read-back returns) */
  int32_t mode;            /* search engine of flyby_render; results are identical bit for bit:
                              0 default (= 1 unless the FLYBY_MODE environment variable names another, for A/B runs),
                              1 raster: triangle-centric scan conversion + exact resolve,
                              2 bvh:    warp-cooperative float32 LBVH packet traversal + exact resolve,
                              3 fused:  warp-cooperative stackless LBVH traversal, exact test at the leaves
                              | FLYBY_SHADE_TOLERANCE (raster engine): depth and normals of pixels whose only candidate is a
                              certain hit are shaded with float32 weights (depth in double) instead of the exact
                              double-precision chain -- hit cell ids stay bit-exact, depth and normals agree with the
                              exact chain within 1e-5 relative (measured: depth equal, normals within 2 float32 ulp);
                              ignored when out_t is requested or barycentric point scalars are set
                              | FLYBY_SHARE_GPU (raster engine): another context renders on the same GPU at the same time
                              (two contexts fed alternately, batch extraction): the exact-resolve kernels are launched as
                              four persistent CTAs per SM -- dispatched at once, with registers to spare -- so that the
                              other context's build and the start of its candidate search run beside them instead of
                              behind their last wave.  Same results; a lone render is ~8 % slower with the flag, two
                              alternating contexts ~10 % faster per mesh */
  double plane_scale;      /* planeScaleFactor, default 1.0 */
  double plane_spacing;    /* planeSpacing, default 1.0 */
  double z_scale;          /* depth multiplier (1/z_far when rescaling), default 1.0 */
  double tolerance;        /* tol of vtkCellLocator::IntersectWithLine: a point within tol of a cell counts as
                              a hit (fly_by_features.cxx:750 uses 1e-10, predict.py:114 1e-8) */
} flyby_render_opts;

/* Per-stage device timings of the last build/render on this context (ms, CUDA
 * events) and traversal counters (when enabled).  ms_normals and ms_tree are
 * measured on the internal side stream, where those two stages run beside the
 * sort and the start of the render: they overlap other work and do not add up
 * to the duration of flyby_build. */
typedef struct {
  float ms_normalise, ms_normals, ms_sort, ms_tree, ms_render;
  int32_t reserved;
  int64_t n_rays, n_hits;
  int64_t nodes_visited, tris_tested; /* filled only by flyby_render with count_work=1 */
  float ms_trace, ms_resolve;         /* the two stages of ms_render (candidate search, exact resolve), summed
                                         over the pixel slabs of the render */
} flyby_stats;

/* ---- lifetime ---------------------------------------------------------- */
/* device >= 0; stream = cudaStream_t (NULL = a stream owned by the context). */
FLYBY_API int flyby_create(int device, void* stream, flyby_ctx** out);
FLYBY_API void flyby_destroy(flyby_ctx* ctx);
FLYBY_API const char* flyby_last_error(const flyby_ctx* ctx);
FLYBY_API int flyby_abi_version(void);
FLYBY_API int flyby_synchronize(flyby_ctx* ctx);

/* ---- mesh: ReadSurf output -> unit surface -> normals -> LBVH ----------- */
/* Replaces GetUnitSurf (utils.py:341) + ComputeNormals (utils.py:493-501) +
 * vtkCellLocator::BuildLocator (fly_by_features.cxx:484-486).
 * verts float32 [V,3], tris int32 [T,3].  flyby_build runs: normalise ->
 * point normals -> Morton codes -> radix sort -> LBVH topology -> AABB refit
 * -> SoA triangle reorder.  All on the GPU. */
FLYBY_API int flyby_set_mesh(flyby_ctx* ctx, const float* verts, int64_t n_verts, const int32_t* tris,
                   int64_t n_tris, const flyby_norm_opts* opts);
FLYBY_API int flyby_build(flyby_ctx* ctx);
/* The source buffers of flyby_set_mesh / flyby_set_point_scalars / flyby_views_custom may be pinned host memory,
 * in which case the copy is asynchronous: keep them valid and unmodified until flyby_build (which synchronises
 * once, for the index validation) or flyby_synchronize has returned. */
/* Per-point scalars that flyby_render writes behind the C feature channels: out_features becomes
 * [views, res, res, C + n_scalars] -- `--point_features <name> --point_features_concat 1` of
 * src/py/fly_by_features.py:334-402 in one pass, without the RGB id-map round trip (utils.py:604-684).
 * feats float32 [V, n_scalars] (host or device; copied).  mode 0: the value at the hit cell's vertex 0, which is
 * what ExtractPointFeaturesClass returns through GetPointIdColors (utils.py:604-621,644-662); mode 1: barycentric
 * interpolation with the exact test's weights, as fly_by_features.cxx:766-793 interpolates its normals.
 * Miss pixels get `zero` (--zero).  n_scalars = 0 removes them; flyby_set_mesh removes them too.  Call after
 * flyby_set_mesh.  Only the orthographic render (flyby_render / flyby_render_async) writes them. */
FLYBY_API int flyby_set_point_scalars(flyby_ctx* ctx, const float* feats, int n_scalars, int mode, float zero);
/* cells reversed by the consistency pass (consistency = 1), 0 otherwise; flipped uint8 [T] optional */
FLYBY_API int flyby_get_winding(flyby_ctx* ctx, int64_t* n_flipped, uint8_t* flipped);
/* Read back what the build produced (parity checks): unit-surf points [V,3],
 * point normals [V,3], centre_scale[4] = centre xyz + scale.  Any may be NULL. */
FLYBY_API int flyby_get_mesh(flyby_ctx* ctx, float* unit_verts, float* normals, double* centre_scale);
/* LBVH introspection for tests: nodes_out [n_nodes,16] 32-bit words (layout in
 * DESIGN.md), slot_cell_ids [T] = original cell id of each Morton-sorted slot. */
FLYBY_API int flyby_get_bvh(flyby_ctx* ctx, int64_t* n_nodes, int32_t* n_top, void* nodes_out,
                  int32_t* slot_cell_ids);

/* ---- viewpoints -------------------------------------------------------- */
/* CreateIcosahedron (utils.py:36-50 + LinearSubdivisionFilter.py:42-67 +
 * normalize_points utils.py:22-31).  dedupe: 0 exact lattice (10 L^2 + 2),
 * 1 float32-equality emulation of the reference's point locator. */
FLYBY_API int flyby_views_icosahedron(flyby_ctx* ctx, int subdivision, double radius, int dedupe);
/* CreateSpiral (utils.py:52-89; fly_by_features.cxx:214-251). */
FLYBY_API int flyby_views_spiral(flyby_ctx* ctx, int n_samples, double turns, double radius);
/* getFlyBy(sphere_points=...) (fly_by_features.py:63-70): float32 [n,3]. */
FLYBY_API int flyby_views_custom(flyby_ctx* ctx, const float* points, int n_views, double radius);
FLYBY_API int flyby_num_views(flyby_ctx* ctx, int* n_views);
FLYBY_API int flyby_get_views(flyby_ctx* ctx, float* points_out /* [n,3] */);

/* ---- ray generation (fly_by_features.cxx:652-702,742-748) --------------- */
/* Writes the segments the render kernel casts: rays_out float64
 * [view_end-view_begin, res, res, 6] = origin xyz, end xyz (i fastest). */
FLYBY_API int flyby_generate_rays(flyby_ctx* ctx, int view_begin, int view_end, int res,
                        const flyby_render_opts* opts, double* rays_out);

/* ---- the hot path: FlyByGenerator.getFlyBy (fly_by_features.py:63-154) on
 * the ray-cast engine of fly_by_features.cxx:742-807 ------------------------
 * out_features float32 [view_end-view_begin, res, res, C], C = 4 when
 * use_z && split_z else 3; out_cell_ids int32 [.., res, res] (-1 = miss);
 * out_t float64 [.., res, res] hit parameter (-1 = miss).  Any may be NULL.
 * Tie-break: smallest t, then lowest cell id. */
FLYBY_API int flyby_render(flyby_ctx* ctx, int view_begin, int view_end, int res,
                 const flyby_render_opts* opts, float* out_features, int32_t* out_cell_ids,
                 double* out_t);
/* Streaming form for host outputs (batch extraction, BASELINE config 3): returns once the work is queued; the
 * render goes to one of two device staging sets and is copied out on a second stream while the next call's
 * mesh is uploaded, built and rendered, so a stream of meshes runs at max(device time, PCIe time).  The host
 * buffers (pinned memory, or the copies serialise) are complete after flyby_synchronize; alternate between two
 * host buffers.  With device outputs it is flyby_render. */
FLYBY_API int flyby_render_async(flyby_ctx* ctx, int view_begin, int view_end, int res,
                       const flyby_render_opts* opts, float* out_features, int32_t* out_cell_ids,
                       double* out_t);
/* Packed output for PCIe-bound consumers of the 8-bit contract (normal_encoding 3 -- the values
 * FlyByGenerator.getFlyBy holds after its RGB read-back, src/py/fly_by_features.py:121-125, are uint8): the three
 * colour channels travel as bytes and the depth as float32, 7 bytes per pixel instead of the 16 of the float tensor,
 * losslessly.  out_rgb uint8 [n, res, res, 3]; out_depth float32 [n, res, res] or NULL (needs use_z && split_z);
 * out_cell_ids optional.  Requires normal_encoding 3, no point scalars, and not (use_z && !split_z).  Host outputs:
 * streaming like flyby_render_async (pinned buffers, complete after flyby_synchronize, alternate two sets); device
 * outputs: in stream order. */
FLYBY_API int flyby_render_packed(flyby_ctx* ctx, int view_begin, int view_end, int res,
                        const flyby_render_opts* opts, uint8_t* out_rgb, float* out_depth, int32_t* out_cell_ids);
/* Raster-compatible views: what FlyByGenerator.getFlyBy reads back from OpenGL
 * (src/py/fly_by_features.py:85-152), produced by the ray caster: a pinhole camera at the viewpoint looking
 * at the origin, view-up (0,0,-1) ((+-1,0,0) exactly at the poles, :93-98), view angle in degrees (vtkCamera
 * default 30), one ray through the centre of every pixel, rows bottom to top as vtkWindowToImageFilter returns
 * them.  Depth is the distance along the viewing direction (the linearised z-buffer of :127-128), clipped to
 * [z_near, z_far]; features follow opts (use_z, split_z, z_scale; normal_encoding 3 = the 8-bit colours of the
 * RGB read-back); plane_scale / plane_spacing are not used.  out_depth float64 (-1 = miss); any output may be NULL. */
FLYBY_API int flyby_render_perspective(flyby_ctx* ctx, int view_begin, int view_end, int res,
                             const flyby_render_opts* opts, double view_angle_deg, double z_near, double z_far,
                             float* out_features, int32_t* out_cell_ids, double* out_depth);
/* Cast caller-supplied segments (rays float64 [n,6]); same outputs per ray
 * plus barycentric weights w [n,3].  vtkCellLocator::IntersectWithLine
 * (fly_by_features.cxx:758, predict.py:128). */
FLYBY_API int flyby_cast(flyby_ctx* ctx, const double* rays, int64_t n_rays, double tolerance,
               int32_t* out_cell_ids, double* out_t, double* out_w);

/* Barycentric lookup of per-point features at the hits of flyby_render (same view range, resolution and
 * options): out[pixel, f] = sum_k w_k feats[point_k(cell), f] with the exact test's weights -- how
 * fly_by_features.cxx:766-790 interpolates its normals, and what the vertex-colour shading of
 * challenge-teeth/teeth_nets.py:121-153 computes.  cell_ids int32 [n_pix] as written by flyby_render;
 * feats float32 [V,F]; out float32 [n_pix,F]; miss -> zero. */
FLYBY_API int flyby_interpolate_features(flyby_ctx* ctx, int view_begin, int view_end, int res,
                               const flyby_render_opts* opts, const int32_t* cell_ids, const float* feats,
                               int n_features, float zero, float* out);

/* ---- point features through the id map (utils.py:604-684) --------------- */
/* mode 0: ids are hit cell ids, value at the cell's vertex 0 (GetPointIdColors
 * utils.py:604-621 colours a cell with its first point's id); mode 1: ids are
 * point ids already (decoded RGB id map, utils.py:644-662).
 * feats float32 [V,F]; out float32 [n_pix,F]; miss / out of range -> zero. */
FLYBY_API int flyby_point_features(flyby_ctx* ctx, const int32_t* cell_ids, int64_t n_pix, const float* feats,
                         int n_feat, float zero, int mode, float* out);

/* The decoded GetPointIdMapActor image (utils.py:604-637, predict_v3.py:34-41,63-67): point id of the hit
 * cell's vertex 0 per pixel, -1 for a miss / out-of-range cell id.  int32 in, int32 out, [n_pix]. */
FLYBY_API int flyby_point_ids(flyby_ctx* ctx, const int32_t* cell_ids, int64_t n_pix, int32_t* out_point_ids);

/* ---- universal-labeling back-projection --------------------------------- */
/* dental_model_seg.py:192-199, predict_v3.py:59-80, predict.py:154-169.
 * votes int32 [n_cells,K] is ACCUMULATED into (caller zeroes it). */
FLYBY_API int flyby_backproject(flyby_ctx* ctx, const int32_t* cell_ids, const int32_t* labels,
                      int64_t n_pix, int n_classes, int32_t* votes);
/* Per-POINT votes of predict_v3.py:59-69 and predict.py:121-157: every hit pixel votes for the vertex 0 of its
 * cell (the point the RGB id map encodes).  votes int32 [V,K], accumulated into.  A miss never votes (the
 * reference's predict_v3 loop lets id -1 index the last row). */
FLYBY_API int flyby_backproject_points(flyby_ctx* ctx, const int32_t* cell_ids, const int32_t* labels,
                             int64_t n_pix, int n_classes, int32_t* votes);
/* Soft votes: probs float32 [n_pix,K] accumulated as round(p * 2^20) into
 * int64 votes_fx [n_cells,K] (integer atomics -> order independent). */
FLYBY_API int flyby_backproject_soft(flyby_ctx* ctx, const int32_t* cell_ids, const float* probs,
                           int64_t n_pix, int n_classes, int64_t* votes_fx);
/* numpy argmax (lowest label wins ties); cells with no vote get dflt. */
FLYBY_API int flyby_votes_argmax(flyby_ctx* ctx, const int32_t* votes, int64_t n_cells, int n_classes,
                       int32_t dflt, int32_t* labels_out);
/* cell label -> the cell's vertex 0, highest cell id wins
 * (dental_model_seg.py:208-211). */
FLYBY_API int flyby_cells_to_points(flyby_ctx* ctx, const int32_t* cell_labels, int32_t dflt,
                          int32_t* point_labels_out /* [V] */);

/* ---- mesh-label post-processing ------------------------ */
/* What every predictor of the reference runs on the point labels right after the
 * back-projection (predict_v3.py:93-133, dental_model_seg.py:228-239).  labels is
 * int32 [V] of the current mesh (flyby_set_mesh + flyby_build), host or device,
 * edited in place; neighbours of a point = the other points of its incident cells. */
/* src/py/post_process.py:285-295 RemoveIslands (+ ConnectedRegion :243-255,
 * NeighborLabel :257-281): connected regions of `label` with fewer than min_count
 * points take the most frequent label around them (ties: the label met first in
 * ascending point id).  As in the reference (:292) nothing is relabelled unless
 * ignore_neg1 is set, and never to -1. */
FLYBY_API int flyby_labels_remove_islands(flyby_ctx* ctx, int32_t* labels, int32_t label,
                                          int64_t min_count, int ignore_neg1);
/* :297-304 ConnectivityLabeling: the connected regions of `label`, in the order of
 * their lowest point id, become start_label, start_label + 1, ...; n_regions
 * (optional) returns their number. */
FLYBY_API int flyby_labels_connectivity(flyby_ctx* ctx, int32_t* labels, int32_t label,
                                        int32_t start_label, int64_t* n_regions);
/* :306-346 ErodeLabel: synchronous sweeps; a point of `label` takes the label of its
 * lowest-id neighbour whose label differs (and is not ignore_label / is target, when
 * has_ignore / has_target); stops when a sweep changes nothing or after `iterations`
 * sweeps (< 0: no limit, the reference's math.inf). */
FLYBY_API int flyby_labels_erode(flyby_ctx* ctx, int32_t* labels, int32_t label, int has_ignore,
                                 int32_t ignore_label, int64_t iterations, int has_target,
                                 int32_t target);
/* :348-375 DilateLabel: `iterations` synchronous sweeps; neighbours of `label` whose
 * label differs (over_target: whose label is exactly target) join it. */
FLYBY_API int flyby_labels_dilate(flyby_ctx* ctx, int32_t* labels, int32_t label, int64_t iterations,
                                  int over_target, int32_t target);
/* :377-380 ReLabel */
FLYBY_API int flyby_labels_relabel(flyby_ctx* ctx, int32_t* labels, int32_t label, int32_t relabel);

/* ---- the one collective: sum of label votes over ranks ------------------ */
/* NCCL is loaded at run time (the libnccl.so.2 torch ships).  Rank 0 calls
 * flyby_nccl_unique_id, the host shares the 128 bytes, every rank calls
 * flyby_nccl_init; flyby_votes_allreduce then runs ncclAllReduce(sum) on the
 * context stream right behind the scatter kernel (no host sync). */
FLYBY_API int flyby_nccl_unique_id(flyby_ctx* ctx, uint8_t id_out[128]);
FLYBY_API int flyby_nccl_init(flyby_ctx* ctx, const uint8_t id[128], int rank, int world);
FLYBY_API int flyby_votes_allreduce(flyby_ctx* ctx, int32_t* votes, int64_t n, int root /* -1 = all */);

/* ---- replay of a step as a CUDA graph ------------------------------------ */
/* Batch extraction (BASELINE configs[2]: 1000 equally shaped meshes) repeats one step -- flyby_set_mesh from the
 * same device buffers, flyby_build, flyby_render into the same device outputs -- some sixty kernel launches of a few
 * microseconds each.  Between flyby_capture_begin and flyby_capture_end those calls are recorded instead of run (the
 * side-stream work becomes parallel branches of the graph); flyby_graph_launch replays the whole step with ONE launch
 * on the context stream, re-reading the source buffers the captured flyby_set_mesh named.  Rules: run the identical
 * step once before capturing it (buffers must have their size: nothing may allocate while capturing); device buffers
 * only; the mesh must be set with FLYBY_BUILD_DEFERRED (no host synchronisation -- a bad mesh is reported by the next
 * synchronising call, as always); flyby_get_stats timings do not cover replays. */
typedef struct flyby_graph flyby_graph;
FLYBY_API int flyby_capture_begin(flyby_ctx* ctx);
FLYBY_API int flyby_capture_end(flyby_ctx* ctx, flyby_graph** out);
FLYBY_API int flyby_graph_launch(flyby_ctx* ctx, flyby_graph* graph);
FLYBY_API void flyby_graph_destroy(flyby_graph* graph);

/* ---- measurement -------------------------------------------------------- */
FLYBY_API int flyby_get_stats(flyby_ctx* ctx, flyby_stats* out);
/* 1: the next flyby_render also counts nodes visited / triangles tested. */
FLYBY_API int flyby_set_count_work(flyby_ctx* ctx, int enable);
/* number of kernels this context has launched since creation */
FLYBY_API int flyby_launch_count(flyby_ctx* ctx, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* FLYBY_B200_H */
