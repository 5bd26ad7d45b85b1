// ctx.h -- internal context of libflyby_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/flyby_b200.h"
#include "flyby_core.cuh"

namespace fb {

// set while a context of this thread captures its stream into a CUDA graph (flyby_capture_begin): allocations are
// not allowed then, the buffers must have their size from an identical warm-up step
inline bool& capture_flag() {
  static thread_local bool f = false;
  return f;
}

// grow-only device allocation
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (capture_flag()) return cudaErrorStreamCaptureUnsupported;  // run the same step once before capturing it
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct MeshDev {
  int64_t V = 0, T = 0;
  DevBuf raw_verts;   // float [V,3] as given
  DevBuf tris;        // int32 [T,3] original order
  DevBuf verts;       // float [V,3] unit surface
  DevBuf normals;     // float4 [V] point normals (w unused)
  DevBuf face_n;      // float [T,3] unit face normals (float32, as vtkPolyDataNormals stores them)
  DevBuf adj_start;   // int32 [V+1] CSR of incident cells
  DevBuf adj_cur;     // int32 [V]
  DevBuf adj;         // int32 [3T]
  DevBuf reduce;      // small: bbox / centre / scale scratch
  // LBVH
  DevBuf codes, codes_alt, idx, idx_alt;  // uint32 [T] radix sort ping-pong
  DevBuf hist;                            // uint32 [256 * nblocks]
  DevBuf scan_tmp;                        // scan block totals
  DevBuf tree_scan_tmp;                   // the same for the scan of the tree build (runs on the side stream)
  DevBuf leaf_box;                        // float [T,6] boxes of sorted slots
  DevBuf node_box;                        // float [T-1,6] internal boxes (Karras numbering)
  DevBuf knode;                           // int32 [T-1,4] child0, child1, parent, flag (Karras numbering)
  DevBuf krange;                          // int2 [T-1] first slot, triangle count of every subtree
  DevBuf remap;                           // int32 [T-1] Karras index -> final index
  DevBuf topflag;                         // int32 [T-1]
  DevBuf nodes;                           // Node [max(T-1,1)] final layout, top-of-tree first
  DevBuf tv0, tv1, tv2;                   // float4 [T] sorted vertices; tv0.w = cell id bits
  DevBuf tvid;                            // int4 [T] sorted (pid0, pid1, pid2, cell id)
  DevBuf trec;                            // float4 [T,4] 64-byte record per slot: v0|cell, v1|pid0, v2|pid1, pid2
  int32_t n_nodes = 0;
  int32_t* n_top_dev = nullptr;  // device word: nodes in the breadth-first top block
  double centre_scale[4] = {0, 0, 0, 1};
  bool has_mesh = false, built = false;
  bool tris_ok = false;   // every index of tris lies in [0, V): checked once per flyby_set_mesh (api.cu ensure_tris_checked)
  // winding consistency pass (winding.cu): bytes [T], 1 = the cell was reversed to (p2,p1,p0); empty when nothing flipped
  DevBuf flip, edge_tab, orient;
  int64_t n_flipped = 0;
  // per-point scalars written by the render behind the feature channels (flyby_set_point_scalars)
  DevBuf scalars;        // float [V, n_scalars]
  int n_scalars = 0, scalar_mode = 0;
  float scalar_zero = 0.f;
};

}  // namespace fb

struct flyby_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;  // D2H of finished view chunks overlaps the next chunk's render
  // Two parts of flyby_build are not read by the candidate search of the default render path: the point normals
  // (+ the adjacency CSR behind them; first needed by the shading of the resolve kernels) and the LBVH topology /
  // boxes / layout.  They run on a side stream -- normals behind the unit surface (ev_unit), tree behind the sort
  // (ev_sorted) -- concurrently with whatever the main stream does next.  Everything that reads them, the next
  // flyby_set_mesh and the next flyby_build wait for ev_tree first (fb::wait_tree).
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_unit = nullptr, ev_sorted = nullptr, ev_tree = nullptr;
  bool tree_pending = false;
  cudaEvent_t chunk_ev[16] = {nullptr};
  char err[512] = {0};
  flyby_norm_opts norm;
  fb::MeshDev mesh;
  // viewpoints
  fb::DevBuf view_pts;  // float [n,3]
  fb::DevBuf views;     // fb::View [n]
  fb::DevBuf viewsf;    // fb::ViewF [n] float32 screening frames
  fb::DevBuf tc_table;  // double [res] plane parameters i / (res - 1) of the current resolution
  fb::DevBuf cand;      // int4 [n_pix] candidate slots between the candidate search and the resolve kernel
  fb::DevBuf cand2;     // int4 [n_pix] candidate slots 4..7 of the scan-conversion path
  fb::DevBuf pix_lists; // uint32 pixel queues between the resolve kernels
  fb::DevBuf raster_tmp;  // uint64 [cap] front (depth bound, slot) + int32 [cap] extra counts, cap = pixels the allocation holds (raster.cu)
  // After a render the side stream re-initialises the pixels it used, beside whatever the main stream does next
  // (upload, unit surface and sort of the next mesh): pixels [0, front_clean) hold "no front, no extras" once
  // ev_front fires.  Only raster.cu touches raster_tmp.
  int64_t front_clean = 0;
  cudaEvent_t ev_front = nullptr, ev_front_free = nullptr;
  fb::DevBuf raster_big;  // queue of big triangles (slot, view, first band, rows per band) for raster_big_kernel
  fb::DevBuf raster_ovf;  // spill list of (pixel, slot) pairs beyond 8 extras + per-overflow-pixel records (raster.cu)
  int n_views = 0;
  double radius = 0;
  // staging for host-side buffers
  fb::DevBuf stage_feat, stage_ids, stage_t, stage_in0, stage_in1, stage_out0, stage_misc;
  fb::DevBuf pp_tmp, pp_tmp2;  // label post-processing scratch (postproc.cu)
  // flyby_render_async: two sets of device staging so that the copy-out of call n overlaps the work of call n + 1
  fb::DevBuf async_feat[2], async_ids[2], async_t[2];
  fb::DevBuf pk_feat, pk_rgb[2], pk_z[2];  // flyby_render_packed: float staging of one view chunk, packed staging sets
  cudaEvent_t async_done[2] = {nullptr, nullptr};  // recorded on copy_stream behind the copies of a set
  int async_flip = 0;
  fb::DevBuf counters;  // uint64 [8]: tile counter, work counters
  // stats
  cudaEvent_t ev[8] = {nullptr};
  // candidate search / exact resolve boundaries of every pixel slab of the last render (ms_trace / ms_resolve are sums)
  static const int kMaxSlabEv = 32;
  cudaEvent_t slab_ev[3 * kMaxSlabEv] = {nullptr};
  int n_slab_ev = 0;
  flyby_stats stats;
  int count_work = 0;
  bool deferred_pending = false;
  // CUDA-graph capture of a step (flyby_capture_begin / _end): no timing events, no allocations, no host syncs
  bool capturing = false, cap_forked = false, cap_front_forked = false, cap_has_build = false;
  bool ev_front_live = false;  // ev_front was last recorded outside a capture (only then may it be waited for outside one)
  int64_t cap_launches0 = 0;
  int64_t last_front_pix = 0;  // pixels of the front buffer the last scan-conversion render used  // a FLYBY_BUILD_DEFERRED build's check flags (counters[7]) have not been read yet
  int64_t launches = 0;
  int sm_count = 148;
  int leaf_max = FB_LEAF_MAX;  // triangles per collapsed leaf (FLYBY_LEAF_MAX overrides, 1..8)
  // NCCL (nccl_glue.cu)
  void* nccl_comm = nullptr;
  int nccl_rank = 0, nccl_world = 1;
};

namespace fb {

inline int fail(flyby_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof c->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

#define FB_CUDA(ctx, expr)                                                                     \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return fb::fail((ctx), FLYBY_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,        \
                      cudaGetErrorString(_e));                                                 \
  } while (0)

#define FB_LAUNCH_CHECK(ctx)                                                                   \
  do {                                                                                         \
    (ctx)->launches++;                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return fb::fail((ctx), FLYBY_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,    \
                      cudaGetErrorString(_e));                                                 \
  } while (0)

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// timing events of flyby_get_stats: skipped inside a captured step
#define FB_STAT_EVENT(ctx, ev, st)                      \
  do {                                                  \
    if (!(ctx)->capturing) FB_CUDA((ctx), cudaEventRecord((ev), (st))); \
  } while (0)
// the side stream waits for an event of the main stream (a fork, when capturing)
#define FB_SIDE_WAIT(ctx, ev)                                                  \
  do {                                                                         \
    FB_CUDA((ctx), cudaStreamWaitEvent((ctx)->side_stream, (ev), 0));          \
    if ((ctx)->capturing) (ctx)->cap_forked = true;                            \
  } while (0)

// stage boundaries of one pixel slab of a render: which = 0 start of the candidate search, 1 start of the exact
// resolve, 2 end (advances to the next slab).  flyby_get_stats sums the two intervals over the slabs.
inline int slab_mark(flyby_ctx* c, int which) {
  if (c->capturing) return FLYBY_OK;  // stats are not part of a captured step
  const int k = c->n_slab_ev;
  if (k < flyby_ctx::kMaxSlabEv) {
    cudaEvent_t& e = c->slab_ev[3 * k + which];
    if (!e && cudaEventCreate(&e) != cudaSuccess) return fail(c, FLYBY_ERR_CUDA, "cudaEventCreate failed");
    cudaError_t err = cudaEventRecord(e, c->stream);
    if (err != cudaSuccess) return fail(c, FLYBY_ERR_CUDA, "cudaEventRecord: %s", cudaGetErrorString(err));
  }
  if (which == 2) c->n_slab_ev = k + 1;
  return FLYBY_OK;
}

// stage launchers (each file implements its own)
int prep_run(flyby_ctx* c);                       // prep.cu: unit surf + normals
int lbvh_run(flyby_ctx* c);                       // lbvh.cu: sort + tree + layout
int exclusive_scan_u32(flyby_ctx* c, uint32_t* data, int64_t n, DevBuf& tmp, cudaStream_t st = nullptr);  // lbvh.cu (st: default c->stream)
int wait_tree(flyby_ctx* c);  // lbvh.cu: make the main stream wait for the side-stream tree build
int views_setup(flyby_ctx* c, const flyby_render_opts* o, int res);                      // trace.cu
int trace_render(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, float* feat,
                 int32_t* ids, double* tt, bool first_chunk = true, bool last_chunk = true);                                     // trace.cu
int pack_rgb8(flyby_ctx* c, const float* feat, int channels, int64_t n_pix, uint8_t* rgb, float* depth, int vec);  // trace.cu
int trace_render_perspective(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, double tan_half,
                             double z_near, double z_far, float* feat, int32_t* ids, double* depth);
int trace_interpolate(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, const int32_t* cell_ids,
                      const float* feats, int F, float zero, float* out);
int trace_rays(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, double* rays);
int trace_cast(flyby_ctx* c, const double* rays, int64_t n, double tol, int32_t* ids, double* tt, double* w);
// postproc.cu: labels int32 [V] on the device, edited in place
int pp_remove_islands(flyby_ctx* c, int32_t* labels, int32_t label, int64_t min_count, int ignore_neg1);
int pp_connectivity(flyby_ctx* c, int32_t* labels, int32_t label, int32_t start_label, int64_t* n_regions);
int pp_erode(flyby_ctx* c, int32_t* labels, int32_t label, int has_ignore, int32_t ignore_label, int64_t iterations,
             int has_target, int32_t target);
int pp_dilate(flyby_ctx* c, int32_t* labels, int32_t label, int64_t iterations, int over_target, int32_t target);
int pp_relabel(flyby_ctx* c, int32_t* labels, int32_t label, int32_t relabel);
int raster_capture_prologue(flyby_ctx* c);                                      // raster.cu
int views_icosahedron(flyby_ctx* c, int L, double radius, int dedupe);          // views.cu
int views_spiral(flyby_ctx* c, int n, double turns, double radius);             // views.cu

}  // namespace fb
