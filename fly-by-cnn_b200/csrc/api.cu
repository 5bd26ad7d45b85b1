// api.cu -- the C-ABI of include/flyby_b200.h: context lifetime, host/device
// buffer staging and call ordering.  All compute is in the kernels of
// prep.cu / lbvh.cu / views.cu / trace.cu / label.cu.
#include <algorithm>
#include <cstdint>
#include <new>

#include "ctx.h"

namespace fb {
int label_point_features(flyby_ctx*, const int32_t*, int64_t, const float*, int, float, int, float*);
int label_backproject(flyby_ctx*, const int32_t*, const int32_t*, int64_t, int, int32_t*);
int label_backproject_soft(flyby_ctx*, const int32_t*, const float*, int64_t, int, int64_t*);
int label_argmax(flyby_ctx*, const int32_t*, int64_t, int, int32_t, int32_t*);
int label_cells_to_points(flyby_ctx*, const int32_t*, int32_t, int32_t*);
int nccl_unique_id(flyby_ctx*, uint8_t*);
int nccl_init(flyby_ctx*, const uint8_t*, int, int);
void nccl_destroy(flyby_ctx*);
int nccl_votes_allreduce(flyby_ctx*, int32_t*, int64_t, int);
const double* prep_centre_scale(flyby_ctx* c);
int label_point_ids(flyby_ctx*, const int32_t*, int64_t, int32_t*);
int label_backproject_points(flyby_ctx*, const int32_t*, const int32_t*, int64_t, int, int32_t*);
int winding_detect(flyby_ctx* c, int* flags_dev);  // winding.cu
int winding_classify(flyby_ctx* c, int* flags_dev, int* flags_host);
int winding_fix(flyby_ctx* c, int* flags_dev);

// input: device pointers are used in place, host pointers are copied to `stage`
template <class T>
static int stage_in(flyby_ctx* c, DevBuf& stage, const T* src, size_t count, const T** out) {
  if (!src || count == 0) { *out = src; return FLYBY_OK; }
  if (is_device_ptr(src)) { *out = src; return FLYBY_OK; }
  FB_CUDA(c, stage.reserve(count * sizeof(T)));
  FB_CUDA(c, cudaMemcpyAsync(stage.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  *out = stage.as<T>();
  return FLYBY_OK;
}

// output: device pointers are written in place; host pointers get a device
// twin that finish() copies back
template <class T>
struct StageOut {
  T* user = nullptr;
  T* dev = nullptr;
  size_t count = 0;
  bool host = false;
  int begin(flyby_ctx* c, DevBuf& stage, T* dst, size_t n) {
    user = dst; count = n; dev = dst; host = false;
    if (!dst || n == 0) return FLYBY_OK;
    if (!is_device_ptr(dst)) {
      host = true;
      FB_CUDA(c, stage.reserve(n * sizeof(T)));
      dev = stage.as<T>();
    }
    return FLYBY_OK;
  }
  int finish(flyby_ctx* c, bool* need_sync) {
    if (host && user && count) {
      FB_CUDA(c, cudaMemcpyAsync(user, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
      *need_sync = true;
    }
    return FLYBY_OK;
  }
};

static int sync_if(flyby_ctx* c, bool need) {
  if (need && c->capturing) return fail(c, FLYBY_ERR_INVALID, "a captured step cannot use host buffers");
  if (need) FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FLYBY_OK;
}

// counts the indices outside [0, V) and replaces them by 0 in the context's copy of the cells, so that even a
// build whose verdict is read later (FLYBY_BUILD_DEFERRED) never gathers out of range
__global__ void check_tris_kernel(int32_t* __restrict__ tris, int64_t n, int64_t V, int32_t* bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && (tris[i] < 0 || tris[i] >= V)) {
    atomicAdd(bad, 1);
    tris[i] = 0;
  }
}

}  // namespace fb

using namespace fb;

#define CHECK_CTX(c) \
  if (!(c)) return FLYBY_ERR_INVALID; \
  { cudaError_t _e = cudaSetDevice((c)->device); \
    if (_e != cudaSuccess) return fail((c), FLYBY_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(_e)); }

extern "C" {

int flyby_abi_version(void) { return FLYBY_ABI_VERSION; }

void flyby_destroy(flyby_ctx* c);

int flyby_create(int device, void* stream, flyby_ctx** out) {
  if (!out) return FLYBY_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return FLYBY_ERR_CUDA;  // no CUDA device: there is no CPU path
  }
  if (cudaSetDevice(device) != cudaSuccess) return FLYBY_ERR_CUDA;
  flyby_ctx* c = new (std::nothrow) flyby_ctx();
  if (!c) return FLYBY_ERR_INVALID;
  c->device = device;
  memset(&c->norm, 0, sizeof c->norm);
  memset(&c->stats, 0, sizeof c->stats);
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { flyby_destroy(c); return FLYBY_ERR_CUDA; }
    c->own_stream = true;
  }
  for (int i = 0; i < 8; i++)
    if (cudaEventCreate(&c->ev[i]) != cudaSuccess) { flyby_destroy(c); return FLYBY_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { flyby_destroy(c); return FLYBY_ERR_CUDA; }
  {
    const char* e = getenv("FLYBY_NO_SIDE_TREE");  // debugging: build the tree on the main stream
    if (!(e && e[0] == '1')) {
      int prio_lo = 0, prio_hi = 0;  // the many small tree kernels slip in ahead of the big render grids
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      if (cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ev_unit, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ev_sorted, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ev_tree, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ev_front, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ev_front_free, cudaEventDisableTiming) != cudaSuccess) { flyby_destroy(c); return FLYBY_ERR_CUDA; }
    }
  }
  for (int i = 0; i < 16; i++)
    if (cudaEventCreateWithFlags(&c->chunk_ev[i], cudaEventDisableTiming) != cudaSuccess) { flyby_destroy(c); return FLYBY_ERR_CUDA; }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (const char* e = getenv("FLYBY_LEAF_MAX")) {
    int v = atoi(e);
    if (v >= 1 && v <= 8) c->leaf_max = v;
  }
  if (c->counters.reserve(16 * sizeof(unsigned long long)) != cudaSuccess) { flyby_destroy(c); return FLYBY_ERR_CUDA; }
  *out = c;
  return FLYBY_OK;
}

void flyby_destroy(flyby_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (c->side_stream) cudaStreamSynchronize(c->side_stream);
  nccl_destroy(c);
  MeshDev& m = c->mesh;
  DevBuf* bufs[] = {&m.raw_verts, &m.tris, &m.verts, &m.normals, &m.face_n, &m.adj_start, &m.adj_cur, &m.adj,
                    &m.reduce, &m.codes, &m.codes_alt, &m.idx, &m.idx_alt, &m.hist, &m.scan_tmp, &m.tree_scan_tmp, &m.leaf_box,
                    &m.node_box, &m.knode, &m.krange, &m.remap, &m.topflag, &m.nodes, &m.tv0, &m.tv1, &m.tv2, &m.tvid, &m.trec,
                    &c->view_pts, &c->views, &c->viewsf, &c->tc_table, &c->cand, &c->cand2, &c->pix_lists, &c->raster_tmp, &c->stage_feat, &c->stage_ids, &c->stage_t, &c->stage_in0,
                    &m.flip, &m.edge_tab, &m.orient, &m.scalars, &c->stage_in1, &c->stage_out0, &c->stage_misc, &c->pp_tmp, &c->pp_tmp2, &c->raster_ovf, &c->raster_big, &c->counters,
                    &c->async_feat[0], &c->async_feat[1], &c->async_ids[0], &c->async_ids[1], &c->async_t[0], &c->async_t[1],
                    &c->pk_feat, &c->pk_rgb[0], &c->pk_rgb[1], &c->pk_z[0], &c->pk_z[1]};
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < 8; i++)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (cudaEvent_t e : c->slab_ev)
    if (e) cudaEventDestroy(e);
  for (int i = 0; i < 16; i++)
    if (c->chunk_ev[i]) cudaEventDestroy(c->chunk_ev[i]);
  for (int i = 0; i < 2; i++)
    if (c->async_done[i]) cudaEventDestroy(c->async_done[i]);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_unit) cudaEventDestroy(c->ev_unit);
  if (c->ev_sorted) cudaEventDestroy(c->ev_sorted);
  if (c->ev_tree) cudaEventDestroy(c->ev_tree);
  if (c->ev_front) cudaEventDestroy(c->ev_front);
  if (c->ev_front_free) cudaEventDestroy(c->ev_front_free);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* flyby_last_error(const flyby_ctx* c) { return c ? c->err : "null context"; }

// counters[6] is raised by the render kernel's watchdog (a scheduling bug must surface as an error, never as a
// hung device); counters[7] holds the verdict of the build-time checks (low word: out-of-range indices, high word:
// winding flags), read here for builds that did not synchronise themselves (FLYBY_BUILD_DEFERRED)
static int check_watchdog(flyby_ctx* c) {
  unsigned long long w[2] = {0, 0};
  FB_CUDA(c, cudaMemcpyAsync(w, c->counters.as<unsigned long long>() + 6, 16, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (w[0]) return fail(c, FLYBY_ERR_CUDA, "render watchdog tripped in %llu warps: results are incomplete", w[0]);
  if (c->deferred_pending) {
    c->deferred_pending = false;
    FB_CUDA(c, cudaMemsetAsync(c->counters.as<unsigned long long>() + 7, 0, 8, c->stream));
    const int nbad = (int)(w[1] & 0xffffffffull), wflags = (int)(w[1] >> 32);
    if (nbad) {
      c->mesh.built = false;
      return fail(c, FLYBY_ERR_INVALID, "a deferred build found %d triangle indices outside [0, n_verts): its results are invalid", nbad);
    }
    if (wflags & 8) {
      c->mesh.built = false;
      return fail(c, FLYBY_ERR_UNSUPPORTED, "a deferred build found a directed edge used by two cells (inconsistent winding or a "
                                            "non-manifold edge): build again without FLYBY_BUILD_DEFERRED to classify / fix it");
    }
  }
  return FLYBY_OK;
}

// Entry points that read m.tris without a build (flyby_set_mesh only) validate the indices first, once per mesh.
static int ensure_tris_checked(flyby_ctx* c) {
  MeshDev& m = c->mesh;
  if (m.tris_ok || m.T <= 0) return FLYBY_OK;
  int32_t* bad = reinterpret_cast<int32_t*>(c->counters.as<unsigned long long>() + 7);
  if (c->deferred_pending) { int rcd = check_watchdog(c); if (rcd) return rcd; }
  FB_CUDA(c, cudaMemsetAsync(bad, 0, 8, c->stream));
  check_tris_kernel<<<(unsigned)cdiv(3 * m.T, 256), 256, 0, c->stream>>>(m.tris.as<int32_t>(), 3 * m.T, m.V, bad);
  FB_LAUNCH_CHECK(c);
  int32_t nbad = 0;
  FB_CUDA(c, cudaMemcpyAsync(&nbad, bad, 4, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (nbad) return fail(c, FLYBY_ERR_INVALID, "%d triangle indices outside [0, n_verts)", nbad);
  m.tris_ok = true;
  return FLYBY_OK;
}


int flyby_synchronize(flyby_ctx* c) {
  CHECK_CTX(c);
  int rc = check_watchdog(c);
  if (rc) return rc;
  FB_CUDA(c, cudaStreamSynchronize(c->copy_stream));  // copy-outs of flyby_render_async
  if (c->side_stream) FB_CUDA(c, cudaStreamSynchronize(c->side_stream));  // tree build behind the last flyby_build
  return FLYBY_OK;
}

int flyby_set_mesh(flyby_ctx* c, const float* verts, int64_t V, const int32_t* tris, int64_t T,
                   const flyby_norm_opts* opts) {
  CHECK_CTX(c);
  if (!verts || V <= 0 || T < 0 || (T > 0 && !tris)) return fail(c, FLYBY_ERR_INVALID, "set_mesh: empty or null mesh");
  if (V >= (int64_t)1 << 31 || T > FB_LEAF_FIRST_MASK) return fail(c, FLYBY_ERR_INVALID, "set_mesh: mesh too large");
  MeshDev& m = c->mesh;
  { int rc = wait_tree(c); if (rc) return rc; }  // the previous build's side-stream work still reads the triangles
  m.has_mesh = false;
  m.built = false;
  m.tris_ok = false;
  m.n_scalars = 0;
  m.n_flipped = 0;
  if (opts) c->norm = *opts; else memset(&c->norm, 0, sizeof c->norm);
  if ((c->norm.consistency & 0xff) > 2 || (c->norm.consistency & ~(0xff | FLYBY_BUILD_DEFERRED)))
    return fail(c, FLYBY_ERR_INVALID, "set_mesh: consistency must be 0, 1 or 2 (| FLYBY_BUILD_DEFERRED)");
  FB_CUDA(c, m.raw_verts.reserve(sizeof(float) * 3 * (size_t)V));
  FB_CUDA(c, m.tris.reserve(sizeof(int32_t) * 3 * (size_t)(T > 0 ? T : 1)));
  FB_CUDA(c, cudaMemcpyAsync(m.raw_verts.p, verts, sizeof(float) * 3 * (size_t)V, cudaMemcpyDefault, c->stream));
  if (T > 0)
    FB_CUDA(c, cudaMemcpyAsync(m.tris.p, tris, sizeof(int32_t) * 3 * (size_t)T, cudaMemcpyDefault, c->stream));
  m.V = V;
  m.T = T;
  m.has_mesh = true;
  return FLYBY_OK;
}

int flyby_build(flyby_ctx* c) {
  CHECK_CTX(c);
  MeshDev& m = c->mesh;
  if (!m.has_mesh) return fail(c, FLYBY_ERR_INVALID, "build: call flyby_set_mesh first");
  {  // side-stream work of the previous build may still read what this build overwrites
    int rc0 = wait_tree(c);
    if (rc0) return rc0;
    c->tree_pending = false;
  }
  // index validation first: an out-of-range index must never reach a gather.  The winding check of
  // vtkPolyDataNormals' consistency pass (winding.cu) shares the one host synchronisation of the build.
  const int cons = c->norm.consistency & 0xff;
  const bool deferred = (c->norm.consistency & FLYBY_BUILD_DEFERRED) != 0;
  if (c->capturing && !deferred) return fail(c, FLYBY_ERR_INVALID, "build: a captured step needs FLYBY_BUILD_DEFERRED (no host synchronisation)");
  int32_t* bad = reinterpret_cast<int32_t*>(c->counters.as<unsigned long long>() + 7);
  if (c->deferred_pending && !deferred) {  // an earlier deferred build's verdict is still unread: read it first
    int rcd = check_watchdog(c);
    if (rcd) return rcd;
  }
  if (!c->deferred_pending && !c->capturing) FB_CUDA(c, cudaMemsetAsync(bad, 0, 8, c->stream));
  if (m.T > 0 && !m.tris_ok) {
    check_tris_kernel<<<(unsigned)cdiv(3 * m.T, 256), 256, 0, c->stream>>>(m.tris.as<int32_t>(), 3 * m.T, m.V, bad);
    FB_LAUNCH_CHECK(c);
  }
  const bool wind = cons != 2 && m.T > 0 && m.n_flipped == 0;
  if (wind) { int rcw = winding_detect(c, bad + 1); if (rcw) return rcw; }
  if (deferred) {
    // no host synchronisation: indices were sanitised on the device, the flags stay in counters[7] until the next
    // synchronising call (flyby_synchronize, a render with host outputs, a non-deferred build) reports them
    c->deferred_pending = true;
    if (c->capturing) c->cap_has_build = true;
    m.tris_ok = true;
  } else {
    int32_t chk[2] = {0, 0};
    FB_CUDA(c, cudaMemcpyAsync(chk, bad, 8, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (chk[0]) return fail(c, FLYBY_ERR_INVALID, "build: %d triangle indices outside [0, n_verts)", chk[0]);
    m.tris_ok = true;
    if (wind && (chk[1] & 8)) {  // a directed edge occurs twice: find out whether a manifold edge is involved
      int rcw = winding_classify(c, bad + 1, &chk[1]);
      if (rcw) return rcw;
    }
    if (wind && (chk[1] & 1)) {
      if (cons == 0)
        return fail(c, FLYBY_ERR_UNSUPPORTED,
                    "build: inconsistently wound cells (two cells traverse a shared edge in the same direction); "
                    "vtkPolyDataNormals would reverse cells -- set flyby_norm_opts.consistency = 1 to reproduce that, 2 to use the cells as given");
      int rcw = winding_fix(c, bad + 1);
      if (rcw) return rcw;
    }
  }
  int rc = prep_run(c);
  if (rc) return rc;
  rc = lbvh_run(c);
  if (rc) return rc;
  if (c->side_stream) {  // normals + tree are complete when this event fires
    FB_CUDA(c, cudaEventRecord(c->ev_tree, c->side_stream));
    c->tree_pending = true;
  }
  m.built = true;
  return FLYBY_OK;
}

int flyby_get_mesh(flyby_ctx* c, float* unit_verts, float* normals, double* centre_scale) {
  CHECK_CTX(c);
  MeshDev& m = c->mesh;
  if (!m.built) return fail(c, FLYBY_ERR_INVALID, "get_mesh: mesh not built");
  { int rc = wait_tree(c); if (rc) return rc; }
  if (unit_verts)
    FB_CUDA(c, cudaMemcpyAsync(unit_verts, m.verts.p, sizeof(float) * 3 * (size_t)m.V, cudaMemcpyDefault, c->stream));
  if (normals)  // device layout is float4 per point
    FB_CUDA(c, cudaMemcpy2DAsync(normals, 12, m.normals.p, 16, 12, (size_t)m.V, cudaMemcpyDefault, c->stream));
  if (centre_scale)
    FB_CUDA(c, cudaMemcpyAsync(centre_scale, prep_centre_scale(c), 4 * sizeof(double), cudaMemcpyDefault, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FLYBY_OK;
}

int flyby_get_bvh(flyby_ctx* c, int64_t* n_nodes, int32_t* n_top, void* nodes_out, int32_t* slot_cell_ids) {
  CHECK_CTX(c);
  MeshDev& m = c->mesh;
  if (!m.built) return fail(c, FLYBY_ERR_INVALID, "get_bvh: mesh not built");
  { int rc = wait_tree(c); if (rc) return rc; }
  if (n_nodes) *n_nodes = m.n_nodes;
  if (n_top) {
    *n_top = 0;
    if (m.n_top_dev) FB_CUDA(c, cudaMemcpyAsync(n_top, m.n_top_dev, 4, cudaMemcpyDefault, c->stream));
  }
  if (nodes_out && m.n_nodes > 0)
    FB_CUDA(c, cudaMemcpyAsync(nodes_out, m.nodes.p, sizeof(Node) * (size_t)m.n_nodes, cudaMemcpyDefault, c->stream));
  if (slot_cell_ids && m.T > 0)  // 4th word of every int4 id record
    FB_CUDA(c, cudaMemcpy2DAsync(slot_cell_ids, 4, reinterpret_cast<char*>(m.tvid.p) + 12, 16, 4, (size_t)m.T,
                                 cudaMemcpyDefault, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FLYBY_OK;
}

int flyby_views_icosahedron(flyby_ctx* c, int L, double radius, int dedupe) {
  CHECK_CTX(c);
  return views_icosahedron(c, L, radius, dedupe);
}

int flyby_views_spiral(flyby_ctx* c, int n, double turns, double radius) {
  CHECK_CTX(c);
  return views_spiral(c, n, turns, radius);
}

int flyby_views_custom(flyby_ctx* c, const float* pts, int n, double radius) {
  CHECK_CTX(c);
  if (!pts || n < 1) return fail(c, FLYBY_ERR_INVALID, "views_custom: no points");
  FB_CUDA(c, c->view_pts.reserve(sizeof(float) * 3 * (size_t)n));
  FB_CUDA(c, cudaMemcpyAsync(c->view_pts.p, pts, sizeof(float) * 3 * (size_t)n, cudaMemcpyDefault, c->stream));
  if (!is_device_ptr(pts)) FB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_views = n;
  c->radius = radius;
  return FLYBY_OK;
}

int flyby_num_views(flyby_ctx* c, int* n) {
  if (!c || !n) return FLYBY_ERR_INVALID;
  *n = c->n_views;
  return FLYBY_OK;
}

int flyby_get_views(flyby_ctx* c, float* out) {
  CHECK_CTX(c);
  if (!out || c->n_views <= 0) return fail(c, FLYBY_ERR_INVALID, "get_views: no viewpoints");
  FB_CUDA(c, cudaMemcpyAsync(out, c->view_pts.p, sizeof(float) * 3 * (size_t)c->n_views, cudaMemcpyDefault, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FLYBY_OK;
}

static int check_render_args(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, bool need_mesh) {
  if (!o) return fail(c, FLYBY_ERR_INVALID, "render: null options");
  if (need_mesh && !c->mesh.built) return fail(c, FLYBY_ERR_INVALID, "render: mesh not built");
  if (c->n_views <= 0) return fail(c, FLYBY_ERR_INVALID, "render: no viewpoints set");
  if (vb < 0 || ve > c->n_views || vb >= ve) return fail(c, FLYBY_ERR_INVALID, "render: bad view range [%d,%d)", vb, ve);
  if (res < 1 || res > 16384) return fail(c, FLYBY_ERR_INVALID, "render: bad resolution %d", res);
  return FLYBY_OK;
}

int flyby_generate_rays(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, double* rays_out) {
  CHECK_CTX(c);
  int rc = check_render_args(c, vb, ve, res, o, false);
  if (rc) return rc;
  if (!rays_out) return fail(c, FLYBY_ERR_INVALID, "generate_rays: null output");
  rc = views_setup(c, o, res);
  if (rc) return rc;
  StageOut<double> so;
  rc = so.begin(c, c->stage_out0, rays_out, (size_t)(ve - vb) * res * res * 6);
  if (rc) return rc;
  rc = trace_rays(c, vb, ve, res, o, so.dev);
  if (rc) return rc;
  bool need = false;
  rc = so.finish(c, &need);
  if (rc) return rc;
  return sync_if(c, need);
}

int flyby_render(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, float* out_features,
                 int32_t* out_cell_ids, double* out_t) {
  CHECK_CTX(c);
  int rc = check_render_args(c, vb, ve, res, o, true);
  if (rc) return rc;
  const size_t per_view = (size_t)res * res;
  const size_t n_pix = (size_t)(ve - vb) * per_view;
  const int channels = ((o->use_z && o->split_z) ? 4 : 3) + c->mesh.n_scalars;  // floats per pixel of out_features
  rc = views_setup(c, o, res);
  if (rc) return rc;
  StageOut<float> sf;
  StageOut<int32_t> si;
  StageOut<double> st;
  if ((rc = sf.begin(c, c->stage_feat, out_features, n_pix * channels))) return rc;
  if ((rc = si.begin(c, c->stage_ids, out_cell_ids, n_pix))) return rc;
  if ((rc = st.begin(c, c->stage_t, out_t, n_pix))) return rc;
  const bool any_host = sf.host || si.host || st.host;
  if (!any_host) return trace_render(c, vb, ve, res, o, sf.dev, si.dev, st.dev);
  if (c->capturing) return fail(c, FLYBY_ERR_INVALID, "render: a captured step needs device outputs");

  // Host outputs: render in view chunks and copy every finished chunk back on the copy
  // stream while the next chunk renders (PCIe D2H overlaps the traversal).
  const int nv = ve - vb;
  int n_chunks = nv < 8 ? 1 : (nv + 11) / 12;
  if (n_chunks > 16) n_chunks = 16;
  // all kernels are queued first so that a pageable destination (synchronous copy) still overlaps
  int bounds[17];
  for (int k = 0; k <= n_chunks; k++) bounds[k] = vb + (int)((long long)nv * k / n_chunks);
  for (int k = 0; k < n_chunks; k++) {
    const size_t off = (size_t)(bounds[k] - vb) * per_view;
    rc = trace_render(c, bounds[k], bounds[k + 1], res, o, sf.dev ? sf.dev + off * channels : nullptr,
                      si.dev ? si.dev + off : nullptr, st.dev ? st.dev + off : nullptr, k == 0, k == n_chunks - 1);
    if (rc) return rc;
    FB_CUDA(c, cudaEventRecord(c->chunk_ev[k], c->stream));
  }
  for (int k = 0; k < n_chunks; k++) {
    const size_t off = (size_t)(bounds[k] - vb) * per_view;
    const size_t cnt = (size_t)(bounds[k + 1] - bounds[k]) * per_view;
    FB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[k], 0));
    if (sf.host && sf.user)
      FB_CUDA(c, cudaMemcpyAsync(sf.user + off * channels, sf.dev + off * channels, cnt * channels * sizeof(float),
                                 cudaMemcpyDeviceToHost, c->copy_stream));
    if (si.host && si.user)
      FB_CUDA(c, cudaMemcpyAsync(si.user + off, si.dev + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
    if (st.host && st.user)
      FB_CUDA(c, cudaMemcpyAsync(st.user + off, st.dev + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  FB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  return check_watchdog(c);
}

// Streaming form of flyby_render for host outputs: returns as soon as the work is queued.  The render goes to
// one of two device staging sets and is copied out chunk by chunk on the copy stream; the next call (next
// mesh) renders into the other set while those copies run, so a stream of meshes is bound by
// max(device time, PCIe time) instead of their sum.  The host buffers are complete after flyby_synchronize
// (or once the call after next has returned its status -- it waits for this set to drain before reusing it).
int flyby_render_async(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, float* out_features,
                       int32_t* out_cell_ids, double* out_t) {
  CHECK_CTX(c);
  const bool any_host = (out_features && !is_device_ptr(out_features)) || (out_cell_ids && !is_device_ptr(out_cell_ids)) ||
                        (out_t && !is_device_ptr(out_t));
  if (!any_host) return flyby_render(c, vb, ve, res, o, out_features, out_cell_ids, out_t);
  if ((out_features && is_device_ptr(out_features)) || (out_cell_ids && is_device_ptr(out_cell_ids)) ||
      (out_t && is_device_ptr(out_t)))
    return fail(c, FLYBY_ERR_INVALID, "render_async: outputs must be all host or all device memory");
  int rc = check_render_args(c, vb, ve, res, o, true);
  if (rc) return rc;
  const size_t per_view = (size_t)res * res;
  const size_t n_pix = (size_t)(ve - vb) * per_view;
  const int channels = ((o->use_z && o->split_z) ? 4 : 3) + c->mesh.n_scalars;  // floats per pixel of out_features
  if ((rc = views_setup(c, o, res))) return rc;
  const int s = c->async_flip;
  c->async_flip ^= 1;
  for (int i = 0; i < 2; i++)
    if (!c->async_done[i]) FB_CUDA(c, cudaEventCreateWithFlags(&c->async_done[i], cudaEventDisableTiming));
  float* df = nullptr;
  int32_t* di = nullptr;
  double* dt = nullptr;
  // growing a staging buffer frees the old one: drain the copies that may still read it first
  if ((out_features && c->async_feat[s].cap < n_pix * channels * sizeof(float)) ||
      (out_cell_ids && c->async_ids[s].cap < n_pix * sizeof(int32_t)) || (out_t && c->async_t[s].cap < n_pix * sizeof(double)))
    FB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  if (out_features) { FB_CUDA(c, c->async_feat[s].reserve(n_pix * channels * sizeof(float))); df = c->async_feat[s].as<float>(); }
  if (out_cell_ids) { FB_CUDA(c, c->async_ids[s].reserve(n_pix * sizeof(int32_t))); di = c->async_ids[s].as<int32_t>(); }
  if (out_t) { FB_CUDA(c, c->async_t[s].reserve(n_pix * sizeof(double))); dt = c->async_t[s].as<double>(); }
  FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->async_done[s], 0));  // the set's previous copy-out (two calls ago)
  const int nv = ve - vb;
  int n_chunks = nv < 8 ? 1 : (nv + 11) / 12;
  if (n_chunks > 16) n_chunks = 16;
  int bounds[17];
  for (int k = 0; k <= n_chunks; k++) bounds[k] = vb + (int)((long long)nv * k / n_chunks);
  for (int k = 0; k < n_chunks; k++) {
    const size_t off = (size_t)(bounds[k] - vb) * per_view;
    const size_t cnt = (size_t)(bounds[k + 1] - bounds[k]) * per_view;
    rc = trace_render(c, bounds[k], bounds[k + 1], res, o, df ? df + off * channels : nullptr, di ? di + off : nullptr,
                      dt ? dt + off : nullptr, k == 0, k == n_chunks - 1);
    if (rc) return rc;
    FB_CUDA(c, cudaEventRecord(c->chunk_ev[k], c->stream));
    FB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[k], 0));
    if (df)
      FB_CUDA(c, cudaMemcpyAsync(out_features + off * channels, df + off * channels, cnt * channels * sizeof(float),
                                 cudaMemcpyDeviceToHost, c->copy_stream));
    if (di) FB_CUDA(c, cudaMemcpyAsync(out_cell_ids + off, di + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
    if (dt) FB_CUDA(c, cudaMemcpyAsync(out_t + off, dt + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  FB_CUDA(c, cudaEventRecord(c->async_done[s], c->copy_stream));
  return FLYBY_OK;
}

// ---- packed output (normal_encoding 3): uint8 colours + float32 depth, 7 bytes per pixel across PCIe instead of 16
int flyby_render_packed(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, uint8_t* out_rgb, float* out_depth,
                        int32_t* out_cell_ids) {
  CHECK_CTX(c);
  int rc = check_render_args(c, vb, ve, res, o, true);
  if (rc) return rc;
  if (!out_rgb) return fail(c, FLYBY_ERR_INVALID, "render_packed: null colour output");
  if (o->normal_encoding != 3)
    return fail(c, FLYBY_ERR_UNSUPPORTED, "render_packed: needs normal_encoding 3 (the 8-bit colours): other encodings are not integers");
  if (o->use_z && !o->split_z)
    return fail(c, FLYBY_ERR_UNSUPPORTED, "render_packed: use_z without split_z multiplies the colours by the depth: not 8-bit values");
  if (c->mesh.n_scalars) return fail(c, FLYBY_ERR_UNSUPPORTED, "render_packed: point scalars are float channels: use flyby_render");
  const int channels = (o->use_z && o->split_z) ? 4 : 3;
  if (out_depth && channels != 4) return fail(c, FLYBY_ERR_INVALID, "render_packed: a depth output needs use_z and split_z");
  const bool dev_out = is_device_ptr(out_rgb);
  if ((out_depth && is_device_ptr(out_depth) != dev_out) || (out_cell_ids && is_device_ptr(out_cell_ids) != dev_out))
    return fail(c, FLYBY_ERR_INVALID, "render_packed: outputs must be all host or all device memory");
  const size_t per_view = (size_t)res * res;
  const size_t n_pix = (size_t)(ve - vb) * per_view;
  if ((rc = views_setup(c, o, res))) return rc;
  const int nv = ve - vb;
  // Two chunks per call (more only to keep the float staging of a chunk under 256 MB): the copy-out of a call overlaps
  // the next call anyway, and every additional cudaMemcpyAsync costs more than its earlier start gains -- streamed
  // bench meshes: 3.15 / 3.04 / 3.14 / 3.48 / 3.80 ms per mesh with 1 / 2 / 4 / 8 / 16 chunks (the copy alone: 2.96)
  int n_chunks = (int)((n_pix * channels * sizeof(float) + (256u << 20) - 1) / (256u << 20));
  if (n_chunks < 2) n_chunks = 2;
  if (n_chunks > 16) n_chunks = 16;
  if (n_chunks > nv) n_chunks = nv;
  {
    static int env_chunks = -1;  // FLYBY_PK_CHUNKS: A/B hook
    if (env_chunks < 0) { const char* e = getenv("FLYBY_PK_CHUNKS"); env_chunks = e ? atoi(e) : 0; }
    if (env_chunks > 0) n_chunks = env_chunks > 16 ? 16 : (env_chunks > nv ? nv : env_chunks);
  }
  int bounds[17];
  size_t max_chunk = 0;
  for (int k = 0; k <= n_chunks; k++) bounds[k] = vb + (int)((long long)nv * k / n_chunks);
  for (int k = 0; k < n_chunks; k++) max_chunk = std::max(max_chunk, (size_t)(bounds[k + 1] - bounds[k]) * per_view);
  FB_CUDA(c, c->pk_feat.reserve(max_chunk * channels * sizeof(float)));
  float* df = c->pk_feat.as<float>();
  uint8_t* drgb = out_rgb;
  float* dz = out_depth;
  int32_t* di = out_cell_ids;
  int s = 0;
  if (!dev_out) {
    s = c->async_flip;
    c->async_flip ^= 1;
    for (int i = 0; i < 2; i++)
      if (!c->async_done[i]) FB_CUDA(c, cudaEventCreateWithFlags(&c->async_done[i], cudaEventDisableTiming));
    // growing a staging buffer frees the old one: drain the copies that may still read it first
    if (c->pk_rgb[s].cap < n_pix * 3 || (out_depth && c->pk_z[s].cap < n_pix * sizeof(float)) ||
        (out_cell_ids && c->async_ids[s].cap < n_pix * sizeof(int32_t)))
      FB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    FB_CUDA(c, c->pk_rgb[s].reserve(n_pix * 3));
    drgb = c->pk_rgb[s].as<uint8_t>();
    if (out_depth) { FB_CUDA(c, c->pk_z[s].reserve(n_pix * sizeof(float))); dz = c->pk_z[s].as<float>(); }
    if (out_cell_ids) { FB_CUDA(c, c->async_ids[s].reserve(n_pix * sizeof(int32_t))); di = c->async_ids[s].as<int32_t>(); }
    FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->async_done[s], 0));  // the set's previous copy-out (two calls ago)
  }
  for (int k = 0; k < n_chunks; k++) {
    const size_t off = (size_t)(bounds[k] - vb) * per_view;
    const size_t cnt = (size_t)(bounds[k + 1] - bounds[k]) * per_view;
    rc = trace_render(c, bounds[k], bounds[k + 1], res, o, df, di ? di + off : nullptr, nullptr, k == 0, k == n_chunks - 1);
    if (rc) return rc;
    // vector path: whole groups of 4 pixels and word-aligned colour / depth rows of this chunk
    const int vec = (cnt % 4 == 0 && off % 4 == 0 && ((uintptr_t)drgb % 4 == 0) && (!dz || (uintptr_t)dz % 16 == 0)) ? 1 : 0;
    if ((rc = fb::pack_rgb8(c, df, channels, (int64_t)cnt, drgb + 3 * off, dz ? dz + off : nullptr, vec))) return rc;
    if (!dev_out) {
      FB_CUDA(c, cudaEventRecord(c->chunk_ev[k], c->stream));
      FB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[k], 0));
      FB_CUDA(c, cudaMemcpyAsync(out_rgb + 3 * off, drgb + 3 * off, cnt * 3, cudaMemcpyDeviceToHost, c->copy_stream));
      if (dz) FB_CUDA(c, cudaMemcpyAsync(out_depth + off, dz + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->copy_stream));
      if (di) FB_CUDA(c, cudaMemcpyAsync(out_cell_ids + off, di + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  if (!dev_out) FB_CUDA(c, cudaEventRecord(c->async_done[s], c->copy_stream));
  return FLYBY_OK;
}

int flyby_render_perspective(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, double view_angle_deg,
                             double z_near, double z_far, float* out_features, int32_t* out_cell_ids, double* out_depth) {
  CHECK_CTX(c);
  int rc = check_render_args(c, vb, ve, res, o, true);
  if (rc) return rc;
  if (!(view_angle_deg > 0.0 && view_angle_deg < 180.0) || !(z_near >= 0.0) || !(z_far > z_near))
    return fail(c, FLYBY_ERR_INVALID, "render_perspective: need 0 < view angle < 180 and 0 <= z_near < z_far");
  const size_t n_pix = (size_t)(ve - vb) * res * res;
  const int channels = (o->use_z && o->split_z) ? 4 : 3;
  const double tan_half = tan(view_angle_deg * 0.5 * 3.14159265358979323846 / 180.0);  // host libm, as the oracle
  StageOut<float> sf;
  StageOut<int32_t> si;
  StageOut<double> sd;
  if ((rc = sf.begin(c, c->stage_feat, out_features, n_pix * channels))) return rc;
  if ((rc = si.begin(c, c->stage_ids, out_cell_ids, n_pix))) return rc;
  if ((rc = sd.begin(c, c->stage_t, out_depth, n_pix))) return rc;
  if ((rc = trace_render_perspective(c, vb, ve, res, o, tan_half, z_near, z_far, sf.dev, si.dev, sd.dev))) return rc;
  bool need = false;
  if ((rc = sf.finish(c, &need))) return rc;
  if ((rc = si.finish(c, &need))) return rc;
  if ((rc = sd.finish(c, &need))) return rc;
  return sync_if(c, need);
}

int flyby_cast(flyby_ctx* c, const double* rays, int64_t n, double tolerance, int32_t* out_ids, double* out_t,
               double* out_w) {
  CHECK_CTX(c);
  if (!c->mesh.built) return fail(c, FLYBY_ERR_INVALID, "cast: mesh not built");
  if (n < 0 || (n > 0 && !rays)) return fail(c, FLYBY_ERR_INVALID, "cast: bad arguments");
  const double* drays;
  int rc = stage_in(c, c->stage_in0, rays, (size_t)n * 6, &drays);
  if (rc) return rc;
  StageOut<int32_t> si;
  StageOut<double> st, sw;
  if ((rc = si.begin(c, c->stage_ids, out_ids, (size_t)n))) return rc;
  if ((rc = st.begin(c, c->stage_t, out_t, (size_t)n))) return rc;
  if ((rc = sw.begin(c, c->stage_out0, out_w, (size_t)n * 3))) return rc;
  rc = trace_cast(c, drays, n, tolerance, si.dev, st.dev, sw.dev);
  if (rc) return rc;
  bool need = !is_device_ptr(rays);
  if ((rc = si.finish(c, &need))) return rc;
  if ((rc = st.finish(c, &need))) return rc;
  if ((rc = sw.finish(c, &need))) return rc;
  return sync_if(c, need);
}

int flyby_point_features(flyby_ctx* c, const int32_t* cell_ids, int64_t n_pix, const float* feats, int F,
                         float zero, int mode, float* out) {
  CHECK_CTX(c);
  if (!c->mesh.has_mesh) return fail(c, FLYBY_ERR_INVALID, "point_features: no mesh");
  if (mode != 0 && mode != 1) return fail(c, FLYBY_ERR_UNSUPPORTED, "point_features: mode must be 0 (cell ids) or 1 (point ids)");
  if (!cell_ids || !feats || !out || F < 1 || n_pix < 0) return fail(c, FLYBY_ERR_INVALID, "point_features: bad arguments");
  const int32_t* dids;
  const float* dfeats;
  int rc;
  if (mode == 0 && (rc = ensure_tris_checked(c))) return rc;  // mode 1 never reads the triangles
  if ((rc = stage_in(c, c->stage_in0, cell_ids, (size_t)n_pix, &dids))) return rc;
  if ((rc = stage_in(c, c->stage_in1, feats, (size_t)c->mesh.V * F, &dfeats))) return rc;
  StageOut<float> so;
  if ((rc = so.begin(c, c->stage_out0, out, (size_t)n_pix * F))) return rc;
  if ((rc = label_point_features(c, dids, n_pix, dfeats, F, zero, mode, so.dev))) return rc;
  bool need = !is_device_ptr(cell_ids) || !is_device_ptr(feats);
  if ((rc = so.finish(c, &need))) return rc;
  return sync_if(c, need);
}

int flyby_point_ids(flyby_ctx* c, const int32_t* cell_ids, int64_t n_pix, int32_t* out) {
  CHECK_CTX(c);
  if (!c->mesh.has_mesh) return fail(c, FLYBY_ERR_INVALID, "point_ids: no mesh");
  if (!cell_ids || !out || n_pix < 0) return fail(c, FLYBY_ERR_INVALID, "point_ids: bad arguments");
  int rc;
  if ((rc = ensure_tris_checked(c))) return rc;
  const int32_t* dids;
  if ((rc = stage_in(c, c->stage_in0, cell_ids, (size_t)n_pix, &dids))) return rc;
  StageOut<int32_t> so;
  if ((rc = so.begin(c, c->stage_out0, out, (size_t)n_pix))) return rc;
  if ((rc = label_point_ids(c, dids, n_pix, so.dev))) return rc;
  bool need = !is_device_ptr(cell_ids);
  if ((rc = so.finish(c, &need))) return rc;
  return sync_if(c, need);
}

int flyby_set_point_scalars(flyby_ctx* c, const float* feats, int S, int mode, float zero) {
  CHECK_CTX(c);
  MeshDev& m = c->mesh;
  if (!m.has_mesh) return fail(c, FLYBY_ERR_INVALID, "set_point_scalars: call flyby_set_mesh first");
  if (S < 0 || S > 64 || (S > 0 && !feats)) return fail(c, FLYBY_ERR_INVALID, "set_point_scalars: 0 <= n_scalars <= 64 and a feature array");
  if (mode != 0 && mode != 1) return fail(c, FLYBY_ERR_UNSUPPORTED, "set_point_scalars: mode must be 0 (vertex 0) or 1 (barycentric)");
  m.n_scalars = 0;
  if (S == 0) return FLYBY_OK;
  FB_CUDA(c, m.scalars.reserve(sizeof(float) * (size_t)m.V * S));
  FB_CUDA(c, cudaMemcpyAsync(m.scalars.p, feats, sizeof(float) * (size_t)m.V * S, cudaMemcpyDefault, c->stream));
  m.n_scalars = S;
  m.scalar_mode = mode;
  m.scalar_zero = zero;
  return FLYBY_OK;
}

int flyby_get_winding(flyby_ctx* c, int64_t* n_flipped, uint8_t* flipped) {
  CHECK_CTX(c);
  MeshDev& m = c->mesh;
  if (!m.built) return fail(c, FLYBY_ERR_INVALID, "get_winding: mesh not built");
  if (n_flipped) *n_flipped = m.n_flipped;
  if (flipped && m.T > 0) {
    if (m.n_flipped > 0) FB_CUDA(c, cudaMemcpyAsync(flipped, m.flip.p, (size_t)m.T, cudaMemcpyDefault, c->stream));
    else if (is_device_ptr(flipped)) FB_CUDA(c, cudaMemsetAsync(flipped, 0, (size_t)m.T, c->stream));
    else memset(flipped, 0, (size_t)m.T);
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return FLYBY_OK;
}

int flyby_interpolate_features(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, const int32_t* cell_ids,
                               const float* feats, int F, float zero, float* out) {
  CHECK_CTX(c);
  int rc = check_render_args(c, vb, ve, res, o, true);
  if (rc) return rc;
  if (!cell_ids || !feats || !out || F < 1) return fail(c, FLYBY_ERR_INVALID, "interpolate_features: bad arguments");
  const size_t n_pix = (size_t)(ve - vb) * res * res;
  if ((rc = views_setup(c, o, res))) return rc;
  const int32_t* dids;
  const float* dfeats;
  if ((rc = stage_in(c, c->stage_in0, cell_ids, n_pix, &dids))) return rc;
  if ((rc = stage_in(c, c->stage_in1, feats, (size_t)c->mesh.V * F, &dfeats))) return rc;
  StageOut<float> so;
  if ((rc = so.begin(c, c->stage_out0, out, n_pix * F))) return rc;
  if ((rc = trace_interpolate(c, vb, ve, res, o, dids, dfeats, F, zero, so.dev))) return rc;
  bool need = !is_device_ptr(cell_ids) || !is_device_ptr(feats);
  if ((rc = so.finish(c, &need))) return rc;
  return sync_if(c, need);
}

int flyby_backproject(flyby_ctx* c, const int32_t* cell_ids, const int32_t* labels, int64_t n_pix, int K,
                      int32_t* votes) {
  CHECK_CTX(c);
  if (!c->mesh.has_mesh) return fail(c, FLYBY_ERR_INVALID, "backproject: no mesh");
  if (!cell_ids || !labels || !votes || K < 1 || n_pix < 0) return fail(c, FLYBY_ERR_INVALID, "backproject: bad arguments");
  const int32_t *dids, *dlab;
  int rc;
  if ((rc = stage_in(c, c->stage_in0, cell_ids, (size_t)n_pix, &dids))) return rc;
  if ((rc = stage_in(c, c->stage_in1, labels, (size_t)n_pix, &dlab))) return rc;
  const size_t nv = (size_t)c->mesh.T * K;
  int32_t* dvotes = votes;
  bool host_votes = !is_device_ptr(votes);
  if (host_votes) {
    FB_CUDA(c, c->stage_out0.reserve(nv * 4 + 4));
    dvotes = c->stage_out0.as<int32_t>();
    FB_CUDA(c, cudaMemcpyAsync(dvotes, votes, nv * 4, cudaMemcpyHostToDevice, c->stream));
  }
  if ((rc = label_backproject(c, dids, dlab, n_pix, K, dvotes))) return rc;
  bool need = !is_device_ptr(cell_ids) || !is_device_ptr(labels);
  if (host_votes) {
    FB_CUDA(c, cudaMemcpyAsync(votes, dvotes, nv * 4, cudaMemcpyDeviceToHost, c->stream));
    need = true;
  }
  return sync_if(c, need);
}

int flyby_backproject_points(flyby_ctx* c, const int32_t* cell_ids, const int32_t* labels, int64_t n_pix, int K,
                             int32_t* votes) {
  CHECK_CTX(c);
  if (!c->mesh.has_mesh) return fail(c, FLYBY_ERR_INVALID, "backproject_points: no mesh");
  if (!cell_ids || !labels || !votes || K < 1 || n_pix < 0) return fail(c, FLYBY_ERR_INVALID, "backproject_points: bad arguments");
  const int32_t *dids, *dlab;
  int rc;
  if ((rc = ensure_tris_checked(c))) return rc;
  if ((rc = stage_in(c, c->stage_in0, cell_ids, (size_t)n_pix, &dids))) return rc;
  if ((rc = stage_in(c, c->stage_in1, labels, (size_t)n_pix, &dlab))) return rc;
  const size_t nv = (size_t)c->mesh.V * K;
  int32_t* dvotes = votes;
  bool host_votes = !is_device_ptr(votes);
  if (host_votes) {
    FB_CUDA(c, c->stage_out0.reserve(nv * 4 + 4));
    dvotes = c->stage_out0.as<int32_t>();
    FB_CUDA(c, cudaMemcpyAsync(dvotes, votes, nv * 4, cudaMemcpyHostToDevice, c->stream));
  }
  if ((rc = label_backproject_points(c, dids, dlab, n_pix, K, dvotes))) return rc;
  bool need = !is_device_ptr(cell_ids) || !is_device_ptr(labels);
  if (host_votes) {
    FB_CUDA(c, cudaMemcpyAsync(votes, dvotes, nv * 4, cudaMemcpyDeviceToHost, c->stream));
    need = true;
  }
  return sync_if(c, need);
}

int flyby_backproject_soft(flyby_ctx* c, const int32_t* cell_ids, const float* probs, int64_t n_pix, int K,
                           int64_t* votes_fx) {
  CHECK_CTX(c);
  if (!c->mesh.has_mesh) return fail(c, FLYBY_ERR_INVALID, "backproject_soft: no mesh");
  if (!cell_ids || !probs || !votes_fx || K < 1 || n_pix < 0) return fail(c, FLYBY_ERR_INVALID, "backproject_soft: bad arguments");
  const int32_t* dids;
  const float* dprobs;
  int rc;
  if ((rc = stage_in(c, c->stage_in0, cell_ids, (size_t)n_pix, &dids))) return rc;
  if ((rc = stage_in(c, c->stage_in1, probs, (size_t)n_pix * K, &dprobs))) return rc;
  const size_t nv = (size_t)c->mesh.T * K;
  int64_t* dvotes = votes_fx;
  bool host_votes = !is_device_ptr(votes_fx);
  if (host_votes) {
    FB_CUDA(c, c->stage_out0.reserve(nv * 8 + 8));
    dvotes = c->stage_out0.as<int64_t>();
    FB_CUDA(c, cudaMemcpyAsync(dvotes, votes_fx, nv * 8, cudaMemcpyHostToDevice, c->stream));
  }
  if ((rc = label_backproject_soft(c, dids, dprobs, n_pix, K, dvotes))) return rc;
  bool need = !is_device_ptr(cell_ids) || !is_device_ptr(probs);
  if (host_votes) {
    FB_CUDA(c, cudaMemcpyAsync(votes_fx, dvotes, nv * 8, cudaMemcpyDeviceToHost, c->stream));
    need = true;
  }
  return sync_if(c, need);
}

int flyby_votes_argmax(flyby_ctx* c, const int32_t* votes, int64_t n_cells, int K, int32_t dflt, int32_t* out) {
  CHECK_CTX(c);
  if (!votes || !out || K < 1 || n_cells < 0) return fail(c, FLYBY_ERR_INVALID, "votes_argmax: bad arguments");
  const int32_t* dv;
  int rc;
  if ((rc = stage_in(c, c->stage_in0, votes, (size_t)n_cells * K, &dv))) return rc;
  StageOut<int32_t> so;
  if ((rc = so.begin(c, c->stage_out0, out, (size_t)n_cells))) return rc;
  if ((rc = label_argmax(c, dv, n_cells, K, dflt, so.dev))) return rc;
  bool need = !is_device_ptr(votes);
  if ((rc = so.finish(c, &need))) return rc;
  return sync_if(c, need);
}

int flyby_cells_to_points(flyby_ctx* c, const int32_t* cell_labels, int32_t dflt, int32_t* out) {
  CHECK_CTX(c);
  if (!c->mesh.has_mesh) return fail(c, FLYBY_ERR_INVALID, "cells_to_points: no mesh");
  if (!cell_labels || !out) return fail(c, FLYBY_ERR_INVALID, "cells_to_points: bad arguments");
  const int32_t* dl;
  int rc;
  if ((rc = ensure_tris_checked(c))) return rc;
  if ((rc = stage_in(c, c->stage_in0, cell_labels, (size_t)c->mesh.T, &dl))) return rc;
  StageOut<int32_t> so;
  if ((rc = so.begin(c, c->stage_out0, out, (size_t)c->mesh.V))) return rc;
  if ((rc = label_cells_to_points(c, dl, dflt, so.dev))) return rc;
  bool need = !is_device_ptr(cell_labels);
  if ((rc = so.finish(c, &need))) return rc;
  return sync_if(c, need);
}

// ---- mesh-label post-processing (postproc.cu): labels int32 [V], host or device, edited in place
namespace {
struct LabelsIO {
  int32_t* user = nullptr;
  int32_t* dev = nullptr;
  bool host = false;
};
int labels_begin(flyby_ctx* c, int32_t* labels, const char* what, LabelsIO* io) {
  if (!c->mesh.has_mesh || !c->mesh.built) return fail(c, FLYBY_ERR_INVALID, "%s: call flyby_set_mesh and flyby_build first", what);
  if (!labels && c->mesh.V > 0) return fail(c, FLYBY_ERR_INVALID, "%s: labels is NULL", what);
  { int rc = wait_tree(c); if (rc) return rc; }  // the adjacency CSR is finished on the side stream
  io->user = io->dev = labels;
  io->host = labels && !is_device_ptr(labels);
  if (io->host) {
    FB_CUDA(c, c->stage_in0.reserve((size_t)c->mesh.V * sizeof(int32_t) + 16));
    io->dev = c->stage_in0.as<int32_t>();
    FB_CUDA(c, cudaMemcpyAsync(io->dev, labels, (size_t)c->mesh.V * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  }
  return FLYBY_OK;
}
int labels_end(flyby_ctx* c, const LabelsIO& io, int rc) {
  if (rc) return rc;
  if (io.host) {
    FB_CUDA(c, cudaMemcpyAsync(io.user, io.dev, (size_t)c->mesh.V * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return FLYBY_OK;
}
}  // namespace

int flyby_labels_remove_islands(flyby_ctx* c, int32_t* labels, int32_t label, int64_t min_count, int ignore_neg1) {
  CHECK_CTX(c);
  LabelsIO io;
  int rc = labels_begin(c, labels, "labels_remove_islands", &io);
  if (rc) return rc;
  return labels_end(c, io, pp_remove_islands(c, io.dev, label, min_count, ignore_neg1));
}
int flyby_labels_connectivity(flyby_ctx* c, int32_t* labels, int32_t label, int32_t start_label, int64_t* n_regions) {
  CHECK_CTX(c);
  LabelsIO io;
  int rc = labels_begin(c, labels, "labels_connectivity", &io);
  if (rc) return rc;
  return labels_end(c, io, pp_connectivity(c, io.dev, label, start_label, n_regions));
}
int flyby_labels_erode(flyby_ctx* c, int32_t* labels, int32_t label, int has_ignore, int32_t ignore_label,
                       int64_t iterations, int has_target, int32_t target) {
  CHECK_CTX(c);
  LabelsIO io;
  int rc = labels_begin(c, labels, "labels_erode", &io);
  if (rc) return rc;
  return labels_end(c, io, pp_erode(c, io.dev, label, has_ignore, ignore_label, iterations, has_target, target));
}
int flyby_labels_dilate(flyby_ctx* c, int32_t* labels, int32_t label, int64_t iterations, int over_target, int32_t target) {
  CHECK_CTX(c);
  LabelsIO io;
  int rc = labels_begin(c, labels, "labels_dilate", &io);
  if (rc) return rc;
  return labels_end(c, io, pp_dilate(c, io.dev, label, iterations, over_target, target));
}
int flyby_labels_relabel(flyby_ctx* c, int32_t* labels, int32_t label, int32_t relabel) {
  CHECK_CTX(c);
  LabelsIO io;
  int rc = labels_begin(c, labels, "labels_relabel", &io);
  if (rc) return rc;
  return labels_end(c, io, pp_relabel(c, io.dev, label, relabel));
}

int flyby_nccl_unique_id(flyby_ctx* c, uint8_t id_out[128]) {
  CHECK_CTX(c);
  return nccl_unique_id(c, id_out);
}
int flyby_nccl_init(flyby_ctx* c, const uint8_t id[128], int rank, int world) {
  CHECK_CTX(c);
  return nccl_init(c, id, rank, world);
}
int flyby_votes_allreduce(flyby_ctx* c, int32_t* votes, int64_t n, int root) {
  CHECK_CTX(c);
  if (!is_device_ptr(votes)) return fail(c, FLYBY_ERR_INVALID, "votes_allreduce: votes must be device memory");
  return nccl_votes_allreduce(c, votes, n, root);
}

// ---- CUDA-graph replay of a step (streams of equally shaped meshes)
struct flyby_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  bool has_build = false;  // the step contains a deferred build: every replay leaves a verdict to be read
};

int flyby_capture_begin(flyby_ctx* c) {
  CHECK_CTX(c);
  if (c->capturing) return fail(c, FLYBY_ERR_INVALID, "capture_begin: already capturing");
  // what the step may lean on must not come from outside the capture: join the side stream's pending work now
  { int rc = wait_tree(c); if (rc) return rc; }
  c->tree_pending = false;
  if (c->ev_front && c->ev_front_live) FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_front, 0));
  // the check flags of deferred builds are sticky inside a graph (a replay must not clear what an earlier replay
  // found): start from clean flags unless an unread verdict is pending
  if (!c->deferred_pending)
    FB_CUDA(c, cudaMemsetAsync(c->counters.as<unsigned long long>() + 7, 0, 8, c->stream));
  FB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
  c->capturing = true;
  c->cap_forked = c->cap_front_forked = c->cap_has_build = false;
  c->cap_launches0 = c->launches;
  capture_flag() = true;
  int rc = raster_capture_prologue(c);
  if (rc) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(c->stream, &g);
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    c->capturing = false;
    capture_flag() = false;
  }
  return rc;
}

int flyby_capture_end(flyby_ctx* c, flyby_graph** out) {
  CHECK_CTX(c);
  if (!out) return FLYBY_ERR_INVALID;
  *out = nullptr;
  if (!c->capturing) return fail(c, FLYBY_ERR_INVALID, "capture_end: not capturing");
  cudaError_t e = cudaSuccess;
  if (c->side_stream && c->cap_forked) {  // join whatever the side stream still does for this step
    e = cudaEventRecord(c->ev_tree, c->side_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_tree, 0);
  }
  cudaGraph_t g = nullptr;
  cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
  c->capturing = false;
  capture_flag() = false;
  c->tree_pending = false;   // replays are ordered on the context stream as a whole
  c->front_clean = 0;        // a replayed render leaves the front buffer used
  c->ev_front_live = false;  // (its last record belongs to the graph)
  if (e == cudaSuccess) e = e2;
  if (e != cudaSuccess || !g) {
    if (g) cudaGraphDestroy(g);
    cudaGetLastError();
    return fail(c, FLYBY_ERR_CUDA, "capture_end: %s (a call inside the capture allocated, synchronised or used host buffers?)",
                cudaGetErrorString(e));
  }
  flyby_graph* fg = new (std::nothrow) flyby_graph();
  if (!fg) { cudaGraphDestroy(g); return FLYBY_ERR_INVALID; }
  fg->graph = g;
  fg->launches = c->launches - c->cap_launches0;
  fg->has_build = c->cap_has_build;
  c->launches = c->cap_launches0;  // captured, not launched
  e = cudaGraphInstantiate(&fg->exec, g, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g);
    delete fg;
    return fail(c, FLYBY_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  }
  *out = fg;
  return FLYBY_OK;
}

int flyby_graph_launch(flyby_ctx* c, flyby_graph* g) {
  CHECK_CTX(c);
  if (!g || !g->exec) return fail(c, FLYBY_ERR_INVALID, "graph_launch: no graph");
  if (c->capturing) return fail(c, FLYBY_ERR_INVALID, "graph_launch: the context is capturing");
  { int rc = wait_tree(c); if (rc) return rc; }
  c->tree_pending = false;
  if (c->ev_front && c->ev_front_live) FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_front, 0));  // side-stream init of the last render
  c->front_clean = 0;
  FB_CUDA(c, cudaGraphLaunch(g->exec, c->stream));
  c->launches += g->launches;
  if (g->has_build) c->deferred_pending = true;
  return FLYBY_OK;
}

void flyby_graph_destroy(flyby_graph* g) {
  if (!g) return;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
}

int flyby_get_stats(flyby_ctx* c, flyby_stats* out) {
  CHECK_CTX(c);
  if (!out) return FLYBY_ERR_INVALID;
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->side_stream) FB_CUDA(c, cudaStreamSynchronize(c->side_stream));
  flyby_stats s = c->stats;
  auto el = [&](int a, int b) -> float {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); return 0.f; }
    return ms;
  };
  if (c->mesh.built) {
    s.ms_normalise = el(0, 1);
    s.ms_normals = el(1, 2);
    // with the side stream the normals (events 1 -> 2) run beside the sort, which then starts at event 1
    if (c->mesh.T > 0) { s.ms_sort = c->side_stream ? el(1, 3) : el(2, 3); s.ms_tree = el(3, 4); }
  }
  s.ms_render = el(5, 6);
  s.ms_trace = s.ms_resolve = 0.f;  // sums over the pixel slabs / view chunks of the last render
  for (int k = 0; k < c->n_slab_ev && k < flyby_ctx::kMaxSlabEv; k++) {
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, c->slab_ev[3 * k], c->slab_ev[3 * k + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&b, c->slab_ev[3 * k + 1], c->slab_ev[3 * k + 2]) == cudaSuccess) {
      s.ms_trace += a;
      s.ms_resolve += b;
    } else {
      cudaGetLastError();
    }
  }
  unsigned long long h[16] = {0};
  FB_CUDA(c, cudaMemcpy(h, c->counters.p, sizeof h, cudaMemcpyDeviceToHost));
  if (getenv("FLYBY_DEBUG_STATS"))
    fprintf(stderr, "[flyby] pixel lists: first %u (bvh mode: 1 candidate; raster mode: several), second %u (bvh: several; raster: overflow), re-traced exactly %llu; raster spill: %u pairs, %llu pixels\n",
            (unsigned)(h[14] & 0xffffffffu), (unsigned)(h[14] >> 32), h[5], (unsigned)(h[13] & 0xffffffffu), h[4]);
  if (getenv("FLYBY_DEBUG_STATS") && h[9])
    fprintf(stderr, "[flyby] packets %llu, lanes/packet %.2f, node visits/packet %.1f, leaf visits/packet %.1f\n", h[9],
            (double)h[10] / h[9], (double)h[8] / h[9], (double)h[11] / h[9]);
  s.nodes_visited = (int64_t)h[1];
  s.tris_tested = (int64_t)h[2];
  s.n_hits = (int64_t)h[3];
  *out = s;
  return FLYBY_OK;
}

int flyby_set_count_work(flyby_ctx* c, int enable) {
  if (!c) return FLYBY_ERR_INVALID;
  c->count_work = enable ? 1 : 0;
  return FLYBY_OK;
}

int flyby_launch_count(flyby_ctx* c, int64_t* n) {
  if (!c || !n) return FLYBY_ERR_INVALID;
  *n = c->launches;
  return FLYBY_OK;
}

}  // extern "C"
