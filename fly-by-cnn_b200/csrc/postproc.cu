// postproc.cu -- mesh-label post-processing, the step that follows the vote back-projection in every
// predictor of the reference (predict_v3.py:93-133, dental_model_seg.py:228-239, predict.py:190-201).
//
// Reference replaced: src/py/post_process.py
//   ConnectedRegion :243-255 / GetNeighborIds utils.py:233-247   -> connected components by min-id hooking
//   NeighborLabel :257-281, RemoveIslands :285-295               -> remove_islands
//   ConnectivityLabeling :297-304                                -> connectivity
//   ErodeLabel :306-346, DilateLabel :348-375, ReLabel :377-380  -> erode / dilate / relabel
// The reference runs Python loops over vtkPolyData links, one point at a time.  Its updates are either
// synchronous by construction (erode / dilate collect first, write afterwards) or touch disjoint regions
// (islands and regions of one label never neighbour each other), so each of them has a data-parallel
// form with the same result: one thread per point over the CSR of incident cells that flyby_build keeps
// for the point normals.  All integer work: results are bit-identical to the oracle.
//
// "Neighbours of p" = the other points of the cells incident to p (GetPointCells + GetCellPoints); the
// adjacency is read as (adj_start, adj, tris) without materialising neighbour lists, so a neighbour shows
// up once per shared cell -- harmless where the operation is idempotent, de-duplicated where it counts.
#include "ctx.h"

namespace fb {

struct Links {
  const uint32_t* start;  // [V+1]
  const int32_t* cells;   // [3T] incident cells, ascending
  const int32_t* tris;    // [T,3]
};

// ---------------------------------------------------------------- connected components of one label
__global__ void cc_init_kernel(const int32_t* __restrict__ labels, int32_t label, int64_t V, int32_t* comp,
                               int32_t* size) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  comp[v] = labels[v] == label ? (int32_t)v : -1;
  size[v] = 0;
}

__device__ __forceinline__ int32_t cc_find(const int32_t* comp, int32_t v) {
  // pointers only ever move to smaller ids, so a stale read still lands on an ancestor
  for (;;) {
    const int32_t p = __ldcg(comp + v);
    if (p == v) return v;
    v = p;
  }
}

// hook the larger root of every edge under the smaller one; repeated until no edge joins two roots
__global__ void cc_hook_kernel(Links L, int64_t V, int32_t* comp, int* changed) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V || __ldcg(comp + v) < 0) return;
  int32_t rv = cc_find(comp, (int32_t)v);
  bool any = false;
  for (uint32_t q = L.start[v]; q < L.start[v + 1]; q++) {
    const int32_t* cell = L.tris + 3 * (int64_t)L.cells[q];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int32_t u = cell[k];
      if (u == v || __ldcg(comp + u) < 0) continue;
      const int32_t ru = cc_find(comp, u);
      if (ru == rv) continue;
      const int32_t hi = max(ru, rv), lo = min(ru, rv);
      atomicMin(comp + hi, lo);
      rv = lo;
      any = true;
    }
  }
  if (any) *changed = 1;
}

__global__ void cc_compress_kernel(int64_t V, int32_t* comp) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V || comp[v] < 0) return;
  comp[v] = cc_find(comp, (int32_t)v);
}

__global__ void cc_size_kernel(int64_t V, const int32_t* __restrict__ comp, int32_t* size) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V || comp[v] < 0) return;
  atomicAdd(size + comp[v], 1);
}

// comp[v] = smallest point id of v's region (or -1), size[root] = points of the region
static int components(flyby_ctx* c, const int32_t* labels, int32_t label, int32_t* comp, int32_t* size, int* changed_dev) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t V = m.V;
  const unsigned g = (unsigned)cdiv(V, 256);
  const Links L{m.adj_start.as<uint32_t>(), m.adj.as<int32_t>(), m.tris.as<int32_t>()};
  cc_init_kernel<<<g, 256, 0, st>>>(labels, label, V, comp, size);
  FB_LAUNCH_CHECK(c);
  // a losing atomicMin drops a link until the next round finds it again through its edge, so the number of rounds
  // is not bounded by log V (long thin regions need more): loop until a round joins nothing
  for (int64_t round = 0; round <= V; round++) {
    FB_CUDA(c, cudaMemsetAsync(changed_dev, 0, sizeof(int), st));
    cc_hook_kernel<<<g, 256, 0, st>>>(L, V, comp, changed_dev);
    FB_LAUNCH_CHECK(c);
    cc_compress_kernel<<<g, 256, 0, st>>>(V, comp);
    FB_LAUNCH_CHECK(c);
    int changed = 0;
    FB_CUDA(c, cudaMemcpyAsync(&changed, changed_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaStreamSynchronize(st));
    if (!changed) {
      cc_size_kernel<<<g, 256, 0, st>>>(V, comp, size);
      FB_LAUNCH_CHECK(c);
      return FLYBY_OK;
    }
  }
  return fail(c, FLYBY_ERR_CUDA, "connected components did not converge");
}

// ---------------------------------------------------------------- remove islands
__global__ void minmax_kernel(const int32_t* __restrict__ labels, int64_t V, int32_t* mm) {
  int32_t lo = INT32_MAX, hi = INT32_MIN;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
    lo = min(lo, labels[v]);
    hi = max(hi, labels[v]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm, lo);
    atomicMax(mm + 1, hi);
  }
}

// flag[r] = 1 for roots of regions smaller than min_count (scanned into a dense island index afterwards)
__global__ void island_flag_kernel(int64_t V, const int32_t* __restrict__ comp, const int32_t* __restrict__ size,
                                   int64_t min_count, uint32_t* flag) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  flag[v] = (comp[v] == v && (int64_t)size[v] < min_count) ? 1u : 0u;
  if (v == 0) flag[V] = 0u;
}

// NeighborLabel: every point n whose label differs votes once for each island it touches
__global__ void island_vote_kernel(Links L, int64_t V, const int32_t* __restrict__ labels, int32_t label,
                                   const int32_t* __restrict__ comp, const int32_t* __restrict__ size, int64_t min_count,
                                   const uint32_t* __restrict__ island, int32_t lmin, int K, int32_t* hist, int32_t* first) {
  const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= V || labels[n] == label) return;
  const uint32_t b = L.start[n], e = L.start[n + 1];
  for (uint32_t q = b; q < e; q++) {
    const int32_t* cell = L.tris + 3 * (int64_t)L.cells[q];
    for (int k = 0; k < 3; k++) {
      const int32_t u = cell[k];
      if (u == n) continue;
      const int32_t r = comp[u];
      if (r < 0 || (int64_t)size[r] >= min_count) continue;
      // the unique() of the reference: count this island only at its first mention around n
      bool seen = false;
      for (uint32_t q2 = b; q2 <= q && !seen; q2++) {
        const int32_t* cell2 = L.tris + 3 * (int64_t)L.cells[q2];
        const int kend = q2 == q ? k : 3;
        for (int k2 = 0; k2 < kend; k2++) {
          const int32_t u2 = cell2[k2];
          if (u2 != n && comp[u2] == r) { seen = true; break; }
        }
      }
      if (seen) continue;
      const int64_t slot = (int64_t)island[r] * K + (labels[n] - lmin);
      atomicAdd(hist + slot, 1);
      atomicMin(first + slot, (int32_t)n);
    }
  }
}

// max(neighbor_labels, key=neighbor_labels.count): the most frequent label, ties to the label whose first
// (lowest point id) occurrence comes first; -1 when the island has no differing neighbour
__global__ void island_decide_kernel(int64_t n_islands, int32_t lmin, int K, const int32_t* __restrict__ hist,
                                     const int32_t* __restrict__ first, int32_t* choice) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_islands) return;
  int32_t best = -1, best_count = 0, best_first = INT32_MAX;
  for (int k = 0; k < K; k++) {
    const int32_t cnt = hist[i * K + k], f = first[i * K + k];
    if (cnt > best_count || (cnt == best_count && cnt > 0 && f < best_first)) {
      best_count = cnt; best_first = f; best = lmin + k;
    }
  }
  choice[i] = best_count > 0 ? best : -1;
}

__global__ void island_apply_kernel(int64_t V, const int32_t* __restrict__ comp, const int32_t* __restrict__ size,
                                    int64_t min_count, const uint32_t* __restrict__ island,
                                    const int32_t* __restrict__ choice, int32_t* labels) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int32_t r = comp[v];
  if (r < 0 || (int64_t)size[r] >= min_count) return;
  const int32_t nl = choice[island[r]];
  if (nl != -1) labels[v] = nl;  // post_process.py:292 (reached only with ignore_neg1)
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct PPScratch {
  int32_t *comp, *size;
  uint32_t* flag;
  int* small;  // [8] changed flag, min, max
};
static int pp_scratch(flyby_ctx* c, PPScratch* s) {
  const size_t V = (size_t)c->mesh.V;
  FB_CUDA(c, c->pp_tmp.reserve((3 * V + 64) * sizeof(int32_t)));
  s->comp = c->pp_tmp.as<int32_t>();
  s->size = s->comp + V;
  s->flag = reinterpret_cast<uint32_t*>(s->size + V);
  s->small = reinterpret_cast<int*>(s->flag + V + 8);
  return FLYBY_OK;
}

int pp_remove_islands(flyby_ctx* c, int32_t* labels, int32_t label, int64_t min_count, int ignore_neg1) {
  // post_process.py:292: the relabelling sits behind `ignore_neg1 == True and ...`; without the flag the
  // reference finds the islands and leaves them alone
  if (!ignore_neg1 || c->mesh.V == 0) return FLYBY_OK;
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t V = m.V;
  const unsigned g = (unsigned)cdiv(V, 256);
  PPScratch s;
  int rc = pp_scratch(c, &s);
  if (rc) return rc;
  if ((rc = components(c, labels, label, s.comp, s.size, s.small))) return rc;
  const int32_t init_mm[2] = {INT32_MAX, INT32_MIN};
  FB_CUDA(c, cudaMemcpyAsync(s.small + 1, init_mm, sizeof init_mm, cudaMemcpyHostToDevice, st));
  minmax_kernel<<<(unsigned)(c->sm_count * 4), 256, 0, st>>>(labels, V, s.small + 1);
  FB_LAUNCH_CHECK(c);
  island_flag_kernel<<<(unsigned)cdiv(V + 1, 256), 256, 0, st>>>(V, s.comp, s.size, min_count, s.flag);
  FB_LAUNCH_CHECK(c);
  if ((rc = exclusive_scan_u32(c, s.flag, V + 1, m.scan_tmp))) return rc;
  int32_t mm[2];
  uint32_t n_islands = 0;
  FB_CUDA(c, cudaMemcpyAsync(mm, s.small + 1, sizeof mm, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(&n_islands, s.flag + V, sizeof n_islands, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  if (n_islands == 0) return FLYBY_OK;
  const int64_t K = (int64_t)mm[1] - mm[0] + 1;
  if (K > 65536) return fail(c, FLYBY_ERR_UNSUPPORTED, "remove_islands: label range %lld too wide", (long long)K);
  const size_t cells = (size_t)n_islands * (size_t)K;
  FB_CUDA(c, c->pp_tmp2.reserve((2 * cells + n_islands + 16) * sizeof(int32_t)));
  int32_t* hist = c->pp_tmp2.as<int32_t>();
  int32_t* first = hist + cells;
  int32_t* choice = first + cells;
  FB_CUDA(c, cudaMemsetAsync(hist, 0, cells * sizeof(int32_t), st));
  fill_i32_kernel<<<(unsigned)cdiv((int64_t)cells, 256), 256, 0, st>>>(first, (int64_t)cells, INT32_MAX);
  FB_LAUNCH_CHECK(c);
  const Links L{m.adj_start.as<uint32_t>(), m.adj.as<int32_t>(), m.tris.as<int32_t>()};
  island_vote_kernel<<<g, 256, 0, st>>>(L, V, labels, label, s.comp, s.size, min_count, s.flag, mm[0], (int)K, hist, first);
  FB_LAUNCH_CHECK(c);
  island_decide_kernel<<<(unsigned)cdiv(n_islands, 256), 256, 0, st>>>(n_islands, mm[0], (int)K, hist, first, choice);
  FB_LAUNCH_CHECK(c);
  island_apply_kernel<<<g, 256, 0, st>>>(V, s.comp, s.size, min_count, s.flag, choice, labels);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// ---------------------------------------------------------------- connectivity labelling
__global__ void root_flag_kernel(int64_t V, const int32_t* __restrict__ comp, uint32_t* flag) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  flag[v] = comp[v] == v ? 1u : 0u;
  if (v == 0) flag[V] = 0u;
}
// regions in the order of their lowest point id get start_label, start_label + 1, ...
__global__ void connectivity_apply_kernel(int64_t V, const int32_t* __restrict__ comp, const uint32_t* __restrict__ rank,
                                          int32_t start_label, int32_t* labels) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V || comp[v] < 0) return;
  labels[v] = start_label + (int32_t)rank[comp[v]];
}

int pp_connectivity(flyby_ctx* c, int32_t* labels, int32_t label, int32_t start_label, int64_t* n_regions) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t V = m.V;
  if (n_regions) *n_regions = 0;
  if (V == 0) return FLYBY_OK;
  const unsigned g = (unsigned)cdiv(V, 256);
  PPScratch s;
  int rc = pp_scratch(c, &s);
  if (rc) return rc;
  if ((rc = components(c, labels, label, s.comp, s.size, s.small))) return rc;
  root_flag_kernel<<<(unsigned)cdiv(V + 1, 256), 256, 0, st>>>(V, s.comp, s.flag);
  FB_LAUNCH_CHECK(c);
  if ((rc = exclusive_scan_u32(c, s.flag, V + 1, m.scan_tmp))) return rc;
  connectivity_apply_kernel<<<g, 256, 0, st>>>(V, s.comp, s.flag, start_label, labels);
  FB_LAUNCH_CHECK(c);
  if (n_regions) {
    uint32_t n = 0;
    FB_CUDA(c, cudaMemcpyAsync(&n, s.flag + V, sizeof n, cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaStreamSynchronize(st));
    *n_regions = n;
  }
  return FLYBY_OK;
}

// ---------------------------------------------------------------- erode / dilate / relabel
// One synchronous sweep: points of `label` take the label of their lowest-id neighbour that qualifies
// (ErodeLabel breaks at the first hit of the sorted unique neighbour list).
__global__ void erode_kernel(Links L, int64_t V, const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t label,
                             int has_ignore, int32_t ignore_label, int has_target, int32_t target, int* changed) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  int32_t mine = in[v];
  if (mine == label) {
    int32_t best_u = INT32_MAX, best_l = label;
    for (uint32_t q = L.start[v]; q < L.start[v + 1]; q++) {
      const int32_t* cell = L.tris + 3 * (int64_t)L.cells[q];
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int32_t u = cell[k];
        if (u == v || u >= best_u) continue;
        const int32_t nl = in[u];
        if (nl != label && (!has_ignore || nl != ignore_label) && (!has_target || nl == target)) { best_u = u; best_l = nl; }
      }
    }
    if (best_u != INT32_MAX) { mine = best_l; *changed = 1; }
  }
  out[v] = mine;
}

// One synchronous sweep: a point joins `label` if one of its neighbours has it and the point itself is
// eligible (any other label, or exactly `target` with dilateOverTarget).
__global__ void dilate_kernel(Links L, int64_t V, const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t label,
                              int over_target, int32_t target) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= V) return;
  int32_t mine = in[v];
  if (over_target ? mine == target : mine != label) {
    bool touch = false;
    for (uint32_t q = L.start[v]; q < L.start[v + 1] && !touch; q++) {
      const int32_t* cell = L.tris + 3 * (int64_t)L.cells[q];
#pragma unroll
      for (int k = 0; k < 3; k++) touch = touch || (cell[k] != v && in[cell[k]] == label);
    }
    if (touch) mine = label;
  }
  out[v] = mine;
}

__global__ void relabel_kernel(int32_t* labels, int64_t V, int32_t label, int32_t relabel) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v < V && labels[v] == label) labels[v] = relabel;
}

#define PP_BATCH 8  // erode sweeps between two looks at the "changed" flags

int pp_erode(flyby_ctx* c, int32_t* labels, int32_t label, int has_ignore, int32_t ignore_label, int64_t iterations,
             int has_target, int32_t target) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t V = m.V;
  if (V == 0 || iterations == 0) return FLYBY_OK;
  const unsigned g = (unsigned)cdiv(V, 256);
  FB_CUDA(c, c->pp_tmp.reserve(((size_t)V + 64) * sizeof(int32_t)));
  int32_t* other = c->pp_tmp.as<int32_t>();
  int* flags = reinterpret_cast<int*>(other + V);
  const Links L{m.adj_start.as<uint32_t>(), m.adj.as<int32_t>(), m.tris.as<int32_t>()};
  int32_t *cur = labels, *nxt = other;
  for (;;) {
    const int batch = (int)((iterations < 0 || iterations > PP_BATCH) ? PP_BATCH : iterations);
    FB_CUDA(c, cudaMemsetAsync(flags, 0, PP_BATCH * sizeof(int), st));
    for (int i = 0; i < batch; i++) {
      erode_kernel<<<g, 256, 0, st>>>(L, V, cur, nxt, label, has_ignore, ignore_label, has_target, target, flags + i);
      FB_LAUNCH_CHECK(c);
      int32_t* t = cur; cur = nxt; nxt = t;
    }
    int h[PP_BATCH];
    FB_CUDA(c, cudaMemcpyAsync(h, flags, sizeof h, cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaStreamSynchronize(st));
    bool stalled = false;  // a sweep without a change ends the loop (post_process.py:343); later sweeps were no-ops
    for (int i = 0; i < batch; i++) stalled = stalled || !h[i];
    if (iterations > 0) iterations -= batch;
    if (stalled || iterations == 0) break;
  }
  if (cur != labels) FB_CUDA(c, cudaMemcpyAsync(labels, cur, (size_t)V * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return FLYBY_OK;
}

int pp_dilate(flyby_ctx* c, int32_t* labels, int32_t label, int64_t iterations, int over_target, int32_t target) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t V = m.V;
  if (V == 0 || iterations <= 0) return FLYBY_OK;
  const unsigned g = (unsigned)cdiv(V, 256);
  FB_CUDA(c, c->pp_tmp.reserve(((size_t)V + 64) * sizeof(int32_t)));
  int32_t* other = c->pp_tmp.as<int32_t>();
  const Links L{m.adj_start.as<uint32_t>(), m.adj.as<int32_t>(), m.tris.as<int32_t>()};
  int32_t *cur = labels, *nxt = other;
  for (int64_t i = 0; i < iterations; i++) {
    dilate_kernel<<<g, 256, 0, st>>>(L, V, cur, nxt, label, over_target, target);
    FB_LAUNCH_CHECK(c);
    int32_t* t = cur; cur = nxt; nxt = t;
  }
  if (cur != labels) FB_CUDA(c, cudaMemcpyAsync(labels, cur, (size_t)V * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  return FLYBY_OK;
}

int pp_relabel(flyby_ctx* c, int32_t* labels, int32_t label, int32_t relabel) {
  const int64_t V = c->mesh.V;
  if (V == 0) return FLYBY_OK;
  relabel_kernel<<<(unsigned)cdiv(V, 256), 256, 0, c->stream>>>(labels, V, label, relabel);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

}  // namespace fb
