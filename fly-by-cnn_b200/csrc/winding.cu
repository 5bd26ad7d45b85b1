// winding.cu -- the consistency pass of vtkPolyDataNormals (ConsistencyOn is VTK's default and what the
// reference runs: src/py/utils.py:493-501, src/app/fly_by_features.cxx:472-478).
//
// VTK walks the cells in waves from the lowest unvisited cell id; a neighbour across an edge (p1,p2) of the
// current cell must traverse that edge as (p2,p1), otherwise vtkPolyData::ReverseCell turns it into
// (p2,p1,p0).  On a consistently wound mesh the pass changes nothing -- the case of every mesh the reference
// ships models for -- so every build only DETECTS, with one kernel: a hash set of DIRECTED edges (one 64-bit CAS per
// edge); a directed edge that occurs twice means two cells traverse a shared edge the same way, the only thing the
// pass ever acts on.  Only then is the table of undirected edges built (key = the two point ids, value = the first
// two cells that use the edge with the direction each traverses it), which tells an inconsistent manifold edge from
// an edge that three or more cells share.  With flyby_norm_opts.consistency = 1 the pass is reproduced: union-find with parity over the
// manifold edges (root = lowest cell id of the component = VTK's seed, parity = reversed relative to the seed),
// cells with parity 1 are reversed in place and remembered in mesh.flip so that "vertex 0 of a cell" keeps
// meaning the caller's vertex 0 (the reference's point-id map is built from the ORIGINAL surface,
// utils.py:604-621).  A component without a consistent orientation (Moebius band) is refused.
//
// Deviation: edges shared by three or more cells are not crossed (VTK's NonManifoldTraversal would
// cross them in an order that depends on its wave sequence); they are reported, never reoriented.
//
// Integer / byte work, L2-atomic-bound: 12 B read per cell, 3 CAS per cell into a table of 8-byte keys (24 B per cell).
#include "ctx.h"

namespace fb {

#define WD_EMPTY 0xffffffffffffffffull
enum { WD_INCONSISTENT = 1, WD_NONMANIFOLD = 2, WD_NONORIENTABLE = 4, WD_DUPLICATE = 8 };

struct EdgeTab {
  unsigned long long* key;  // [cap] (lo << 32) | hi, WD_EMPTY = free
  int* cells;               // [cap, 2] (cell << 1) | (traverses lo -> hi), -1 = free, cells[.,1] == -2: three or more cells
  uint32_t mask;
};

__device__ __forceinline__ uint32_t wd_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (uint32_t)k;
}

__global__ void edge_insert_kernel(const int32_t* __restrict__ tris, int64_t T, EdgeTab tab) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
  const int32_t t[3] = {tris[3 * c], tris[3 * c + 1], tris[3 * c + 2]};
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const uint32_t a = (uint32_t)t[j], b = (uint32_t)t[(j + 1) % 3];
    if (a == b) continue;  // degenerate edge
    const uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
    const unsigned long long key = ((unsigned long long)lo << 32) | hi;
    const int val = (int)((c << 1) | (a < b ? 1 : 0));
    uint32_t h = wd_hash(key) & tab.mask;
    for (;;) {
      const unsigned long long old = atomicCAS(tab.key + h, WD_EMPTY, key);
      if (old == WD_EMPTY || old == key) break;
      h = (h + 1) & tab.mask;
    }
    if (atomicCAS(tab.cells + 2 * (size_t)h, -1, val) != -1)
      if (atomicCAS(tab.cells + 2 * (size_t)h + 1, -1, val) != -1) atomicExch(tab.cells + 2 * (size_t)h + 1, -2);
  }
}

// The check every build runs: a DIRECTED edge (a -> b) that occurs in two cells means two cells traverse a shared
// edge the same way -- the only thing the consistency pass ever acts on.  One 64-bit CAS per edge into a table of
// 8-byte keys; a consistently wound manifold (the normal case) never raises the flag and nothing else runs.
__global__ void edge_directed_kernel(const int32_t* __restrict__ tris, int64_t T, unsigned long long* keys, uint32_t mask,
                                     int* flags) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
  const int32_t t[3] = {tris[3 * c], tris[3 * c + 1], tris[3 * c + 2]};
  bool dup = false;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const uint32_t a = (uint32_t)t[j], b = (uint32_t)t[(j + 1) % 3];
    if (a == b) continue;
    const unsigned long long key = ((unsigned long long)a << 32) | b;
    uint32_t h = wd_hash(key) & mask;
    for (;;) {
      const unsigned long long old = atomicCAS(keys + h, WD_EMPTY, key);
      if (old == WD_EMPTY) break;
      if (old == key) { dup = true; break; }
      h = (h + 1) & mask;
    }
  }
  if (dup) atomicOr(flags, WD_DUPLICATE);
}

__global__ void edge_check_kernel(EdgeTab tab, int* flags) {
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h > tab.mask) return;
  const int x = tab.cells[2 * (size_t)h], y = tab.cells[2 * (size_t)h + 1];
  int f = 0;
  if (y == -2) f = WD_NONMANIFOLD;
  else if (x >= 0 && y >= 0 && ((x ^ y) & 1) == 0) f = WD_INCONSISTENT;  // both traverse the edge the same way
  if (__any_sync(0xffffffffu, f != 0)) {
    for (int o = 16; o > 0; o >>= 1) f |= __shfl_xor_sync(0xffffffffu, f, o);
    if ((threadIdx.x & 31) == 0 && (f & ~*(volatile int*)flags)) atomicOr(flags, f);
  }
}

// ---- union-find with parity: par[c] = (parent << 1) | (c reversed relative to its parent)
__global__ void orient_init_kernel(uint32_t* par, int64_t T) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c < T) par[c] = (uint32_t)c << 1;
}

__device__ __forceinline__ uint32_t orient_find(const uint32_t* par, uint32_t v) {  // (root << 1) | parity of v
  uint32_t acc = 0;
  for (;;) {
    const uint32_t p = __ldcg(par + v);
    if ((p >> 1) == v) return (v << 1) | acc;
    acc ^= p & 1u;
    v = p >> 1;
  }
}

__global__ void orient_union_kernel(EdgeTab tab, uint32_t* par, int* flags, int* changed) {
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h > tab.mask) return;
  const int x = tab.cells[2 * (size_t)h], y = tab.cells[2 * (size_t)h + 1];
  if (x < 0 || y < 0) return;  // boundary edge, or one that three or more cells share
  const uint32_t rel = ((x ^ y) & 1) == 0 ? 1u : 0u;  // same direction: exactly one of the two must be reversed
  const uint32_t fa = orient_find(par, (uint32_t)x >> 1), fb2 = orient_find(par, (uint32_t)y >> 1);
  const uint32_t ra = fa >> 1, rb = fb2 >> 1, q = (fa ^ fb2 ^ rel) & 1u;
  if (ra == rb) {
    if (q) atomicOr(flags, WD_NONORIENTABLE);
    return;
  }
  const uint32_t hi = ra > rb ? ra : rb, lo = ra > rb ? rb : ra;
  atomicMin(par + hi, (lo << 1) | q);  // every stored link is a valid relation, so losing a race only costs a round
  *changed = 1;
}

__global__ void orient_compress_kernel(uint32_t* par, int64_t T) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c < T) par[c] = orient_find(par, (uint32_t)c);
}

__global__ void orient_apply_kernel(const uint32_t* __restrict__ par, int64_t T, int32_t* tris, uint8_t* flip,
                                    unsigned long long* n_flipped) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool f = c < T && (par[c] & 1u);
  if (c < T) flip[c] = f ? 1 : 0;
  if (f) {  // vtkPolyData::ReverseCell: (p0,p1,p2) -> (p2,p1,p0)
    const int32_t t0 = tris[3 * c];
    tris[3 * c] = tris[3 * c + 2];
    tris[3 * c + 2] = t0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, f);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(n_flipped, (unsigned long long)__popc(m));
}

// Queues the cheap detection on the context stream; flags_dev (int) receives WD_DUPLICATE when some directed edge
// occurs twice.
int winding_detect(flyby_ctx* c, int* flags_dev) {
  MeshDev& m = c->mesh;
  if (m.T <= 0) return FLYBY_OK;
  uint64_t cap = 1024;
  while (cap < 4 * (uint64_t)m.T) cap <<= 1;
  FB_CUDA(c, m.edge_tab.reserve((size_t)cap * 16));  // sized for winding_classify as well
  FB_CUDA(c, cudaMemsetAsync(m.edge_tab.p, 0xff, (size_t)cap * 8, c->stream));
  edge_directed_kernel<<<(unsigned)cdiv(m.T, 256), 256, 0, c->stream>>>(m.tris.as<int32_t>(), m.T,
                                                                         m.edge_tab.as<unsigned long long>(),
                                                                         (uint32_t)(cap - 1), flags_dev);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// Only after a duplicate was seen: the table of UNDIRECTED edges with the first two cells of each, which tells an
// inconsistent manifold edge (WD_INCONSISTENT) from an edge that three or more cells share (WD_NONMANIFOLD).
// flags_host receives the bits (synchronises).
int winding_classify(flyby_ctx* c, int* flags_dev, int* flags_host) {
  MeshDev& m = c->mesh;
  uint64_t cap = 1024;
  while (cap < 4 * (uint64_t)m.T) cap <<= 1;
  EdgeTab tab{m.edge_tab.as<unsigned long long>(), reinterpret_cast<int*>(m.edge_tab.as<unsigned long long>() + cap),
              (uint32_t)(cap - 1)};
  FB_CUDA(c, cudaMemsetAsync(m.edge_tab.p, 0xff, (size_t)cap * 16, c->stream));
  FB_CUDA(c, cudaMemsetAsync(flags_dev, 0, sizeof(int), c->stream));
  edge_insert_kernel<<<(unsigned)cdiv(m.T, 256), 256, 0, c->stream>>>(m.tris.as<int32_t>(), m.T, tab);
  FB_LAUNCH_CHECK(c);
  edge_check_kernel<<<(unsigned)cdiv((int64_t)cap, 256), 256, 0, c->stream>>>(tab, flags_dev);
  FB_LAUNCH_CHECK(c);
  FB_CUDA(c, cudaMemcpyAsync(flags_host, flags_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return FLYBY_OK;
}

// Reorients the cells (consistency = 1).  Synchronises: this is the rare path of a mesh that needs it.
int winding_fix(flyby_ctx* c, int* flags_dev) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  uint64_t cap = 1024;
  while (cap < 4 * (uint64_t)m.T) cap <<= 1;
  EdgeTab tab{m.edge_tab.as<unsigned long long>(), reinterpret_cast<int*>(m.edge_tab.as<unsigned long long>() + cap),
              (uint32_t)(cap - 1)};
  FB_CUDA(c, m.orient.reserve(sizeof(uint32_t) * (size_t)m.T + 16));
  FB_CUDA(c, m.flip.reserve((size_t)m.T));
  uint32_t* par = m.orient.as<uint32_t>();
  unsigned long long* scratch = c->counters.as<unsigned long long>() + 15;  // changed flag / flip count
  const unsigned gt = (unsigned)cdiv(m.T, 256), ge = (unsigned)cdiv((int64_t)cap, 256);
  orient_init_kernel<<<gt, 256, 0, st>>>(par, m.T);
  FB_LAUNCH_CHECK(c);
  for (int64_t round = 0;; round++) {
    if (round > m.T) return fail(c, FLYBY_ERR_CUDA, "winding: orientation did not converge");
    FB_CUDA(c, cudaMemsetAsync(scratch, 0, 8, st));
    orient_union_kernel<<<ge, 256, 0, st>>>(tab, par, flags_dev, reinterpret_cast<int*>(scratch));
    FB_LAUNCH_CHECK(c);
    orient_compress_kernel<<<gt, 256, 0, st>>>(par, m.T);
    FB_LAUNCH_CHECK(c);
    int changed = 0;
    FB_CUDA(c, cudaMemcpyAsync(&changed, scratch, sizeof(int), cudaMemcpyDeviceToHost, st));
    FB_CUDA(c, cudaStreamSynchronize(st));
    if (!changed) break;
  }
  int flags = 0;
  FB_CUDA(c, cudaMemcpyAsync(&flags, flags_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  if (flags & WD_NONORIENTABLE)
    return fail(c, FLYBY_ERR_UNSUPPORTED, "build: the mesh has a component without a consistent orientation (non-orientable)");
  FB_CUDA(c, cudaMemsetAsync(scratch, 0, 8, st));
  orient_apply_kernel<<<gt, 256, 0, st>>>(par, m.T, m.tris.as<int32_t>(), m.flip.as<uint8_t>(), scratch);
  FB_LAUNCH_CHECK(c);
  unsigned long long n = 0;
  FB_CUDA(c, cudaMemcpyAsync(&n, scratch, 8, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  m.n_flipped = (int64_t)n;
  return FLYBY_OK;
}

}  // namespace fb
