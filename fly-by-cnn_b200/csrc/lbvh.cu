// lbvh.cu -- GPU LBVH build: Morton codes, hand-written LSD radix sort (no CUB),
// Karras radix-tree topology, bottom-up AABB refit, top-of-tree-first node
// layout and the Morton-ordered SoA triangle arrays.
//
// Replaces vtkCellLocator::BuildLocator (reference
// src/app/fly_by_features.cxx:484-486, src/py/predict.py:70-72).
//
// Every kernel is an HBM/L2-bound integer pass over T keys or T-1 nodes.
// Bytes per triangle over the whole build: codes 4 passes x
// (8 B read for the histogram + 8 B read + 8 B write for the scatter) = 96 B,
// tree 64 B node + 48 B SoA vertices + 16 B ids written once.
#include "ctx.h"

namespace fb {

const float* prep_unit_box(flyby_ctx* c);

// ============================================================ exclusive scan
// 256 threads x 8 items per block; recursion on the block totals.
#define SCAN_ITEMS 8
#define SCAN_TILE (256 * SCAN_ITEMS)

__global__ void __launch_bounds__(256) scan_tile_kernel(uint32_t* data, int64_t n, uint32_t* block_sums) {
  __shared__ uint32_t warp_tot[8];
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? data[base + k] : 0u;
    sum += v[k];
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < 8; w++)
    if (w < warp) woff += warp_tot[w];
  uint32_t run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 255 && block_sums) block_sums[blockIdx.x] = woff + incl;
}

__global__ void scan_add_kernel(uint32_t* data, int64_t n, const uint32_t* __restrict__ block_offs) {
  int64_t i = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
  uint32_t off = block_offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++, i += 256)
    if (i < n) data[i] += off;
}

static int scan_rec(flyby_ctx* c, uint32_t* data, int64_t n, uint32_t* tmp, cudaStream_t st) {
  int64_t nb = cdiv(n, SCAN_TILE);
  scan_tile_kernel<<<(unsigned)nb, 256, 0, st>>>(data, n, nb > 1 ? tmp : nullptr);
  FB_LAUNCH_CHECK(c);
  if (nb > 1) {
    int rc = scan_rec(c, tmp, nb, tmp + nb, st);
    if (rc) return rc;
    scan_add_kernel<<<(unsigned)nb, 256, 0, st>>>(data, n, tmp);
    FB_LAUNCH_CHECK(c);
  }
  return FLYBY_OK;
}

int wait_tree(flyby_ctx* c) {
  if (c->tree_pending) FB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_tree, 0));
  return FLYBY_OK;
}

int exclusive_scan_u32(flyby_ctx* c, uint32_t* data, int64_t n, DevBuf& tmp, cudaStream_t st) {
  if (!st) st = c->stream;
  if (n <= 0) return FLYBY_OK;
  // total temp over all recursion levels < n/2048 + n/2048^2 + ... + 8
  size_t need = 0;
  for (int64_t m = cdiv(n, SCAN_TILE); m > 1; m = cdiv(m, SCAN_TILE)) need += (size_t)m;
  need += 16;
  FB_CUDA(c, tmp.reserve(need * sizeof(uint32_t)));
  return scan_rec(c, data, n, tmp.as<uint32_t>(), st);
}

// ============================================================ Morton codes
__global__ void morton_kernel(const float* __restrict__ v, const int32_t* __restrict__ tris, int64_t T,
                              const float* __restrict__ ubox, uint32_t* __restrict__ codes,
                              uint32_t* __restrict__ idx) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
  int i0 = tris[3 * c], i1 = tris[3 * c + 1], i2 = tris[3 * c + 2];
  float q[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float a = v[3 * (int64_t)i0 + k], b = v[3 * (int64_t)i1 + k], e = v[3 * (int64_t)i2 + k];
    float mn = fminf(a, fminf(b, e)), mx = fmaxf(a, fmaxf(b, e));
    float ext = ubox[3 + k] - ubox[k];
    q[k] = ext > 0.f ? (0.5f * (mn + mx) - ubox[k]) / ext : 0.5f;
  }
  codes[c] = morton30(q[0], q[1], q[2]);
  idx[c] = (uint32_t)c;
}

// ============================================================ radix sort
// Stable LSD sort of (code, index) pairs.  The Morton codes have 30 bits: three passes of 10 bits.  Per pass:
//   radix_hist_kernel   per-block digit histogram -> hist[digit][block]
//   exclusive scan      over the digit-major table = global base per (digit, block)
//   radix_scatter_kernel stable ranking inside the block with warp match votes
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)  // 2048 keys per block, 256 per warp
#define RS_BITS 10
#define RS_RADIX (1 << RS_BITS)
#define RS_PASSES 3

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist,
                  int nblocks) {
  __shared__ uint32_t h[RS_RADIX];
  for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) h[d] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; k++) {
    int64_t i = base + k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (RS_RADIX - 1u)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) hist[(int64_t)d * nblocks + blockIdx.x] = h[d];
}

__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
                     int shift, const uint32_t* __restrict__ hist_scanned, int nblocks,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  // per-warp digit counts, turned into per-warp bases, then used as running cursors
  __shared__ uint32_t cnt[8][RS_RADIX];
  __shared__ uint32_t gbase[RS_RADIX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8 * RS_RADIX; i += RS_THREADS) (&cnt[0][0])[i] = 0;
  for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) gbase[d] = hist_scanned[(int64_t)d * nblocks + blockIdx.x];
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_TILE / 8);
  uint32_t k[RS_ITEMS], v[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int64_t i = wbase + r * 32 + lane;
    bool ok = i < n;
    k[r] = ok ? keys[i] : 0u;
    v[r] = ok ? vals[i] : 0u;
    if (ok) atomicAdd(&cnt[warp][(k[r] >> shift) & (RS_RADIX - 1u)], 1u);
  }
  __syncthreads();
  // exclusive prefix over warps for each digit (every thread owns RS_RADIX / RS_THREADS digits)
  for (int d = threadIdx.x; d < RS_RADIX; d += RS_THREADS) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      uint32_t t = cnt[w][d];
      cnt[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    int64_t i = wbase + r * 32 + lane;
    bool ok = i < n;
    uint32_t d = (k[r] >> shift) & (RS_RADIX - 1u);
    uint32_t peers = __match_any_sync(0xffffffffu, ok ? d : (RS_RADIX | (uint32_t)lane));
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (ok) pos = gbase[d] + cnt[warp][d] + rank;
    __syncwarp();
    if (ok && rank == 0) cnt[warp][d] += __popc(peers);
    __syncwarp();
    if (ok) {
      keys_out[pos] = k[r];
      vals_out[pos] = v[r];
    }
  }
}

// ============================================================ tree
// sorted slot -> SoA vertices, ids and leaf boxes
__global__ void reorder_kernel(const float* __restrict__ v, const int32_t* __restrict__ tris, int64_t T,
                               const uint32_t* __restrict__ idx, float4* __restrict__ tv0,
                               float4* __restrict__ tv1, float4* __restrict__ tv2,
                               int4* __restrict__ tvid, float4* __restrict__ trec, float* __restrict__ leaf_box) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= T) return;
  int cell = (int)idx[s];
  int i0 = tris[3 * (int64_t)cell], i1 = tris[3 * (int64_t)cell + 1], i2 = tris[3 * (int64_t)cell + 2];
  float p0[3], p1[3], p2[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    p0[k] = v[3 * (int64_t)i0 + k];
    p1[k] = v[3 * (int64_t)i1 + k];
    p2[k] = v[3 * (int64_t)i2 + k];
    leaf_box[6 * s + k] = fminf(p0[k], fminf(p1[k], p2[k]));
    leaf_box[6 * s + 3 + k] = fmaxf(p0[k], fmaxf(p1[k], p2[k]));
  }
  tv0[s] = make_float4(p0[0], p0[1], p0[2], __int_as_float(cell));
  tv1[s] = make_float4(p1[0], p1[1], p1[2], 0.f);
  tv2[s] = make_float4(p2[0], p2[1], p2[2], 0.f);
  tvid[s] = make_int4(i0, i1, i2, cell);
  // 64-byte record of the same data for the per-pixel kernels (one fetch instead of four)
  trec[4 * s + 0] = make_float4(p0[0], p0[1], p0[2], __int_as_float(cell));
  trec[4 * s + 1] = make_float4(p1[0], p1[1], p1[2], __int_as_float(i0));
  trec[4 * s + 2] = make_float4(p2[0], p2[1], p2[2], __int_as_float(i1));
  trec[4 * s + 3] = make_float4(__int_as_float(i2), 0.f, 0.f, 0.f);
}

// knode[i] = (child0, child1, parent, arrival flag); leaf children are ~slot
__global__ void karras_kernel(const uint32_t* __restrict__ codes, int n, int4* __restrict__ knode,
                              int32_t* __restrict__ leaf_parent, int2* __restrict__ krange) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int first, last, split;
  karras_node(codes, n, i, &first, &last, &split);
  krange[i] = make_int2(first, last - first + 1);  // the subtree covers slots [first, first + count)
  int c0 = (split == first) ? ~split : split;
  int c1 = (split + 1 == last) ? ~(split + 1) : split + 1;
  knode[i].x = c0;
  knode[i].y = c1;
  knode[i].w = 0;
  if (i == 0) knode[0].z = -1;
  if (c0 >= 0) knode[c0].z = i; else leaf_parent[~c0] = i;
  if (c1 >= 0) knode[c1].z = i; else leaf_parent[~c1] = i;
}

// bottom-up: the second thread to reach a node merges its children's boxes
// Refit without a dependency chain: the subtree of internal node i covers the Morton-sorted slots
// krange[i] = [first, first + count), so its box is a range union of leaf boxes.  Two levels of pre-reduced
// boxes (32-leaf chunks, 1024-leaf super chunks) keep a range query at <= 62 leaves + 62 chunks + count / 1024
// super chunks; min / max are exact and order independent, so the boxes equal the bottom-up ones bit for bit.
__global__ void __launch_bounds__(256) box_reduce32_kernel(const float* __restrict__ in, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float b[6];
#pragma unroll
  for (int k = 0; k < 6; k++) b[k] = i < n ? in[6 * (int64_t)i + k] : (k < 3 ? INFINITY : -INFINITY);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const float v = __shfl_xor_sync(0xffffffffu, b[k], o);
      b[k] = k < 3 ? fminf(b[k], v) : fmaxf(b[k], v);
    }
  if ((threadIdx.x & 31) == 0 && i < n)
#pragma unroll
    for (int k = 0; k < 6; k++) out[6 * (int64_t)(i >> 5) + k] = b[k];
}

__device__ __forceinline__ void box_merge(float* b, const float* __restrict__ src, int64_t idx) {
  const float* p = src + 6 * idx;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    b[k] = fminf(b[k], __ldg(p + k));
    b[3 + k] = fmaxf(b[3 + k], __ldg(p + 3 + k));
  }
}

__global__ void __launch_bounds__(256) refit_range_kernel(int n_internal, const int2* __restrict__ krange,
                                                          const float* __restrict__ leaf_box,
                                                          const float* __restrict__ chunk_box,
                                                          const float* __restrict__ super_box, float* __restrict__ node_box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  const int2 r = krange[i];
  int a = r.x;
  const int e = r.x + r.y;  // [a, e)
  float b[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
  // leaves up to the next chunk boundary, chunks up to the next super boundary, supers, then back down
  while (a < e && (a & 31)) box_merge(b, leaf_box, a++);
  while (a + 32 <= e && (a & 1023)) { box_merge(b, chunk_box, a >> 5); a += 32; }
  while (a + 1024 <= e) { box_merge(b, super_box, a >> 10); a += 1024; }
  while (a + 32 <= e) { box_merge(b, chunk_box, a >> 5); a += 32; }
  while (a < e) box_merge(b, leaf_box, a++);
#pragma unroll
  for (int k = 0; k < 6; k++) node_box[6 * (int64_t)i + k] = b[k];
}

// breadth-first prefix of the tree (at most ntop_max nodes) -> remap[karras] = rank
// One block of 1024 threads; every level is appended with a block-wide scan so
// the order is deterministic.
__global__ void __launch_bounds__(1024)
top_bfs_kernel(const int4* __restrict__ knode, int n_internal, int ntop_max, int32_t* remap,
               int32_t* topflag, int32_t* n_top_out) {
  __shared__ int32_t order[1024];
  __shared__ int32_t warp_tot[32];
  __shared__ int head_s, count_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { order[0] = 0; head_s = 0; count_s = 1; }
  __syncthreads();
  for (;;) {
    int head = head_s, count = count_s;
    if (head >= count || count >= ntop_max) break;
    int item = head + tid;
    int c0 = -1, c1 = -1, k = 0;
    if (item < count) {
      int4 kn = knode[order[item]];
      if (kn.x >= 0) { c0 = kn.x; k++; }
      if (kn.y >= 0) { c1 = kn.y; k++; }
    }
    int incl = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < 32; w++) {
      if (w < warp) woff += warp_tot[w];
      total += warp_tot[w];
    }
    int pos = count + woff + incl - k;
    if (c0 >= 0) { if (pos < ntop_max) order[pos] = c0; pos++; }
    if (c1 >= 0) { if (pos < ntop_max) order[pos] = c1; }
    __syncthreads();
    if (tid == 0) {
      head_s = count;
      int nc = count + total;
      count_s = nc < ntop_max ? nc : ntop_max;
    }
    __syncthreads();
  }
  int count = count_s;
  if (count > n_internal) count = n_internal;
  for (int i = tid; i < count; i += blockDim.x) {
    remap[order[i]] = i;
    topflag[order[i]] = 0;  // 0 = top (not counted by the non-top scan)
  }
  if (tid == 0) *n_top_out = count;
}

__global__ void fill_u32_kernel(uint32_t* p, int64_t n, uint32_t v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// final index of a Karras node: top nodes keep their BFS rank, the rest follow
// in Karras order (which keeps subtrees contiguous)
__device__ __forceinline__ int final_index(int k, const int32_t* remap, const uint32_t* nontop_scan,
                                           const int32_t* nontop_flag, int n_top) {
  return nontop_flag[k] ? n_top + (int)nontop_scan[k] : remap[k];
}

// child reference in the final layout: a subtree of at most leaf_max triangles becomes one leaf
// (its slots are contiguous), anything larger stays an internal node
__device__ __forceinline__ int final_child(int k, const int2* krange, int leaf_max, const int32_t* remap,
                                           const uint32_t* nontop_scan, const int32_t* nontop_flag, int n_top) {
  if (k < 0) return k;  // single-triangle leaf: ~slot == leaf_encode(slot, 1)
  const int2 r = krange[k];
  if (r.y <= leaf_max) return leaf_encode(r.x, r.y);
  return final_index(k, remap, nontop_scan, nontop_flag, n_top);
}

__global__ void layout_kernel(int n_internal, const int4* __restrict__ knode, const int2* __restrict__ krange,
                              int leaf_max, const float* __restrict__ leaf_box, const float* __restrict__ node_box,
                              const int32_t* __restrict__ remap, const uint32_t* __restrict__ nontop_scan,
                              const int32_t* __restrict__ nontop_flag, const int32_t* __restrict__ n_top_p,
                              Node* __restrict__ nodes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  const int n_top = *n_top_p;
  int4 kn = knode[i];
  Node out;
  const float* b0 = kn.x < 0 ? leaf_box + 6 * (int64_t)(~kn.x) : node_box + 6 * (int64_t)kn.x;
  const float* b1 = kn.y < 0 ? leaf_box + 6 * (int64_t)(~kn.y) : node_box + 6 * (int64_t)kn.y;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    out.box[k] = b0[k];
    out.box[6 + k] = b1[k];
  }
  out.c0 = final_child(kn.x, krange, leaf_max, remap, nontop_scan, nontop_flag, n_top);
  out.c1 = final_child(kn.y, krange, leaf_max, remap, nontop_scan, nontop_flag, n_top);
  int sib = -1;
  if (kn.z >= 0) {
    int4 pn = knode[kn.z];
    int s = (pn.x == i) ? pn.y : pn.x;
    sib = final_child(s, krange, leaf_max, remap, nontop_scan, nontop_flag, n_top);
    out.parent = final_index(kn.z, remap, nontop_scan, nontop_flag, n_top);
  } else {
    out.parent = -1;
  }
  out.sibling = sib;
  nodes[final_index(i, remap, nontop_scan, nontop_flag, n_top)] = out;
}

// a mesh with a single triangle: one node with the leaf as both children (the
// duplicate test of the same cell cannot change the result)
__global__ void single_leaf_kernel(const float* __restrict__ leaf_box, Node* nodes) {
  if (threadIdx.x != 0) return;
  Node out;
  for (int k = 0; k < 6; k++) { out.box[k] = leaf_box[k]; out.box[6 + k] = leaf_box[k]; }
  out.c0 = ~0;
  out.c1 = ~0;
  out.parent = -1;
  out.sibling = -1;
  nodes[0] = out;
}

// ============================================================ driver
#define FB_NTOP_MAX 1024

int lbvh_run(flyby_ctx* c) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t T = m.T;
  m.n_nodes = 0;
  m.n_top_dev = nullptr;
  if (T == 0) return FLYBY_OK;
  const int nblocks = (int)cdiv(T, RS_TILE);
  const size_t Tz = (size_t)T;
  FB_CUDA(c, m.codes.reserve(4 * Tz));
  FB_CUDA(c, m.codes_alt.reserve(4 * Tz));
  FB_CUDA(c, m.idx.reserve(4 * Tz));
  FB_CUDA(c, m.idx_alt.reserve(4 * Tz));
  FB_CUDA(c, m.hist.reserve(4 * RS_RADIX * (size_t)nblocks));
  FB_CUDA(c, m.leaf_box.reserve(24 * Tz));
  FB_CUDA(c, m.node_box.reserve(24 * Tz));
  FB_CUDA(c, m.knode.reserve(16 * Tz));
  FB_CUDA(c, m.krange.reserve(8 * Tz));
  FB_CUDA(c, m.remap.reserve(4 * Tz + 64));
  FB_CUDA(c, m.topflag.reserve(4 * Tz + 64));
  FB_CUDA(c, m.nodes.reserve(sizeof(Node) * Tz));
  FB_CUDA(c, m.tv0.reserve(16 * Tz));
  FB_CUDA(c, m.tv1.reserve(16 * Tz));
  FB_CUDA(c, m.tv2.reserve(16 * Tz));
  FB_CUDA(c, m.tvid.reserve(16 * Tz));
  FB_CUDA(c, m.trec.reserve(64 * Tz));

  const unsigned gT = (unsigned)cdiv(T, 256);
  morton_kernel<<<gT, 256, 0, st>>>(m.verts.as<float>(), m.tris.as<int32_t>(), T, prep_unit_box(c),
                                    m.codes.as<uint32_t>(), m.idx.as<uint32_t>());
  FB_LAUNCH_CHECK(c);
  uint32_t *k0 = m.codes.as<uint32_t>(), *k1 = m.codes_alt.as<uint32_t>();
  uint32_t *v0 = m.idx.as<uint32_t>(), *v1 = m.idx_alt.as<uint32_t>();
  for (int pass = 0; pass < RS_PASSES; pass++) {
    int shift = RS_BITS * pass;
    radix_hist_kernel<<<nblocks, RS_THREADS, 0, st>>>(k0, T, shift, m.hist.as<uint32_t>(), nblocks);
    FB_LAUNCH_CHECK(c);
    int rc = exclusive_scan_u32(c, m.hist.as<uint32_t>(), RS_RADIX * (int64_t)nblocks, m.scan_tmp);
    if (rc) return rc;
    radix_scatter_kernel<<<nblocks, RS_THREADS, 0, st>>>(k0, v0, T, shift, m.hist.as<uint32_t>(),
                                                         nblocks, k1, v1);
    FB_LAUNCH_CHECK(c);
    uint32_t* t = k0; k0 = k1; k1 = t;
    t = v0; v0 = v1; v1 = t;
  }
  // after an odd number of passes the sorted data sits in the alternate buffers: k0 / v0 point at it
  FB_STAT_EVENT(c, c->ev[3], st);

  reorder_kernel<<<gT, 256, 0, st>>>(m.verts.as<float>(), m.tris.as<int32_t>(), T, v0,
                                     m.tv0.as<float4>(), m.tv1.as<float4>(), m.tv2.as<float4>(),
                                     m.tvid.as<int4>(), m.trec.as<float4>(), m.leaf_box.as<float>());
  FB_LAUNCH_CHECK(c);
  if (T == 1) {
    single_leaf_kernel<<<1, 32, 0, st>>>(m.leaf_box.as<float>(), m.nodes.as<Node>());
    FB_LAUNCH_CHECK(c);
    fill_u32_kernel<<<1, 32, 0, st>>>(m.remap.as<uint32_t>(), 1, 1u);
    FB_LAUNCH_CHECK(c);
    m.n_nodes = 1;
    m.n_top_dev = m.remap.as<int32_t>();
    FB_STAT_EVENT(c, c->ev[4], st);
    return FLYBY_OK;
  }
  const int n = (int)T, ni = n - 1;
  // From here on nothing is needed by the scan-conversion render: topology, boxes and layout go to the side
  // stream, ordered behind the sort / reorder by ev_sorted; readers of the tree call wait_tree().
  cudaStream_t ss = c->side_stream ? c->side_stream : st;
  if (ss != st) {
    FB_CUDA(c, cudaEventRecord(c->ev_sorted, st));
    FB_SIDE_WAIT(c, c->ev_sorted);
  }
  int32_t* leaf_parent = reinterpret_cast<int32_t*>(v1);  // the sort's spare index buffer is free again
  karras_kernel<<<(unsigned)cdiv(ni, 256), 256, 0, ss>>>(k0, n, m.knode.as<int4>(), leaf_parent, m.krange.as<int2>());
  FB_LAUNCH_CHECK(c);
  {
    const int n_chunk = (n + 31) / 32, n_super = (n_chunk + 31) / 32;
    FB_CUDA(c, m.hist.reserve(sizeof(float) * 6 * ((size_t)n_chunk + (size_t)n_super) + 64));  // sort scratch, free again
    float* chunk_box = m.hist.as<float>();
    float* super_box = chunk_box + 6 * (size_t)n_chunk;
    box_reduce32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, ss>>>(m.leaf_box.as<float>(), n, chunk_box);
    FB_LAUNCH_CHECK(c);
    box_reduce32_kernel<<<(unsigned)cdiv(n_chunk, 256), 256, 0, ss>>>(chunk_box, n_chunk, super_box);
    FB_LAUNCH_CHECK(c);
    refit_range_kernel<<<(unsigned)cdiv(ni, 256), 256, 0, ss>>>(ni, m.krange.as<int2>(), m.leaf_box.as<float>(), chunk_box,
                                                                 super_box, m.node_box.as<float>());
    FB_LAUNCH_CHECK(c);
  }
  // top-of-tree first: flag = 1 for every node, the BFS clears it for the top set
  int32_t* n_top_dev = m.remap.as<int32_t>() + ni;  // one spare word behind the table
  fill_u32_kernel<<<(unsigned)cdiv(ni, 256), 256, 0, ss>>>(m.topflag.as<uint32_t>(), ni, 1u);
  FB_LAUNCH_CHECK(c);
  top_bfs_kernel<<<1, 1024, 0, ss>>>(m.knode.as<int4>(), ni, FB_NTOP_MAX, m.remap.as<int32_t>(),
                                     m.topflag.as<int32_t>(), n_top_dev);
  FB_LAUNCH_CHECK(c);
  // scan of the non-top flags (kept separately: the layout needs flag and rank)
  uint32_t* nontop_scan = k1;  // the sort's spare key buffer
  FB_CUDA(c, cudaMemcpyAsync(nontop_scan, m.topflag.p, 4 * (size_t)ni, cudaMemcpyDeviceToDevice, ss));
  int rc = exclusive_scan_u32(c, nontop_scan, ni, m.tree_scan_tmp, ss);
  if (rc) return rc;
  layout_kernel<<<(unsigned)cdiv(ni, 256), 256, 0, ss>>>(
      ni, m.knode.as<int4>(), m.krange.as<int2>(), c->leaf_max, m.leaf_box.as<float>(), m.node_box.as<float>(),
      m.remap.as<int32_t>(),
      nontop_scan, m.topflag.as<int32_t>(), n_top_dev, m.nodes.as<Node>());
  FB_LAUNCH_CHECK(c);
  FB_STAT_EVENT(c, c->ev[4], ss);
  m.n_nodes = ni;
  m.n_top_dev = n_top_dev;
  return FLYBY_OK;
}

}  // namespace fb
