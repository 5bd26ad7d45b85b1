// label.cu -- point-feature lookup through the cell-id map and the
// universal-labeling back-projection (per-pixel CNN labels -> mesh cells).
//
// Reference: src/py/utils.py:604-684 (id map + ExtractPointFeatures),
// src/py/dental_model_seg.py:192-211, src/py/predict_v3.py:59-80,
// src/py/predict.py:154-169, src/py/challenge-teeth/teeth_challenge_predict.py:60-72.
// The reference's scatter loses votes on duplicate indices (index_put with
// repeats) and counts background pixels into the last row; here every vote is
// an integer atomic, so the result is exact and independent of the order of
// pixels, views and GPUs.
//
// HBM-bound byte work: 8 B read per pixel (id + label), one 4 B L2 atomic per
// (cell, label) group of a warp; votes table [cells, K] int32 stays in L2.
#include "ctx.h"

namespace fb {

// mode 0: ids are hit CELL ids, the value is taken at the cell's vertex 0 (GetPointIdColors,
//         utils.py:604-621, colours every cell with the id of its first point);
// mode 1: ids are POINT ids already (decoded from an RGB id map, utils.py:644-662).
// "vertex 0 of cell c" as the caller numbered the cell: a cell the consistency pass reversed (winding.cu) is
// stored as (p2,p1,p0)
__device__ __forceinline__ int32_t cell_vertex0(const int32_t* __restrict__ tris, const uint8_t* __restrict__ flip, int64_t c) {
  return tris[3 * c + ((flip && flip[c]) ? 2 : 0)];
}

__global__ void point_features_kernel(const int32_t* __restrict__ ids, int64_t n_pix,
                                      const int32_t* __restrict__ tris, const uint8_t* __restrict__ flip, int64_t n_cells,
                                      int64_t n_points, const float* __restrict__ feats, int F, float zero, int mode,
                                      float* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pix * F) return;
  int64_t p = i / F;
  int k = (int)(i % F);
  int64_t c = ids[p];
  int64_t pt = -1;
  if (mode == 0) { if (c >= 0 && c < n_cells) pt = cell_vertex0(tris, flip, c); }
  else if (c >= 0 && c < n_points) pt = c;
  out[i] = pt >= 0 ? feats[pt * F + k] : zero;
}

// the decoded point-id map of GetPointIdMapActor (utils.py:604-637): vertex 0 of the hit cell, -1 = background
__global__ void point_ids_kernel(const int32_t* __restrict__ ids, int64_t n_pix, const int32_t* __restrict__ tris,
                                 const uint8_t* __restrict__ flip, int64_t n_cells, int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const int64_t c = ids[i];
  out[i] = (c >= 0 && c < n_cells) ? cell_vertex0(tris, flip, c) : -1;
}

// votes[cell, label] += 1.  Lanes of a warp that target the same counter are
// merged with a match vote so that one atomic carries the whole group.
// tris != nullptr: per-POINT votes (predict_v3.py:59-69, predict.py:121-157) -- the row is the vertex 0 of the hit cell.
__global__ void backproject_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ labels,
                                   int64_t n_pix, int K, int64_t n_cells, int32_t* votes,
                                   const int32_t* __restrict__ tris, const uint8_t* __restrict__ flip) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int32_t c = -1, l = -1;
  if (i < n_pix) {
    c = ids[i];
    l = labels[i];
  }
  bool ok = c >= 0 && c < n_cells && l >= 0 && l < K;
  if (ok && tris) c = cell_vertex0(tris, flip, c);
  unsigned long long key = ok ? (unsigned long long)c * (unsigned)K + (unsigned)l
                              : (0x8000000000000000ull | (unsigned)lane);
  unsigned peers = __match_any_sync(0xffffffffu, key);
  if (ok && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&votes[key], __popc(peers));
}

// soft votes in 2^-20 fixed point (integer atomics keep the sum order-free)
__global__ void backproject_soft_kernel(const int32_t* __restrict__ ids, const float* __restrict__ probs,
                                        int64_t n_pix, int K, int64_t n_cells,
                                        unsigned long long* votes_fx) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pix * K) return;
  int64_t p = i / K;
  int k = (int)(i % K);
  int32_t c = ids[p];
  if (c < 0 || c >= n_cells) return;
  long long q = llrintf(probs[i] * 1048576.0f);
  if (q != 0) atomicAdd(&votes_fx[(int64_t)c * K + k], (unsigned long long)q);
}

__global__ void votes_argmax_kernel(const int32_t* __restrict__ votes, int64_t n_cells, int K, int32_t dflt,
                                    int32_t* __restrict__ out) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  int32_t best = 0, arg = -1;
  for (int k = 0; k < K; k++) {
    int32_t v = votes[c * K + k];
    if (v > best) { best = v; arg = k; }
  }
  out[c] = arg >= 0 ? arg : dflt;
}

// cell label -> vertex 0 of the cell; the highest cell id wins (the reference
// loop overwrites in ascending cell order, dental_model_seg.py:210-211)
__global__ void cell_owner_kernel(const int32_t* __restrict__ tris, const uint8_t* __restrict__ flip, int64_t T, int32_t* owner) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c < T) atomicMax(&owner[cell_vertex0(tris, flip, c)], (int32_t)c);
}
__global__ void owner_label_kernel(const int32_t* __restrict__ owner, const int32_t* __restrict__ cell_labels,
                                   int64_t V, int32_t dflt, int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < V) out[i] = owner[i] >= 0 ? cell_labels[owner[i]] : dflt;
}

int label_point_features(flyby_ctx* c, const int32_t* ids, int64_t n_pix, const float* feats, int F,
                         float zero, int mode, float* out) {
  if (n_pix * F <= 0) return FLYBY_OK;
  point_features_kernel<<<(unsigned)cdiv(n_pix * F, 256), 256, 0, c->stream>>>(
      ids, n_pix, c->mesh.tris.as<int32_t>(), c->mesh.n_flipped > 0 ? c->mesh.flip.as<uint8_t>() : nullptr, c->mesh.T,
      c->mesh.V, feats, F, zero, mode, out);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

int label_point_ids(flyby_ctx* c, const int32_t* ids, int64_t n_pix, int32_t* out) {
  if (n_pix <= 0) return FLYBY_OK;
  point_ids_kernel<<<(unsigned)cdiv(n_pix, 256), 256, 0, c->stream>>>(
      ids, n_pix, c->mesh.tris.as<int32_t>(), c->mesh.n_flipped > 0 ? c->mesh.flip.as<uint8_t>() : nullptr, c->mesh.T, out);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

int label_backproject_points(flyby_ctx* c, const int32_t* ids, const int32_t* labels, int64_t n_pix, int K, int32_t* votes) {
  if (n_pix <= 0) return FLYBY_OK;
  backproject_kernel<<<(unsigned)cdiv(n_pix, 256), 256, 0, c->stream>>>(
      ids, labels, n_pix, K, c->mesh.T, votes, c->mesh.tris.as<int32_t>(),
      c->mesh.n_flipped > 0 ? c->mesh.flip.as<uint8_t>() : nullptr);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

int label_backproject(flyby_ctx* c, const int32_t* ids, const int32_t* labels, int64_t n_pix, int K,
                      int32_t* votes) {
  if (n_pix <= 0) return FLYBY_OK;
  backproject_kernel<<<(unsigned)cdiv(n_pix, 256), 256, 0, c->stream>>>(ids, labels, n_pix, K, c->mesh.T, votes, nullptr, nullptr);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

int label_backproject_soft(flyby_ctx* c, const int32_t* ids, const float* probs, int64_t n_pix, int K,
                           int64_t* votes_fx) {
  if (n_pix <= 0) return FLYBY_OK;
  backproject_soft_kernel<<<(unsigned)cdiv(n_pix * K, 256), 256, 0, c->stream>>>(
      ids, probs, n_pix, K, c->mesh.T, reinterpret_cast<unsigned long long*>(votes_fx));
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

int label_argmax(flyby_ctx* c, const int32_t* votes, int64_t n_cells, int K, int32_t dflt, int32_t* out) {
  if (n_cells <= 0) return FLYBY_OK;
  votes_argmax_kernel<<<(unsigned)cdiv(n_cells, 256), 256, 0, c->stream>>>(votes, n_cells, K, dflt, out);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

int label_cells_to_points(flyby_ctx* c, const int32_t* cell_labels, int32_t dflt, int32_t* out) {
  MeshDev& m = c->mesh;
  FB_CUDA(c, c->stage_misc.reserve(sizeof(int32_t) * (size_t)m.V));
  int32_t* owner = c->stage_misc.as<int32_t>();
  FB_CUDA(c, cudaMemsetAsync(owner, 0xff, sizeof(int32_t) * (size_t)m.V, c->stream));
  if (m.T > 0) {
    cell_owner_kernel<<<(unsigned)cdiv(m.T, 256), 256, 0, c->stream>>>(
        m.tris.as<int32_t>(), m.n_flipped > 0 ? m.flip.as<uint8_t>() : nullptr, m.T, owner);
    FB_LAUNCH_CHECK(c);
  }
  owner_label_kernel<<<(unsigned)cdiv(m.V, 256), 256, 0, c->stream>>>(owner, cell_labels, m.V, dflt, out);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

}  // namespace fb
