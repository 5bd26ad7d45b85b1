// prep.cu -- mesh normalisation (unit surface) and point normals on the GPU.
//
// Replaces GetUnitSurf/ScaleSurf (reference src/py/utils.py:249-290,341;
// C++ src/app/fly_by_features.cxx:393-450) and ComputeNormals /
// vtkPolyDataNormals with splitting off (src/py/utils.py:493-501;
// src/app/fly_by_features.cxx:472-478).
//
// Both stages are HBM-bound streaming passes over V points / T cells
// (12 B per point read + 12 B written; 12 B per cell index read + 36 B of
// gathered vertices); they are a few microseconds next to the ray cast.
// Results are bit-reproducible: min/max reductions are order-free, the centre
// of mass uses a fixed chunked summation order, and point normals add the
// float32 face normals in ascending cell id through a CSR adjacency.
#include "ctx.h"

namespace fb {

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// red layout (uint32 words): [0..2] min xyz (ordered), [3..5] max xyz (ordered),
// then doubles at byte 32: centre xyz, scale, sum xyz (chunked), maxnorm bits
struct Reduce {
  unsigned int lo[3], hi[3];
  unsigned int pad[2];
  double centre[3];
  double scale;
  double sum[3];
  unsigned long long maxnorm_bits;
  float ubox[6];  // bbox of the unit surface (float32 points)
};

__global__ void reduce_init_kernel(Reduce* r) {
  if (threadIdx.x == 0) {
    for (int k = 0; k < 3; k++) {
      r->lo[k] = 0xffffffffu;
      r->hi[k] = 0u;
      r->sum[k] = 0.0;
    }
    r->maxnorm_bits = 0ull;
  }
}

__global__ void bbox_kernel(const float* __restrict__ v, int64_t V, Reduce* r) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V;
       i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float p = v[3 * i + k];
      lo[k] = fminf(lo[k], p);
      hi[k] = fmaxf(hi[k], p);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(&r->lo[k], f2ord(lo[k]));
      atomicMax(&r->hi[k], f2ord(hi[k]));
    }
  }
}

// centre of mass: fixed order = 1024-point chunks summed left to right, then
// the chunk sums left to right (the oracle uses the same order)
#define FB_COM_CHUNK 1024
__global__ void com_chunk_kernel(const float* __restrict__ v, int64_t V, double* chunk_sums) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t b = c * FB_COM_CHUNK;
  if (b >= V) return;
  int64_t e = b + FB_COM_CHUNK < V ? b + FB_COM_CHUNK : V;
  double s[3] = {0, 0, 0};
  for (int64_t i = b; i < e; i++)
    for (int k = 0; k < 3; k++) s[k] = FB_ADD(s[k], (double)v[3 * i + k]);
  for (int k = 0; k < 3; k++) chunk_sums[3 * c + k] = s[k];
}

__global__ void centre_kernel(Reduce* r, const double* chunk_sums, int64_t n_chunks, int64_t V,
                              flyby_norm_opts o) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double lo[3], hi[3];
  for (int k = 0; k < 3; k++) {
    lo[k] = (double)ord2f(r->lo[k]);
    hi[k] = (double)ord2f(r->hi[k]);
  }
  if (o.centre_mode == 1 && !o.has_translate) {
    double s[3] = {0, 0, 0};
    for (int64_t c = 0; c < n_chunks; c++)
      for (int k = 0; k < 3; k++) s[k] = FB_ADD(s[k], chunk_sums[3 * c + k]);
    for (int k = 0; k < 3; k++) r->centre[k] = FB_DIV(s[k], (double)V);
  } else {
    for (int k = 0; k < 3; k++)
      r->centre[k] = o.has_translate ? o.translate[k] : FB_DIV(FB_ADD(lo[k], hi[k]), 2.0);
  }
  if (o.has_scale) r->scale = o.scale_factor;
  else if (o.scale_mode == 0) {
    double d[3] = {FB_SUB(hi[0], r->centre[0]), FB_SUB(hi[1], r->centre[1]), FB_SUB(hi[2], r->centre[2])};
    r->scale = FB_DIV(1.0, norm3(d));
  }
}

__global__ void maxnorm_kernel(const float* __restrict__ v, int64_t V, Reduce* r) {
  double m = 0.0;
  double c[3] = {r->centre[0], r->centre[1], r->centre[2]};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V;
       i += (int64_t)gridDim.x * blockDim.x) {
    double d[3] = {FB_SUB((double)v[3 * i], c[0]), FB_SUB((double)v[3 * i + 1], c[1]),
                   FB_SUB((double)v[3 * i + 2], c[2])};
    double n = norm3(d);
    m = n > m ? n : m;
  }
  // non-negative doubles order like their bit patterns
  atomicMax(&r->maxnorm_bits, (unsigned long long)__double_as_longlong(m));
}

__global__ void scale_from_maxnorm_kernel(Reduce* r) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    r->scale = FB_DIV(1.0, __longlong_as_double((long long)r->maxnorm_bits));
}

__global__ void transform_kernel(const float* __restrict__ v, int64_t n3, const Reduce* r,
                                 float* __restrict__ out, int skip) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n3) return;
  if (skip) { out[i] = v[i]; return; }
  int k = (int)(i % 3);
  out[i] = (float)FB_MUL(FB_SUB((double)v[i], r->centre[k]), r->scale);
}

// bbox of the unit surface = image of the raw bbox (x -> float((x-c)*s) is monotone)
__global__ void unit_box_kernel(Reduce* r, int skip) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int k = 0; k < 3; k++) {
    float a = ord2f(r->lo[k]), b = ord2f(r->hi[k]);
    if (!skip) {
      a = (float)FB_MUL(FB_SUB((double)a, r->centre[k]), r->scale);
      b = (float)FB_MUL(FB_SUB((double)b, r->centre[k]), r->scale);
    } else {
      r->centre[k] = 0.0;
    }
    r->ubox[k] = fminf(a, b);
    r->ubox[3 + k] = fmaxf(a, b);
  }
  if (skip) r->scale = 1.0;
}

// ---------------------------------------------------------------- normals
__global__ void face_normal_kernel(const float* __restrict__ v, const int32_t* __restrict__ tris,
                                   int64_t T, float* __restrict__ fn, uint32_t* adj_count) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
  int i0 = tris[3 * c], i1 = tris[3 * c + 1], i2 = tris[3 * c + 2];
  double a[3], b[3], n[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double p0 = v[3 * (int64_t)i0 + k], p1 = v[3 * (int64_t)i1 + k], p2 = v[3 * (int64_t)i2 + k];
    a[k] = FB_SUB(p2, p1);
    b[k] = FB_SUB(p0, p1);
  }
  cross3(a, b, n);
  double len = norm3(n);
#pragma unroll
  for (int k = 0; k < 3; k++) fn[3 * c + k] = (len != 0.0) ? (float)FB_DIV(n[k], len) : 0.0f;
  atomicAdd(&adj_count[i0], 1u);
  atomicAdd(&adj_count[i1], 1u);
  atomicAdd(&adj_count[i2], 1u);
}

__global__ void adj_fill_kernel(const int32_t* __restrict__ tris, int64_t T,
                                const uint32_t* __restrict__ adj_start, uint32_t* adj_cur,
                                int32_t* __restrict__ adj) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    int v = tris[3 * c + j];
    uint32_t s = atomicAdd(&adj_cur[v], 1u);
    adj[adj_start[v] + s] = (int32_t)c;
  }
}

__device__ void sort_small(int32_t* a, int n) {
  if (n <= 32) {
    for (int i = 1; i < n; i++) {
      int32_t x = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; j--; }
      a[j + 1] = x;
    }
    return;
  }
  // heap sort for the rare high-valence vertex
  for (int start = n / 2 - 1; start >= 0; start--) {
    int root = start;
    for (;;) {
      int ch = 2 * root + 1;
      if (ch >= n) break;
      if (ch + 1 < n && a[ch] < a[ch + 1]) ch++;
      if (a[root] >= a[ch]) break;
      int32_t t = a[root]; a[root] = a[ch]; a[ch] = t;
      root = ch;
    }
  }
  for (int end = n - 1; end > 0; end--) {
    int32_t t = a[0]; a[0] = a[end]; a[end] = t;
    int root = 0;
    for (;;) {
      int ch = 2 * root + 1;
      if (ch >= end) break;
      if (ch + 1 < end && a[ch] < a[ch + 1]) ch++;
      if (a[root] >= a[ch]) break;
      int32_t u = a[root]; a[root] = a[ch]; a[ch] = u;
      root = ch;
    }
  }
}

__global__ void point_normal_kernel(int64_t V, const uint32_t* __restrict__ adj_start,
                                    int32_t* __restrict__ adj, const float* __restrict__ fn,
                                    float4* __restrict__ normals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= V) return;
  uint32_t b = adj_start[i], e = adj_start[i + 1];
  sort_small(adj + b, (int)(e - b));
  float s[3] = {0.f, 0.f, 0.f};
  for (uint32_t q = b; q < e; q++) {
    int64_t c = adj[q];
#pragma unroll
    for (int k = 0; k < 3; k++) s[k] = FB_FADD(s[k], fn[3 * c + k]);
  }
  float ss = FB_FADD(FB_FADD(FB_FMUL(s[0], s[0]), FB_FMUL(s[1], s[1])), FB_FMUL(s[2], s[2]));
  double len = FB_SQRT((double)ss);
  float4 o = make_float4(s[0], s[1], s[2], 0.f);
  if (len != 0.0) {
    o.x = (float)FB_DIV((double)s[0], len);
    o.y = (float)FB_DIV((double)s[1], len);
    o.z = (float)FB_DIV((double)s[2], len);
  }
  normals[i] = o;
}

// ---------------------------------------------------------------- driver
int prep_run(flyby_ctx* c) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int64_t V = m.V, T = m.T;
  const float* raw = m.raw_verts.as<float>();
  FB_CUDA(c, m.reduce.reserve(sizeof(Reduce)));
  FB_CUDA(c, m.verts.reserve(sizeof(float) * 3 * (size_t)V));
  FB_CUDA(c, m.normals.reserve(sizeof(float4) * (size_t)V));
  FB_CUDA(c, m.face_n.reserve(sizeof(float) * 3 * (size_t)(T > 0 ? T : 1)));
  FB_CUDA(c, m.adj_start.reserve(sizeof(uint32_t) * (size_t)(V + 1)));
  FB_CUDA(c, m.adj_cur.reserve(sizeof(uint32_t) * (size_t)(V + 1)));
  FB_CUDA(c, m.adj.reserve(sizeof(int32_t) * 3 * (size_t)(T > 0 ? T : 1)));
  Reduce* red = m.reduce.as<Reduce>();
  const int skip = c->norm.skip;

  FB_STAT_EVENT(c, c->ev[0], st);
  reduce_init_kernel<<<1, 32, 0, st>>>(red);
  FB_LAUNCH_CHECK(c);
  int blocks = (int)(cdiv(V, 256) < 1184 ? cdiv(V, 256) : 1184);
  if (blocks < 1) blocks = 1;
  bbox_kernel<<<blocks, 256, 0, st>>>(raw, V, red);
  FB_LAUNCH_CHECK(c);
  int64_t n_chunks = 0;
  double* chunk_sums = nullptr;
  if (!skip && c->norm.centre_mode == 1 && !c->norm.has_translate) {
    n_chunks = cdiv(V, FB_COM_CHUNK);
    FB_CUDA(c, c->stage_misc.reserve(sizeof(double) * 3 * (size_t)n_chunks));
    chunk_sums = c->stage_misc.as<double>();
    com_chunk_kernel<<<(unsigned)cdiv(n_chunks, 128), 128, 0, st>>>(raw, V, chunk_sums);
    FB_LAUNCH_CHECK(c);
  }
  if (!skip) {
    centre_kernel<<<1, 32, 0, st>>>(red, chunk_sums, n_chunks, V, c->norm);
    FB_LAUNCH_CHECK(c);
    if (!c->norm.has_scale && c->norm.scale_mode == 1) {
      maxnorm_kernel<<<blocks, 256, 0, st>>>(raw, V, red);
      FB_LAUNCH_CHECK(c);
      scale_from_maxnorm_kernel<<<1, 32, 0, st>>>(red);
      FB_LAUNCH_CHECK(c);
    }
  }
  transform_kernel<<<(unsigned)cdiv(3 * V, 256), 256, 0, st>>>(raw, 3 * V, red, m.verts.as<float>(), skip);
  FB_LAUNCH_CHECK(c);
  unit_box_kernel<<<1, 32, 0, st>>>(red, skip);
  FB_LAUNCH_CHECK(c);
  FB_STAT_EVENT(c, c->ev[1], st);

  // point normals: not needed before the shading, so they go to the side stream (see ctx.h)
  cudaStream_t ss = c->side_stream ? c->side_stream : st;
  if (ss != st) {
    FB_CUDA(c, cudaEventRecord(c->ev_unit, st));  // the unit surface is complete on the main stream ...
    FB_SIDE_WAIT(c, c->ev_unit);  // ... before the side stream reads it
  }
  FB_CUDA(c, cudaMemsetAsync(m.adj_start.p, 0, sizeof(uint32_t) * (size_t)(V + 1), ss));
  FB_CUDA(c, cudaMemsetAsync(m.adj_cur.p, 0, sizeof(uint32_t) * (size_t)(V + 1), ss));
  if (T > 0) {
    face_normal_kernel<<<(unsigned)cdiv(T, 256), 256, 0, ss>>>(
        m.verts.as<float>(), m.tris.as<int32_t>(), T, m.face_n.as<float>(), m.adj_start.as<uint32_t>());
    FB_LAUNCH_CHECK(c);
  }
  int rc = exclusive_scan_u32(c, m.adj_start.as<uint32_t>(), V + 1, ss != st ? m.tree_scan_tmp : m.scan_tmp, ss);
  if (rc) return rc;
  if (T > 0) {
    adj_fill_kernel<<<(unsigned)cdiv(T, 256), 256, 0, ss>>>(
        m.tris.as<int32_t>(), T, m.adj_start.as<uint32_t>(), m.adj_cur.as<uint32_t>(), m.adj.as<int32_t>());
    FB_LAUNCH_CHECK(c);
  }
  point_normal_kernel<<<(unsigned)cdiv(V, 128), 128, 0, ss>>>(
      V, m.adj_start.as<uint32_t>(), m.adj.as<int32_t>(), m.face_n.as<float>(), m.normals.as<float4>());
  FB_LAUNCH_CHECK(c);
  FB_STAT_EVENT(c, c->ev[2], ss);
  return FLYBY_OK;
}

// exported for lbvh.cu / api.cu
const float* prep_unit_box(flyby_ctx* c) { return c->mesh.reduce.as<Reduce>()->ubox; }
const double* prep_centre_scale(flyby_ctx* c) { return c->mesh.reduce.as<Reduce>()->centre; }

}  // namespace fb
