// raster.cu -- triangle-centric candidate search for the fly-by views (default render path).
//
// The rays of a view are parallel and sit on a regular pixel grid, and a triangle of a
// 100k-2M triangle surface covers only a few pixels.  Instead of walking the LBVH once per
// ray packet, every triangle is projected into the view frame once and tested -- with the
// same float32 screening predicate the trace kernel uses (flyby_core.cuh) -- against the
// handful of pixels inside its bounding box:
//   pass A  uz[pixel]  = min over SURE triangles of their far depth bound   (atomicMin)
//   pass B  cand[pixel] += every triangle that is not a certain miss and whose near depth
//           bound is not behind uz[pixel]                                   (atomicAdd slot)
// resolve_kernel (screen.cu) then decides every pixel exactly, in double precision, from its
// 1-2 candidates; a pixel with more than 8 candidates is re-traced exactly through the LBVH.
// The candidate set is a superset of what the exact test can accept in front of the nearest
// certain hit, so the result equals the LBVH traversal and the oracle bit for bit; the order
// in which atomics land does not matter (the exact tie-break is (t, cell id)).
//
// Work: T x views threads, ~45 float instructions per covered pixel; bytes: 48 B per triangle
// per view (L2 resident), 4-8 B of L2 atomics per covered pixel.
//
// Reference replaced: the per-view / per-plane-point loops of
// src/app/fly_by_features.cxx:652-852 (this is the search part of IntersectWithLine).
#include "trace_common.cuh"

namespace fb {

#define RT_THREADS 256
#define RT_BIG 48  // bounding boxes with more pixels than this are rasterised by a whole warp

struct RasterParams {
  RenderParams r;
  unsigned int* uz;  // [n_pix] ordered-float bits of the nearest certain far depth
  int* cnt;          // [n_pix] candidates found
  int* cand;         // [n_pix, 4] candidate slots 0..3
  int* cand2;        // [n_pix, 4] candidate slots 4..7 (touched only by pixels that need them)
  int T;
};

__device__ __forceinline__ unsigned int ford(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int PASS>
__device__ __forceinline__ void visit_pixel(const RasterParams& p, const ViewF& V, const TriS& t, int slot,
                                            int64_t view_base, int res, int i, int j) {
  const float qx = fmaf((float)i, V.qdx, V.q0x), qy = fmaf((float)j, V.qdy, V.q0y);
  const int cls = screen_pixel(V, t, qx, qy);
  if (cls == SC_MISS) return;
  const int64_t pix = view_base + (int64_t)j * res + i;
  if (PASS == 0) {
    if (cls == SC_SURE) atomicMin(p.uz + pix, ford(t.zhi));
  } else {
    if (ford(t.zlo) <= __ldcg(p.uz + pix)) {
      const int k = atomicAdd(p.cnt + pix, 1);
      if (k < 4) p.cand[pix * 4 + k] = slot;
      else if (k < 8) p.cand2[pix * 4 + (k - 4)] = slot;
    }
  }
}

template <int PASS>
__global__ void __launch_bounds__(RT_THREADS) raster_kernel(const RasterParams p) {
  __shared__ ViewF s_vf;
  const int view = blockIdx.y;  // local view index of this render call
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.r.viewsf + p.r.view_begin + view);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_vf);
    if (threadIdx.x < (int)(sizeof(ViewF) / 4)) dst[threadIdx.x] = __ldg(src + threadIdx.x);
  }
  __syncthreads();
  const ViewF& V = s_vf;
  const int res = p.r.res;
  const int64_t view_base = (int64_t)view * res * res;
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * RT_THREADS + threadIdx.x;

  TriS t;
  PixRange r;
  int npx = 0;
  if (slot < p.T) {
    const float4 a = __ldg(p.r.tv0 + slot), b = __ldg(p.r.tv1 + slot), e = __ldg(p.r.tv2 + slot);
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
    screen_setup(V, 0.0f, P0, P1, P2, &t);  // ray origins lie in the view plane: depth 0 (+- e_z)
    if (t.flags & 1) {
      r = pixel_range(V, t, res);
      if (r.i1 >= r.i0 && r.j1 >= r.j0) npx = (r.i1 - r.i0 + 1) * (r.j1 - r.j0 + 1);
    }
  }
  // small triangles: one thread walks the pixels of its bounding box
  if (npx > 0 && npx <= RT_BIG) {
    for (int j = r.j0; j <= r.j1; j++)
      for (int i = r.i0; i <= r.i1; i++) visit_pixel<PASS>(p, V, t, slot, view_base, res, i, j);
  }
  // large triangles (coarse meshes, close-ups): the warp takes them one at a time, 32 pixels per step
  unsigned big = __ballot_sync(0xffffffffu, npx > RT_BIG);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    TriS u;
    u.ax = __shfl_sync(0xffffffffu, t.ax, src); u.ay = __shfl_sync(0xffffffffu, t.ay, src);
    u.bx = __shfl_sync(0xffffffffu, t.bx, src); u.by = __shfl_sync(0xffffffffu, t.by, src);
    u.cx = __shfl_sync(0xffffffffu, t.cx, src); u.cy = __shfl_sync(0xffffffffu, t.cy, src);
    u.lox = __shfl_sync(0xffffffffu, t.lox, src); u.hix = __shfl_sync(0xffffffffu, t.hix, src);
    u.loy = __shfl_sync(0xffffffffu, t.loy, src); u.hiy = __shfl_sync(0xffffffffu, t.hiy, src);
    u.zlo = __shfl_sync(0xffffffffu, t.zlo, src); u.zhi = __shfl_sync(0xffffffffu, t.zhi, src);
    u.area2 = __shfl_sync(0xffffffffu, t.area2, src);
    u.flags = __shfl_sync(0xffffffffu, t.flags, src);
    const int i0 = __shfl_sync(0xffffffffu, r.i0, src), i1 = __shfl_sync(0xffffffffu, r.i1, src);
    const int j0 = __shfl_sync(0xffffffffu, r.j0, src), j1 = __shfl_sync(0xffffffffu, r.j1, src);
    const int uslot = __shfl_sync(0xffffffffu, slot, src);
    const int w = i1 - i0 + 1;
    const long long n = (long long)w * (j1 - j0 + 1);
    for (long long k = lane; k < n; k += 32)
      visit_pixel<PASS>(p, V, u, uslot, view_base, res, i0 + (int)(k % w), j0 + (int)(k / w));
  }
}

__global__ void raster_init_kernel(unsigned int* uz, int* cnt, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    uz[i] = 0xffffffffu;
    cnt[i] = 0;
  }
}

int resolve_launch(flyby_ctx* c, const RenderParams& rp, const int4* cand, const int* cnt, const int4* cand2);  // screen.cu

int raster_render(flyby_ctx* c, const RenderParams& rp) {
  cudaStream_t st = c->stream;
  const int64_t n_pix = (int64_t)rp.n_views * rp.res * rp.res;
  FB_CUDA(c, c->cand.reserve(sizeof(int4) * (size_t)n_pix));
  FB_CUDA(c, c->cand2.reserve(sizeof(int4) * (size_t)n_pix));
  FB_CUDA(c, c->raster_tmp.reserve(8 * (size_t)n_pix));
  RasterParams p;
  p.r = rp;
  p.uz = c->raster_tmp.as<unsigned int>();
  p.cnt = reinterpret_cast<int*>(p.uz + n_pix);
  p.cand = c->cand.as<int>();
  p.cand2 = c->cand2.as<int>();
  p.T = (int)c->mesh.T;
  raster_init_kernel<<<(unsigned)cdiv(n_pix, 256), 256, 0, st>>>(p.uz, p.cnt, n_pix);
  FB_LAUNCH_CHECK(c);
  dim3 grid((unsigned)cdiv(p.T, RT_THREADS), (unsigned)rp.n_views);
  raster_kernel<0><<<grid, RT_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  raster_kernel<1><<<grid, RT_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  FB_CUDA(c, cudaEventRecord(c->ev[7], st));  // between the candidate search and the exact resolve
  return resolve_launch(c, rp, c->cand.as<int4>(), p.cnt, c->cand2.as<int4>());
}

}  // namespace fb
