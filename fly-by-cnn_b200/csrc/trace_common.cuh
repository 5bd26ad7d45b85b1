// trace_common.cuh -- pieces shared by the render kernels (raster.cu: scan conversion + exact resolve,
// screen.cu: float32 LBVH trace kernel + exact resolve, trace.cu: fused exact LBVH kernel, pinhole views).
#pragma once
#include "ctx.h"

namespace fb {

const float* prep_unit_box(flyby_ctx* c);

// ---------------------------------------------------------------- node fetch
struct GlobalFetch {
  const Node* gl;
  __device__ __forceinline__ Node node(int i) const {
    const float4* p = reinterpret_cast<const float4*>(gl + i);
    Node n;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    n.box[0] = a.x; n.box[1] = a.y; n.box[2] = a.z; n.box[3] = a.w;
    n.box[4] = b.x; n.box[5] = b.y; n.box[6] = b.z; n.box[7] = b.w;
    n.box[8] = c.x; n.box[9] = c.y; n.box[10] = c.z; n.box[11] = c.w;
    n.c0 = __float_as_int(d.x); n.c1 = __float_as_int(d.y);
    n.parent = __float_as_int(d.z); n.sibling = __float_as_int(d.w);
    return n;
  }
  __device__ __forceinline__ int parent(int i) const { return __ldg(&gl[i].parent); }
  __device__ __forceinline__ int sibling(int i) const { return __ldg(&gl[i].sibling); }
};

// top-of-tree in shared memory, the rest through the read-only path
struct TopFetch {
  const Node* gl;
  const Node* sm;
  int ntop;
  __device__ __forceinline__ Node node(int i) const {
    Node n;
    float4 a, b, c, d;
    if (i < ntop) {
      const float4* p = reinterpret_cast<const float4*>(sm + i);
      a = p[0]; b = p[1]; c = p[2]; d = p[3];
    } else {
      const float4* p = reinterpret_cast<const float4*>(gl + i);
      a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3);
    }
    n.box[0] = a.x; n.box[1] = a.y; n.box[2] = a.z; n.box[3] = a.w;
    n.box[4] = b.x; n.box[5] = b.y; n.box[6] = b.z; n.box[7] = b.w;
    n.box[8] = c.x; n.box[9] = c.y; n.box[10] = c.z; n.box[11] = c.w;
    n.c0 = __float_as_int(d.x); n.c1 = __float_as_int(d.y);
    n.parent = __float_as_int(d.z); n.sibling = __float_as_int(d.w);
    return n;
  }
  __device__ __forceinline__ int parent(int i) const {
    return i < ntop ? sm[i].parent : __ldg(&gl[i].parent);
  }
  __device__ __forceinline__ int sibling(int i) const {
    return i < ntop ? sm[i].sibling : __ldg(&gl[i].sibling);
  }
};

struct WorkCount {
  unsigned int nodes = 0, tris = 0;
  __device__ __forceinline__ void node() { nodes++; }
  __device__ __forceinline__ void tri() { tris++; }
};

// ---------------------------------------------------------------- division by a launch-invariant divisor
// n / d for 32-bit n with a multiply-high and two shifts (Granlund & Montgomery): pixel_ray splits a pixel id into
// (view, row, column) once per pixel, which as two hardware-less 32-bit divisions cost ~40 instructions
struct FastDiv {
  uint32_t mul, sh1, sh2;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{0u, 0u, 0u};
  if (d <= 1) return f;  // q = n
  uint32_t l = 0;
  while ((1ull << l) < d) l++;
  f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1ull);
  f.sh1 = 1;
  f.sh2 = l - 1;
  return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.mul, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

// ---------------------------------------------------------------- render
struct RenderParams {
  const Node* nodes;
  const int32_t* n_top_dev;
  const float4 *tv0, *tv1, *tv2;
  const int4* tvid;
  const float4* trec;   // 64-byte triangle records (resolve kernels)
  const double* tc;     // [res] i / (res - 1), exactly as view_ray computes it
  const float4* normals;
  const View* views;
  const float* ubox;
  int view_begin, n_views, res;
  int use_z, split_z, enc, channels;
  double zscale, spacing;
  double tol2;  // tol^2 of IntersectWithLine
  float* feat;
  int32_t* ids;
  double* tt;
  unsigned long long* counters;  // [0] tile counter, [1] nodes, [2] tris, [3] hits
  int tiles_x, tiles_y;          // padded tile grid (x even, y multiple of 4)
  long long n_tiles;
  int top_cap;                   // nodes that fit in the dynamic shared memory
  const ViewF* viewsf;           // float32 screening frame of every view
  // per-point scalars behind the feature channels (flyby_set_point_scalars): pixel stride = channels + n_scal
  const float* scal;             // [V, n_scal] or nullptr
  const uint8_t* flip;           // [T] cells reversed by the consistency pass (their vertex 0 is p2), or nullptr
  int n_scal, scal_mode, stride;
  float scal_zero;
  int approx;                    // FLYBY_SHADE_TOLERANCE: shade_approx for pixels whose only candidate is a certain hit
  int share;                     // FLYBY_SHARE_GPU: the resolve kernels leave room for another context's candidate search
  FastDiv div_view, div_res;     // by res * res and by res
};

#ifndef RK_MINB
#define RK_MINB 2
#endif
#define RK_THREADS 256
#define RK_WARPS (RK_THREADS / 32)
#define RK_QCAP 96

__device__ __forceinline__ void write_pixel(const RenderParams& p, int64_t pix, const float* px, int cell,
                                            double t) {
  if (p.feat) {
    if (p.channels == 4 && p.n_scal == 0) {  // (3 image channels + 1 scalar also make a stride of 4: not this path)
      __stcs(reinterpret_cast<float4*>(p.feat) + pix, make_float4(px[0], px[1], px[2], px[3]));
    } else {
      float* o = p.feat + pix * p.stride;
      __stcs(o, px[0]);
      __stcs(o + 1, px[1]);
      __stcs(o + 2, px[2]);
      if (p.channels == 4) __stcs(o + 3, px[3]);
      if (cell < 0)  // background: the scalar channels get `zero` (--zero); hits write theirs in write_hit_scalars
        for (int k = 0; k < p.n_scal; k++) __stcs(o + p.channels + k, p.scal_zero);
    }
  }
  if (p.ids) __stcs(p.ids + pix, cell);
  if (p.tt) __stcs(p.tt + pix, t);
}

// Scalar channels of a hit pixel (src/py/fly_by_features.py:385-392 concatenates them behind the image channels).
// mode 0: the value at the cell's vertex 0 -- what the reference reads back through the point-id colour map
// (GetPointIdColors utils.py:604-621 colours a cell with its first point; ExtractPointFeaturesClass :644-662);
// mode 1: sum_k w_k f_k with the exact test's weights, products and sums in double, left to right
// (fly_by_features.cxx:766-793), the expression of interpolate_kernel / orc_interpolate_features.
__device__ __forceinline__ void write_hit_scalars(const RenderParams& p, int64_t pix, const double* w, int pid0, int pid1,
                                                  int pid2, int cell) {
  if (!p.n_scal || !p.feat) return;
  float* o = p.feat + pix * p.stride + p.channels;
  const int S = p.n_scal;
  if (p.scal_mode == 0) {
    const int v0 = (p.flip && p.flip[cell]) ? pid2 : pid0;  // vertex 0 of the cell as the caller numbered it
    for (int k = 0; k < S; k++) __stcs(o + k, __ldg(p.scal + (int64_t)v0 * S + k));
  } else {
    for (int k = 0; k < S; k++) {
      const double a = FB_MUL(w[0], (double)__ldg(p.scal + (int64_t)pid0 * S + k));
      const double b = FB_MUL(w[1], (double)__ldg(p.scal + (int64_t)pid1 * S + k));
      const double e = FB_MUL(w[2], (double)__ldg(p.scal + (int64_t)pid2 * S + k));
      __stcs(o + k, (float)FB_ADD(FB_ADD(a, b), e));
    }
  }
}

// conservative float32 test of a ray against the mesh bounds (same maths as the
// node test, so a ray rejected here would also be rejected at the root)
__device__ __forceinline__ bool hits_bounds(const float* ubox, const double* o, const View& vw) {
  float of[3] = {(float)o[0], (float)o[1], (float)o[2]};
  float tn;
  return slab(ubox, of, vw.invf, vw.slack, 1.0f + vw.tslack, &tn);
}


// One bulk async copy (TMA engine, SASS UBLKCP) of the breadth-first top of the tree into shared
// memory, completion through an mbarrier.  Must be called by every thread of the CTA.
__device__ __forceinline__ void stage_top_of_tree(Node* top, const Node* nodes, int ntop, unsigned long long* bar_mem) {
  if (ntop > 0) {
    const uint32_t bytes = (uint32_t)ntop * (uint32_t)sizeof(Node);
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_mem);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(top);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
          "l"(nodes), "r"(bytes), "r"(bar)
          : "memory");
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(0u)
          : "memory");
    }
  } else {
    __syncthreads();
  }
}

// ---------------------------------------------------------------- exact resolve helpers
struct PixelRay {
  double o[3], d[3];
  double dlen;  // |d|
};
__device__ __forceinline__ PixelRay pixel_ray(const RenderParams& p, int64_t pix_id) {
  // a render call covers fewer than 2^32 pixels (checked by the host side): 32-bit index arithmetic
  const uint32_t per_view = (uint32_t)p.res * (uint32_t)p.res;
  const uint32_t v = fastdiv((uint32_t)pix_id, p.div_view);
  const uint32_t pix = (uint32_t)pix_id - v * per_view;
  const View& vw = p.views[p.view_begin + (int)v];
  PixelRay r;
  // view_ray() with the two divisions i / (res - 1) taken from the table (bit-identical values)
  const uint32_t pj = fastdiv(pix, p.div_res), pi = pix - pj * (uint32_t)p.res;
  const double tc0 = p.tc[pi], tc1 = p.tc[pj];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float pf = (float)FB_ADD(FB_ADD(vw.origin[k], FB_MUL(tc0, vw.v1[k])), FB_MUL(tc1, vw.v2[k]));
    r.o[k] = FB_ADD(FB_MUL((double)pf, p.spacing), vw.delta[k]);
  }
  r.d[0] = vw.dir[0]; r.d[1] = vw.dir[1]; r.d[2] = vw.dir[2];
  r.dlen = vw.dlen;
  return r;
}

__device__ __forceinline__ void shade_and_write(const RenderParams& p, int64_t pix_id, const PixelRay& r, double t,
                                                const double* w, int pid0, int pid1, int pid2, int cell) {
  float px[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.feat) {
    const double x[3] = {FB_ADD(r.o[0], FB_MUL(t, r.d[0])), FB_ADD(r.o[1], FB_MUL(t, r.d[1])), FB_ADD(r.o[2], FB_MUL(t, r.d[2]))};
    const float4 n0 = __ldg(p.normals + pid0), n1 = __ldg(p.normals + pid1), n2 = __ldg(p.normals + pid2);
    const float N0[3] = {n0.x, n0.y, n0.z}, N1[3] = {n1.x, n1.y, n1.z}, N2[3] = {n2.x, n2.y, n2.z};
    shade(r.o, x, w, N0, N1, N2, p.use_z, p.split_z, p.enc, p.zscale, px);
  }
  write_pixel(p, pix_id, px, cell, t);
  write_hit_scalars(p, pix_id, w, pid0, pid1, pid2, cell);
}

// Rare path, kept out of line so that it does not inflate the registers of resolve_kernel:
// a ray with more candidates than the list holds (layered / degenerate geometry) is re-traced
// exactly, on its own, with the stackless per-ray traversal.
struct ExactHit {
  double t, w0, w1, w2;
  int slot, cell;
};
static __device__ __noinline__ ExactHit resolve_overflow(const RenderParams& p, int view, double o0, double o1, double o2,
                                                         const ViewF* vf = nullptr, float qx = 0.f, float qy = 0.f) {
  const View& vw = p.views[view];
  const double o[3] = {o0, o1, o2};
  const double d[3] = {vw.dir[0], vw.dir[1], vw.dir[2]};
  GlobalFetch f{p.nodes};
  NoCount nc;
  float invf[3] = {vw.invf[0], vw.invf[1], vw.invf[2]};
  const Hit h = traverse(f, p.tv0, p.tv1, p.tv2, o, d, vw.dirf, invf, vw.slack, vw.tslack, p.tol2, nc, vf, qx, qy, 0.0f);
  ExactHit r;
  r.t = DBL_MAX; r.w0 = r.w1 = r.w2 = 0.0; r.slot = -1; r.cell = 0x7fffffff;
  if (h.slot >= 0) {
    const float4 a = p.tv0[h.slot], b = p.tv1[h.slot], e = p.tv2[h.slot];
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
    double x[3], w[3];
    tri_exact(o, d, P0, P1, P2, DBL_MAX, p.tol2, &r.t, x, w);
    r.w0 = w[0]; r.w1 = w[1]; r.w2 = w[2];
    r.slot = h.slot;
    r.cell = h.cell;
  }
  return r;
}


}  // namespace fb
