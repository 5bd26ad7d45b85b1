// flyby_core.cuh -- per-element device logic of the fly-by path.
//
// Everything here is __host__ __device__ so that tests/hostsim can compile the
// very same traversal / intersection / tree-construction code with g++ and
// debug it without a GPU.  The product only ever runs it inside the CUDA
// kernels of this directory; there is no CPU execution path in the library.
//
// Arithmetic contract (DESIGN.md "Exactness"):
//  * everything that decides WHICH cell a ray hits, and the values written to
//    the output tensor, is IEEE double with explicitly rounded operations
//    (no FMA contraction) in a fixed evaluation order;
//  * the BVH traversal around it is float32 and CONSERVATIVE: it may visit a
//    box the exact ray misses, it never skips a box the exact hit lies in.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };  // host simulator build only
#endif

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#define FB_COLD static __host__ __device__ __noinline__
#else
#define FB_HD inline
#define FB_COLD static __attribute__((noinline))
#endif

#if defined(__CUDA_ARCH__)
#define FB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define FB_FRCP(a) __frcp_rn((a))
#define FB_MUL(a, b) __dmul_rn((a), (b))
#define FB_ADD(a, b) __dadd_rn((a), (b))
#define FB_SUB(a, b) __dsub_rn((a), (b))
#define FB_DIV(a, b) __ddiv_rn((a), (b))
#define FB_SQRT(a) __dsqrt_rn((a))
#define FB_FADD(a, b) __fadd_rn((a), (b))
#define FB_FMUL(a, b) __fmul_rn((a), (b))
#define FB_F2F_RU(a) __double2float_ru((a))
#else  // host build: compile with -ffp-contract=off
#define FB_FMA(a, b, c) fma((a), (b), (c))
#define FB_FRCP(a) (1.0f / (a))
#define FB_MUL(a, b) ((a) * (b))
#define FB_ADD(a, b) ((a) + (b))
#define FB_SUB(a, b) ((a) - (b))
#define FB_DIV(a, b) ((a) / (b))
#define FB_SQRT(a) sqrt((a))
#define FB_FADD(a, b) ((float)(a) + (float)(b))
#define FB_FMUL(a, b) ((float)(a) * (float)(b))
static inline float fb_f2f_ru(double a) {
  float f = (float)a;
  return ((double)f < a) ? nextafterf(f, INFINITY) : f;
}
#define FB_F2F_RU(a) fb_f2f_ru((a))
#endif

namespace fb {

// ---------------------------------------------------------------- vectors
FB_HD double dot3(const double* a, const double* b) {
  return FB_ADD(FB_ADD(FB_MUL(a[0], b[0]), FB_MUL(a[1], b[1])), FB_MUL(a[2], b[2]));
}
FB_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = FB_SUB(FB_MUL(a[1], b[2]), FB_MUL(a[2], b[1]));
  c[1] = FB_SUB(FB_MUL(a[2], b[0]), FB_MUL(a[0], b[2]));
  c[2] = FB_SUB(FB_MUL(a[0], b[1]), FB_MUL(a[1], b[0]));
}
FB_HD double norm3(const double* a) { return FB_SQRT(dot3(a, a)); }

// ---------------------------------------------------------------- views
// One record per viewpoint; rebuilt by view_setup_kernel for every render
// call because it depends on plane_scale / plane_spacing.
// Reference: src/app/fly_by_features.cxx:652-702 (basis and vtkPlaneSource
// corner points) and :742-748 (segment from each plane point).
struct View {
  double sp[3], n[3], x[3], y[3];
  double origin[3], v1[3], v2[3];  // plane Origin, Point1-Origin, Point2-Origin
  double delta[3];                 // sp - sp*spacing
  double dir[3];                   // end - origin = -(n*R)*2
  float dirf[3], invf[3];          // float32 direction and its reciprocal (traversal only)
  float slack;                     // box inflation covering float32 ray rounding
  float tslack;                    // slack in units of the segment parameter
  double dlen;                     // |dir| (shade_approx: depth = t |dir|)
};

FB_HD void view_setup(const float* spf, double radius, double plane_scale, double spacing, View* v, double tol = 0.0) {
  for (int k = 0; k < 3; k++) v->sp[k] = (double)spf[k];
  double len = norm3(v->sp);
  for (int k = 0; k < 3; k++) v->n[k] = FB_DIV(v->sp[k], len);
  bool north = fabs(v->n[0]) <= 1e-8 && fabs(v->n[1]) <= 1e-8 && fabs(FB_SUB(v->n[2], 1.0)) <= 1e-8;
  bool south = fabs(v->n[0]) <= 1e-8 && fabs(v->n[1]) <= 1e-8 && fabs(FB_ADD(v->n[2], 1.0)) <= 1e-8;
  if (north) {
    v->x[0] = 1; v->x[1] = 0; v->x[2] = 0; v->y[0] = 0; v->y[1] = 1; v->y[2] = 0;
  } else if (south) {
    v->x[0] = -1; v->x[1] = 0; v->x[2] = 0; v->y[0] = 0; v->y[1] = -1; v->y[2] = 0;
  } else {
    double up[3] = {0.0, 0.0, 1.0}, c[3];
    cross3(v->n, up, c);
    double l = norm3(c);
    for (int k = 0; k < 3; k++) v->x[k] = FB_DIV(c[k], l);
    cross3(v->n, v->x, c);
    l = norm3(c);
    for (int k = 0; k < 3; k++) v->y[k] = FB_DIV(c[k], l);
  }
  double amax = 1.0;
  for (int k = 0; k < 3; k++) {
    double hx = FB_MUL(FB_MUL(v->x[k], 0.5), plane_scale);
    double hy = FB_MUL(FB_MUL(v->y[k], 0.5), plane_scale);
    v->origin[k] = FB_SUB(FB_SUB(v->sp[k], hx), hy);
    double p1 = FB_ADD(v->origin[k], FB_MUL(v->x[k], plane_scale));
    double p2 = FB_ADD(v->origin[k], FB_MUL(v->y[k], plane_scale));
    v->v1[k] = FB_SUB(p1, v->origin[k]);
    v->v2[k] = FB_SUB(p2, v->origin[k]);
    v->delta[k] = FB_SUB(v->sp[k], FB_MUL(v->sp[k], spacing));
    v->dir[k] = -FB_MUL(FB_MUL(v->n[k], radius), 2.0);
    v->dirf[k] = (float)v->dir[k];
    v->invf[k] = 1.0f / v->dirf[k];
    // largest coordinate any ray of this view can take (origin or end)
    double reach = fabs(v->sp[k]) * fabs(spacing) + fabs(v->delta[k]) + fabs(plane_scale * spacing) +
                   fabs(v->dir[k]);
    if (reach > amax) amax = reach;
  }
  // float32 rounding of origin/direction moves a ray by < 4 ulp(amax); 64 ulp of slack.  The exact test also
  // accepts a point within tol of a triangle (dist^2 <= tol^2), i.e. up to tol outside its box: 2 tol more.
  v->slack = (float)(amax * 64.0 * 5.9604644775390625e-08 + 2.0 * fabs(tol));
  double dl = norm3(v->dir);
  v->tslack = (float)(dl > 0.0 ? 4.0 * (double)v->slack / dl : 0.0);
  v->dlen = dl;
}

// origin of ray (i, j) of a res x res plane; i is the fast axis (vtkPlaneSource
// with resolution res-1, points stored as float32)
FB_HD void view_ray(const View& v, int res, int i, int j, double spacing, double* o) {
  double tc0 = (res > 1) ? FB_DIV((double)i, (double)(res - 1)) : 0.0;
  double tc1 = (res > 1) ? FB_DIV((double)j, (double)(res - 1)) : 0.0;
  for (int k = 0; k < 3; k++) {
    float pf = (float)FB_ADD(FB_ADD(v.origin[k], FB_MUL(tc0, v.v1[k])), FB_MUL(tc1, v.v2[k]));
    o[k] = FB_ADD(FB_MUL((double)pf, spacing), v.delta[k]);
  }
}

// ---------------------------------------------------------------- exact test
// vtkTriangle::IntersectWithLine -> vtkPlane::IntersectWithLine ->
// vtkTriangle::EvaluatePosition, restated (reference call site
// src/app/fly_by_features.cxx:750-758,778).  P* are float32 vertices.
// Returns true on hit and fills t; x and w are filled when the plane is hit.
FB_HD double dist2pt(const double* a, const double* b) {
  double d0 = FB_SUB(a[0], b[0]), d1 = FB_SUB(a[1], b[1]), d2 = FB_SUB(a[2], b[2]);
  return FB_ADD(FB_ADD(FB_MUL(d0, d0), FB_MUL(d1, d1)), FB_MUL(d2, d2));
}
// vtkLine::DistanceToLine: squared distance from x to the segment p1-p2
FB_HD double dist2seg(const double* x, const double* p1, const double* p2) {
  double p21[3] = {FB_SUB(p2[0], p1[0]), FB_SUB(p2[1], p1[1]), FB_SUB(p2[2], p1[2])};
  double num = FB_ADD(FB_ADD(FB_MUL(p21[0], FB_SUB(x[0], p1[0])), FB_MUL(p21[1], FB_SUB(x[1], p1[1]))),
                      FB_MUL(p21[2], FB_SUB(x[2], p1[2])));
  double denom = dot3(p21, p21);
  double tolerance = FB_MUL(1.e-05, num);
  if (tolerance < 0.0) tolerance = -tolerance;
  if (denom < tolerance) return dist2pt(p1, x);
  double t = FB_DIV(num, denom);
  if (t < 0.0) return dist2pt(p1, x);
  if (t > 1.0) return dist2pt(p2, x);
  double c[3] = {FB_ADD(p1[0], FB_MUL(t, p21[0])), FB_ADD(p1[1], FB_MUL(t, p21[1])), FB_ADD(p1[2], FB_MUL(t, p21[2]))};
  return dist2pt(c, x);
}
// vtkTriangle::EvaluatePosition for a point outside the triangle: squared distance
// to the nearest edge / vertex, picked from the signs of the weights
// (w[k] belongs to vertex k; pt1 = P1, pt2 = P2, pt3 = P0).  Rarely executed.
// Not inlined and fed by value: the hot path then keeps every vertex in registers
// (an inlined copy made the render kernel spill to local memory).
FB_COLD double outside_dist2(double x0, double x1, double x2, double a0, double a1, double a2, double b0,
                             double b1, double b2, double c0, double c1, double c2, double w0, double w1,
                             double w2) {
  const double x[3] = {x0, x1, x2}, pt1[3] = {a0, a1, a2}, pt2[3] = {b0, b1, b2}, pt3[3] = {c0, c1, c2};
  const double w[3] = {w0, w1, w2};
  double dp, d1, d2;
  if (w[1] < 0.0 && w[2] < 0.0) {
    dp = dist2pt(x, pt3); d1 = dist2seg(x, pt1, pt3); d2 = dist2seg(x, pt3, pt2);
  } else if (w[2] < 0.0 && w[0] < 0.0) {
    dp = dist2pt(x, pt1); d1 = dist2seg(x, pt1, pt3); d2 = dist2seg(x, pt1, pt2);
  } else if (w[1] < 0.0 && w[0] < 0.0) {
    dp = dist2pt(x, pt2); d1 = dist2seg(x, pt2, pt3); d2 = dist2seg(x, pt1, pt2);
  } else if (w[0] < 0.0) {
    return dist2seg(x, pt1, pt2);
  } else if (w[1] < 0.0) {
    return dist2seg(x, pt2, pt3);
  } else if (w[2] < 0.0) {
    return dist2seg(x, pt1, pt3);
  } else {
    return DBL_MAX;
  }
  double r = (dp < d1) ? dp : d1;
  if (d2 < r) r = d2;
  return r;
}

// tol2 = tol^2 of vtkCellLocator::IntersectWithLine (1e-10 in fly_by_features.cxx:750)
FB_HD bool tri_exact(const double* o, const double* d, const float* P0, const float* P1,
                     const float* P2, double t_limit, double tol2, double* t_out, double* x, double* w) {
  double pt1[3] = {(double)P1[0], (double)P1[1], (double)P1[2]};
  double pt2[3] = {(double)P2[0], (double)P2[1], (double)P2[2]};
  double pt3[3] = {(double)P0[0], (double)P0[1], (double)P0[2]};
  double a[3] = {FB_SUB(pt3[0], pt2[0]), FB_SUB(pt3[1], pt2[1]), FB_SUB(pt3[2], pt2[2])};
  double b[3] = {FB_SUB(pt1[0], pt2[0]), FB_SUB(pt1[1], pt2[1]), FB_SUB(pt1[2], pt2[2])};
  double n[3];
  cross3(a, b, n);
  if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0) return false;
  double num = FB_SUB(dot3(n, pt1), dot3(n, o));
  double den = dot3(n, d);
  double fabsden = den < 0.0 ? -den : den;
  double fabstol = num < 0.0 ? FB_MUL(-num, 1.0e-06) : FB_MUL(num, 1.0e-06);
  if (fabsden <= fabstol) return false;
  double t = FB_DIV(num, den);
  if (!(t >= 0.0 && t <= 1.0)) return false;
  if (t > t_limit) return false;  // cannot beat the current nearest hit
  for (int k = 0; k < 3; k++) x[k] = FB_ADD(o[k], FB_MUL(t, d[k]));
  double xo[3] = {FB_SUB(x[0], pt1[0]), FB_SUB(x[1], pt1[1]), FB_SUB(x[2], pt1[2])};
  double tp = dot3(n, xo), n2 = dot3(n, n);
  // drop the dominant axis of n (first max wins); only the two kept components
  // of the projected point are needed, selected without indexed arrays
  int idx = 0;
  double maxc = 0.0;
  {
    double f0 = n[0] < 0 ? -n[0] : n[0], f1 = n[1] < 0 ? -n[1] : n[1], f2 = n[2] < 0 ? -n[2] : n[2];
    if (f0 > maxc) { maxc = f0; idx = 0; }
    if (f1 > maxc) { maxc = f1; idx = 1; }
    if (f2 > maxc) { maxc = f2; idx = 2; }
  }
  const bool a0 = (idx == 0), b2 = (idx == 2);  // kept axes: ia = a0 ? 1 : 0, ib = b2 ? 1 : 2
  double xa = a0 ? x[1] : x[0], xb = b2 ? x[1] : x[2];
  double na = a0 ? n[1] : n[0], nb = b2 ? n[1] : n[2];
  double p1a = a0 ? pt1[1] : pt1[0], p1b = b2 ? pt1[1] : pt1[2];
  double p2a = a0 ? pt2[1] : pt2[0], p2b = b2 ? pt2[1] : pt2[2];
  double p3a = a0 ? pt3[1] : pt3[0], p3b = b2 ? pt3[1] : pt3[2];
  double cpa = (n2 != 0.0) ? FB_SUB(xa, FB_DIV(FB_MUL(tp, na), n2)) : xa;
  double cpb = (n2 != 0.0) ? FB_SUB(xb, FB_DIV(FB_MUL(tp, nb), n2)) : xb;
  double r0 = FB_SUB(cpa, p3a), r1 = FB_SUB(cpb, p3b);
  double c10 = FB_SUB(p1a, p3a), c11 = FB_SUB(p1b, p3b);
  double c20 = FB_SUB(p2a, p3a), c21 = FB_SUB(p2b, p3b);
  double det = FB_SUB(FB_MUL(c10, c21), FB_MUL(c20, c11));
  if (det == 0.0) return false;
  double pc0 = FB_DIV(FB_SUB(FB_MUL(r0, c21), FB_MUL(c20, r1)), det);
  double pc1 = FB_DIV(FB_SUB(FB_MUL(c10, r1), FB_MUL(r0, c11)), det);
  w[0] = FB_SUB(1.0, FB_ADD(pc0, pc1));
  w[1] = pc0;
  w[2] = pc1;
  if (w[0] >= 0.0 && w[0] <= 1.0 && w[1] >= 0.0 && w[1] <= 1.0 && w[2] >= 0.0 && w[2] <= 1.0) {
    *t_out = t;
    return true;
  }
  // Outside.  vtkTriangle::IntersectWithLine still reports a hit when the squared
  // distance from x to the triangle is <= tol^2 (this keeps a ray through a shared
  // edge from falling between the two cells).  Cheap safe rejection first: a point
  // with weight w_k < 0 lies |w_k| * height_k beyond the edge opposite vertex k,
  // height_k = |n| / |edge_k|, so w_k^2 |n|^2 > 4 tol^2 |edge_k|^2 rules the slow
  // path out (factor 4 covers the rounding of w, n and the edge).
  {
    const double lim = FB_MUL(4.0, tol2);
    bool far = false;
    if (w[0] < 0.0) {  // edge P1-P2
      double e[3] = {FB_SUB(pt2[0], pt1[0]), FB_SUB(pt2[1], pt1[1]), FB_SUB(pt2[2], pt1[2])};
      far = far || FB_MUL(FB_MUL(w[0], w[0]), n2) > FB_MUL(lim, dot3(e, e));
    }
    if (w[1] < 0.0) {  // edge P2-P0
      double e[3] = {FB_SUB(pt3[0], pt2[0]), FB_SUB(pt3[1], pt2[1]), FB_SUB(pt3[2], pt2[2])};
      far = far || FB_MUL(FB_MUL(w[1], w[1]), n2) > FB_MUL(lim, dot3(e, e));
    }
    if (w[2] < 0.0) {  // edge P0-P1
      double e[3] = {FB_SUB(pt1[0], pt3[0]), FB_SUB(pt1[1], pt3[1]), FB_SUB(pt1[2], pt3[2])};
      far = far || FB_MUL(FB_MUL(w[2], w[2]), n2) > FB_MUL(lim, dot3(e, e));
    }
    if (far) return false;
  }
  if (outside_dist2(x[0], x[1], x[2], pt1[0], pt1[1], pt1[2], pt2[0], pt2[1], pt2[2], pt3[0], pt3[1], pt3[2],
                    w[0], w[1], w[2]) <= tol2) {
    *t_out = t;
    return true;
  }
  return false;
}

// ---------------------------------------------------------------- LBVH
// 64-byte node: the two child boxes live in the parent so one fetch decides
// both descents.  Words (DESIGN.md "Data layout"):
//   f[0..5]  child0 lo xyz, hi xyz      f[6..11] child1 lo xyz, hi xyz
//   i[12] child0  i[13] child1   (>= 0 internal node index, < 0 leaf: see leaf_first / leaf_count)
//   i[14] parent (-1 for the root)    i[15] sibling (same encoding as a child)
struct __attribute__((aligned(16))) Node {
  float box[12];
  int32_t c0, c1, parent, sibling;
};
// A negative child is a LEAF: a run of consecutive Morton-sorted triangle slots (a radix-tree
// subtree always covers a contiguous slot range, so small subtrees are collapsed into one leaf).
//   v = ~child ;  first slot = v & LEAF_FIRST_MASK ;  count = (v >> LEAF_SHIFT) + 1   (1..8)
#define FB_LEAF_SHIFT 27
#define FB_LEAF_FIRST_MASK 0x07ffffff
#define FB_LEAF_MAX 4  // triangles per collapsed leaf
FB_HD int leaf_encode(int first, int count) { return ~(first | ((count - 1) << FB_LEAF_SHIFT)); }
FB_HD int leaf_first(int child) { return (~child) & FB_LEAF_FIRST_MASK; }
FB_HD int leaf_count(int child) { return ((~child) >> FB_LEAF_SHIFT) + 1; }

FB_HD uint32_t expand10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
// 30-bit Morton code of a point in [0,1]^3 (x is the most significant axis)
FB_HD uint32_t morton30(float x, float y, float z) {
  x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
  y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
  z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
  return (expand10((uint32_t)x) << 2) | (expand10((uint32_t)y) << 1) | expand10((uint32_t)z);
}

FB_HD int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}
// common-prefix length of sorted keys i and j (-1 when j is out of range);
// equal codes fall back to the position, so every key is distinct.
FB_HD int lcp(const uint32_t* codes, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint32_t a = codes[i], b = codes[j];
  if (a != b) return clz32(a ^ b);
  return 32 + clz32((uint32_t)i ^ (uint32_t)j);
}
// Karras 2012: range and split of internal node i over n sorted keys (n >= 2).
FB_HD void karras_node(const uint32_t* codes, int n, int i, int* first, int* last, int* split) {
  int d = (lcp(codes, n, i, i + 1) - lcp(codes, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = lcp(codes, n, i, i - d);
  int lmax = 2;
  while (lcp(codes, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (lcp(codes, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = lcp(codes, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (lcp(codes, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  *split = i + s * d + (d < 0 ? -1 : 0);
  *first = d > 0 ? i : j;
  *last = d > 0 ? j : i;
}

// ---------------------------------------------------------------- traversal
struct Hit {
  double t;
  int32_t slot;  // Morton-sorted triangle slot, -1 = miss
  int32_t cell;  // original cell id
};

// Conservative float32 slab test of one child box against a segment.
FB_HD bool slab(const float* b, const float* o, const float* inv, float slack, float tcull,
                float* tnear) {
  float t0x = (b[0] - slack - o[0]) * inv[0], t1x = (b[3] + slack - o[0]) * inv[0];
  float t0y = (b[1] - slack - o[1]) * inv[1], t1y = (b[4] + slack - o[1]) * inv[1];
  float t0z = (b[2] - slack - o[2]) * inv[2], t1z = (b[5] + slack - o[2]) * inv[2];
  // fminf/fmaxf drop NaNs (0 * inf), which keeps an axis-parallel ray inside its slab
  float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
  float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tcull));
  *tnear = tn;
  return tn <= tf;
}

// Per-ray constants of the fused-multiply-add slab test used by the packet
// traversal: t = b * inv - o * inv, with the slack folded into the addends.
// Directions are uniform per view, so the near / far plane of every axis is
// known up front (neg[k] selects hi as the near plane).  |inv| is clamped so an
// axis-parallel ray (d = 0) still culls boxes it lies outside of.
struct SlabRay {
  float inv[3];
  float cn[3], cf[3];  // near: fma(b_near, inv, cn)   far: fma(b_far, inv, cf)
};
FB_HD float clamp_inv(float dirf) {
  float a = fabsf(dirf);
  float inv = a > 1e-30f ? 1.0f / a : 1e30f;
  return dirf < 0.f ? -inv : inv;  // -0.0f counts as positive: the ray does not move on that axis
}
FB_HD void slab_ray_setup(const float* of, const float* inv, float slack, SlabRay* r) {
  for (int k = 0; k < 3; k++) {
    r->inv[k] = inv[k];
    float oi = of[k] * inv[k];
    float sk = slack * fabsf(inv[k]) * 1.0001f;
    r->cn[k] = -oi - sk;
    r->cf[k] = -oi + sk;
  }
}
// b = {lo xyz, hi xyz}; neg* = direction component is negative (warp-uniform)
FB_HD bool slab_fma(const float* b, const SlabRay& r, bool negx, bool negy, bool negz, float tcull,
                    float* tnear) {
  float nx = negx ? b[3] : b[0], fx = negx ? b[0] : b[3];
  float ny = negy ? b[4] : b[1], fy = negy ? b[1] : b[4];
  float nz = negz ? b[5] : b[2], fz = negz ? b[2] : b[5];
  float tn = fmaxf(fmaxf(fmaf(nx, r.inv[0], r.cn[0]), fmaf(ny, r.inv[1], r.cn[1])),
                   fmaxf(fmaf(nz, r.inv[2], r.cn[2]), 0.0f));
  float tf = fminf(fminf(fmaf(fx, r.inv[0], r.cf[0]), fmaf(fy, r.inv[1], r.cf[1])),
                   fminf(fmaf(fz, r.inv[2], r.cf[2]), tcull));
  *tnear = tn;
  return tn <= tf;
}

// ---------------------------------------------------------------- float32 screening
// The rays of one view are parallel (orthographic), so "does the ray hit the
// triangle" is a 2-D point-in-triangle question in the view plane.  screen_tri
// answers it in float32 with explicit error margins and returns
//   SC_MISS  the exact test is certain to reject the triangle,
//   SC_SURE  the exact test is certain to accept it,
//   SC_MAYBE anything the margins cannot decide (near an edge, edge-on, slivers),
// plus a depth interval [zlo, zhi] (distance below the view plane) that contains
// the exact hit if there is one.  The render kernel uses it to run the long
// double-precision test once per ray, on the few candidates that can still be
// the nearest hit, instead of on every leaf a ray's box test touches.  It never
// decides a result: ids, t and weights always come from tri_exact.
struct ViewF {
  float xf[3], yf[3], nf[3];  // view-plane basis and viewing normal
  float rp;                   // sp . n  (distance of the plane from the origin)
  float e_p;                  // bound on the error of a projected coordinate
  float e_z;                  // bound on the error of a projected depth
  float two_r;                // length of the segment (2 R)
  float d0;                   // absolute in-plane margin on top of the rounding bounds
  float q0x, q0y, qdx, qdy;   // view-plane position of pixel (i, j): (q0x + i qdx, q0y + j qdy)
  float iqx, iqy;             // 1 / qdx, 1 / qdy
  int enabled;                // 0: the view radius is tiny against the mesh, every candidate is SC_MAYBE
};
enum { SC_MISS = 0, SC_MAYBE = 1, SC_SURE = 2 };

FB_HD void viewf_setup(const View& v, double radius, double spacing, float max_abs_coord, double tol, ViewF* f) {
  for (int k = 0; k < 3; k++) {
    f->xf[k] = (float)v.x[k];
    f->yf[k] = (float)v.y[k];
    f->nf[k] = (float)v.n[k];
  }
  double rp = v.sp[0] * v.n[0] + v.sp[1] * v.n[1] + v.sp[2] * v.n[2];
  f->rp = (float)rp;
  // e_proj: a projected coordinate is a 3-term float dot product (two fmaf and a product: three roundings of
  // partial sums bounded by |P|_2 |x|_2) of a point with |coord| <= max_abs_coord and a unit vector rounded to
  // float (one more 2^-24 |P|_2): error <= 4 * 2^-24 * sqrt(3) * max_abs_coord; twice that is used.
  // e_q: the view-plane position of a ray is taken from the pixel index, q = fmaf(i, qd, q0).  The ray the exact
  // test sees starts at a float32-rounded plane point (vtkPlaneSource): off by <= 2^-24 |spacing| |pf|_2 with
  // |pf|_2 <= |sp|_2 + |v1|_2 + |v2|_2; evaluating q in float adds <= 2 * 2^-24 |spacing| max(|v1|_2, |v2|_2)
  // (rounding of qd times i, of q0 and of the fmaf).  Twice the sum is used.
  const double sp2 = sqrt(v.sp[0] * v.sp[0] + v.sp[1] * v.sp[1] + v.sp[2] * v.sp[2]);
  const double v12 = sqrt(v.v1[0] * v.v1[0] + v.v1[1] * v.v1[1] + v.v1[2] * v.v1[2]);
  const double v22 = sqrt(v.v2[0] * v.v2[0] + v.v2[1] * v.v2[1] + v.v2[2] * v.v2[2]);
  const float e_q = (float)(1.1920928955078125e-07 * fabs(spacing) * (sp2 + v12 + v22 + 2.0 * (v12 > v22 ? v12 : v22)) * 1.0001) + 1.0e-12f;
  f->e_p = 8.3e-07f * max_abs_coord + e_q;
  f->e_z = 9.5367431640625e-07f * (max_abs_coord + 1.0f + fabsf((float)rp)) + e_q;
  f->two_r = (float)(2.0 * radius);
  // d0: absolute in-plane margin on top of the rounding bounds.  It has to cover what the exact test itself
  // adds at a boundary: its dist^2 <= tol^2 acceptance just outside an edge (tol = 1e-10 by default) and its
  // own double rounding (< 1e-11 in these units for the triangles that get past the strip test).
  f->d0 = f->e_p + (float)(2.0 * fabs(tol)) + 1.0e-9f;
  f->q0x = f->q0y = f->qdx = f->qdy = f->iqx = f->iqy = 0.f;  // set by viewf_pixels once the resolution is known
  // the SC_SURE argument needs |n.d| > 1e-6 |n.(P1-o)| for non-grazing triangles: holds when R is not tiny
  f->enabled = radius >= 0.05 * ((double)max_abs_coord + 1.0) ? 1 : 0;
}

// pixel (i, j) of a res x res plane -> ideal view-plane position of its ray (see view_setup / view_ray):
// o = sp + spacing ((i/(res-1) - 1/2) v1 + (j/(res-1) - 1/2) v2), v1 = s x, v2 = s y
FB_HD void viewf_pixels(const View& v, double spacing, int res, ViewF* f) {
  double sx = spacing * (v.v1[0] * v.x[0] + v.v1[1] * v.x[1] + v.v1[2] * v.x[2]);
  double sy = spacing * (v.v2[0] * v.y[0] + v.v2[1] * v.y[1] + v.v2[2] * v.y[2]);
  f->q0x = (float)(-0.5 * sx);
  f->q0y = (float)(-0.5 * sy);
  f->qdx = res > 1 ? (float)(sx / (double)(res - 1)) : 0.f;
  f->qdy = res > 1 ? (float)(sy / (double)(res - 1)) : 0.f;
  f->iqx = 1.0f / f->qdx;
  f->iqy = 1.0f / f->qdy;
}

// 2-D position and depth of a ray origin in the frame of its view (from the double origin)
FB_HD void screen_ray(const View& v, const double* o, float* qx, float* qy, float* zo) {
  double r[3] = {o[0] - v.sp[0], o[1] - v.sp[1], o[2] - v.sp[2]};
  *qx = (float)(r[0] * v.x[0] + r[1] * v.x[1] + r[2] * v.x[2]);
  *qy = (float)(r[0] * v.y[0] + r[1] * v.y[1] + r[2] * v.y[2]);
  *zo = (float)(-(r[0] * v.n[0] + r[1] * v.n[1] + r[2] * v.n[2]));
}

// Per-triangle part of the screening test: the projection into the view frame and everything
// that does not depend on the ray.  flags: bit 0 = the depth range overlaps the segment at all,
// bit 1 = eligible for SC_SURE (not grazing, depth range strictly inside the segment).
struct TriS {
  float ax, ay, bx, by, cx, cy;  // projected vertices
  float lox, hix, loy, hiy;      // bounding box of the projection
  float zlo, zhi;                // depth interval (distance below the view plane), widened by e_z
  float area2;                   // twice the signed projected area
  int flags;
};

FB_HD void screen_setup(const ViewF& V, float zo, const float* P0, const float* P1, const float* P2, TriS* t) {
  t->ax = fmaf(P0[0], V.xf[0], fmaf(P0[1], V.xf[1], P0[2] * V.xf[2]));
  t->ay = fmaf(P0[0], V.yf[0], fmaf(P0[1], V.yf[1], P0[2] * V.yf[2]));
  t->bx = fmaf(P1[0], V.xf[0], fmaf(P1[1], V.xf[1], P1[2] * V.xf[2]));
  t->by = fmaf(P1[0], V.yf[0], fmaf(P1[1], V.yf[1], P1[2] * V.yf[2]));
  t->cx = fmaf(P2[0], V.xf[0], fmaf(P2[1], V.xf[1], P2[2] * V.xf[2]));
  t->cy = fmaf(P2[0], V.yf[0], fmaf(P2[1], V.yf[1], P2[2] * V.yf[2]));
  t->lox = fminf(t->ax, fminf(t->bx, t->cx));
  t->hix = fmaxf(t->ax, fmaxf(t->bx, t->cx));
  t->loy = fminf(t->ay, fminf(t->by, t->cy));
  t->hiy = fmaxf(t->ay, fmaxf(t->by, t->cy));
  const float az = V.rp - fmaf(P0[0], V.nf[0], fmaf(P0[1], V.nf[1], P0[2] * V.nf[2]));
  const float bz = V.rp - fmaf(P1[0], V.nf[0], fmaf(P1[1], V.nf[1], P1[2] * V.nf[2]));
  const float cz = V.rp - fmaf(P2[0], V.nf[0], fmaf(P2[1], V.nf[1], P2[2] * V.nf[2]));
  const float zmin = fminf(az, fminf(bz, cz)), zmax = fmaxf(az, fmaxf(bz, cz));
  t->zlo = zmin - V.e_z;
  t->zhi = zmax + V.e_z;
  t->area2 = (t->bx - t->ax) * (t->cy - t->ay) - (t->by - t->ay) * (t->cx - t->ax);
  int flags = 0;
  // the segment covers depths [zo, zo + 2R]
  if (!(t->zhi < zo - V.e_z || t->zlo > zo + V.two_r + V.e_z)) flags |= 1;
  // Sure-eligible only if the triangle is not grazing -- the exact test rejects
  // |n.d| <= 1e-6 |n.(P1-o)|.  The depth of the plane changes by at most (zmax - zmin) over a
  // 2-D extent of at least the smallest altitude, which is >= |area2| / (longest edge) >= |area2| / Dq
  // (Dq = L1 size of the bounding box): tan(theta) <= dz Dq / |area2|, so dz Dq <= 500 |area2| keeps
  // cos(theta) > 1e-3 -- and only if its depth range lies strictly inside the segment.
  const float Dq = (t->hix - t->lox) + (t->hiy - t->loy);
  const float amin = fabsf(t->area2) - 8.0f * V.e_p * Dq;  // |area2| less its own rounding bound
  if ((zmax - zmin + 4.0f * V.e_z) * Dq <= 500.0f * amin && t->zlo > zo + V.e_z && t->zhi < zo + V.two_r - V.e_z)
    flags |= 2;
  t->flags = flags;
}

// Per-ray part: q = position of the ray in the view plane.
FB_HD int screen_pixel(const ViewF& V, const TriS& t, float qx, float qy) {
  if (!(t.flags & 1)) return SC_MISS;
  const float ep = V.e_p, m0 = 2.0f * ep + V.d0;
  if (qx < t.lox - m0 || qx > t.hix + m0 || qy < t.loy - m0 || qy > t.hiy + m0) return SC_MISS;
  // edge functions E_k = e_k x r_k, each with a margin that bounds its total float error:
  // moving both edge ends and the query point by ep changes E by at most 2 ep (|r|_1 + |e|_1),
  // the product rounding adds 2^-22 (|ex ry| + |ey rx|), and d0 |e|_1 (>= d0 |e|_2) is the
  // absolute distance margin.
  const float r0x = qx - t.ax, r0y = qy - t.ay, r1x = qx - t.bx, r1y = qy - t.by, r2x = qx - t.cx, r2y = qy - t.cy;
  const float e0x = t.bx - t.ax, e0y = t.by - t.ay, e1x = t.cx - t.bx, e1y = t.cy - t.by;
  const float e2x = t.ax - t.cx, e2y = t.ay - t.cy;
  const float p0a = e0x * r0y, p0b = e0y * r0x, p1a = e1x * r1y, p1b = e1y * r1x, p2a = e2x * r2y, p2b = e2y * r2x;
  const float E0 = p0a - p0b, E1 = p1a - p1b, E2 = p2a - p2b;
  const float l0 = fabsf(e0x) + fabsf(e0y), l1 = fabsf(e1x) + fabsf(e1y), l2 = fabsf(e2x) + fabsf(e2y);
  const float k2 = 2.0f * ep, kd = k2 + V.d0, kr = 2.384185791015625e-07f;
  const float g0 = fmaf(k2, fabsf(r0x) + fabsf(r0y), fmaf(kd, l0, kr * (fabsf(p0a) + fabsf(p0b))));
  const float g1 = fmaf(k2, fabsf(r1x) + fabsf(r1y), fmaf(kd, l1, kr * (fabsf(p1a) + fabsf(p1b))));
  const float g2 = fmaf(k2, fabsf(r2x) + fabsf(r2y), fmaf(kd, l2, kr * (fabsf(p2a) + fabsf(p2b))));
  const float area2 = (E0 + E1) + E2;
  const float G = g0 + g1 + g2;
  if (fabsf(area2) <= 2.0f * G) {
    // (nearly) edge-on: the orientation is not reliable, but every point of the triangle (and of its
    // d0 neighbourhood) satisfies |E_k| <= |area2| + d0 |e_k| for each edge k -- it lies in the thin
    // strip between an edge line and the opposite vertex.  Outside that strip is a certain miss.
    const float lim = fabsf(area2) + G;
    if (fabsf(E0) > lim + g0 || fabsf(E1) > lim + g1 || fabsf(E2) > lim + g2) return SC_MISS;
    return SC_MAYBE;
  }
  const float s = area2 > 0.f ? 1.0f : -1.0f;
  const float v0 = s * E0, v1 = s * E1, v2 = s * E2;
  if (v0 < -g0 || v1 < -g1 || v2 < -g2) return SC_MISS;
  if (!(v0 > g0 && v1 > g1 && v2 > g2)) return SC_MAYBE;
  return (t.flags & 2) ? SC_SURE : SC_MAYBE;
}

struct PixRange {  // pixels of a view that can see a projected bounding box
  int i0, i1, j0, j1;
};
FB_HD int screen_tri(const ViewF& V, float qx, float qy, float zo, const float* P0, const float* P1,
                     const float* P2, float* zlo, float* zhi) {
  if (!V.enabled) {
    *zlo = -INFINITY;
    *zhi = INFINITY;
    return SC_MAYBE;
  }
  TriS t;
  screen_setup(V, zo, P0, P1, P2, &t);
  *zlo = t.zlo;
  *zhi = t.zhi;
  return screen_pixel(V, t, qx, qy);
}

// ---------------------------------------------------------------- scan-conversion form (raster.cu)
// The same screening predicate with everything that does not depend on the pixel hoisted into the
// per-triangle setup.  The per-pixel margins of screen_pixel grow with |r| = |q - vertex|; inside the
// padded bounding box |r|_1 <= R1 = (box size)_1 + 2 m0, so
//     g_k = k2 R1 + kd l_k + kr l_k R1   >=   the g_k screen_pixel would use at any such pixel.
// Larger margins only turn decisions into SC_MAYBE, so the guarantees of screen_pixel carry over.
struct TriR {
  float ax, ay, bx, by, cx, cy;        // projected vertices
  float e0x, e0y, e1x, e1y, e2x, e2y;  // edges a->b, b->c, c->a
  float g0, g1, g2, G;                 // margins of the three edge functions and their sum
  float x0, x1, y0, y1;                // padded bounding box (lox - m0 ...)
  float zlo, zhi;                      // depth interval, widened by e_z
  int flags;                           // 1 depth range overlaps the segment, 2 SURE-eligible
};

// projection, padded box and pixel range; false when no pixel of the view can see the triangle's box
FB_HD bool raster_project(const ViewF& V, const float* P0, const float* P1, const float* P2, int res, TriR* t,
                          PixRange* r) {
  t->ax = fmaf(P0[0], V.xf[0], fmaf(P0[1], V.xf[1], P0[2] * V.xf[2]));
  t->ay = fmaf(P0[0], V.yf[0], fmaf(P0[1], V.yf[1], P0[2] * V.yf[2]));
  t->bx = fmaf(P1[0], V.xf[0], fmaf(P1[1], V.xf[1], P1[2] * V.xf[2]));
  t->by = fmaf(P1[0], V.yf[0], fmaf(P1[1], V.yf[1], P1[2] * V.yf[2]));
  t->cx = fmaf(P2[0], V.xf[0], fmaf(P2[1], V.xf[1], P2[2] * V.xf[2]));
  t->cy = fmaf(P2[0], V.yf[0], fmaf(P2[1], V.yf[1], P2[2] * V.yf[2]));
  // a triangle with a NaN coordinate can never be hit by the exact test (every comparison fails): drop it here,
  // fminf / fmaxf would otherwise hide the NaN and every pixel of the box would be undecided
  const float chk = ((t->ax + t->bx) + t->cx) + ((t->ay + t->by) + t->cy);
  if (chk != chk) return false;
  const float m0 = 2.0f * V.e_p + V.d0;
  t->x0 = fminf(t->ax, fminf(t->bx, t->cx)) - m0;
  t->x1 = fmaxf(t->ax, fmaxf(t->bx, t->cx)) + m0;
  t->y0 = fminf(t->ay, fminf(t->by, t->cy)) - m0;
  t->y1 = fmaxf(t->ay, fmaxf(t->by, t->cy)) + m0;
  const float ix = V.iqx, iy = V.iqy;
  // a superset of the pixels inside the padded box (raster_pixel repeats the box test on the real coordinates)
  const float fi0 = fmaxf(0.0f, ceilf((t->x0 - V.q0x) * ix - 0.01f)), fi1 = fminf((float)(res - 1), floorf((t->x1 - V.q0x) * ix + 0.01f));
  const float fj0 = fmaxf(0.0f, ceilf((t->y0 - V.q0y) * iy - 0.01f)), fj1 = fminf((float)(res - 1), floorf((t->y1 - V.q0y) * iy + 0.01f));
  if (!(fi1 >= fi0 && fj1 >= fj0)) return false;
  r->i0 = (int)fi0; r->i1 = (int)fi1; r->j0 = (int)fj0; r->j1 = (int)fj1;
  return true;
}

// the rest of the per-triangle work (depth range, flags, margins); zo = depth of the ray origins
FB_HD void raster_setup(const ViewF& V, float zo, const float* P0, const float* P1, const float* P2, TriR* t) {
  const float az = V.rp - fmaf(P0[0], V.nf[0], fmaf(P0[1], V.nf[1], P0[2] * V.nf[2]));
  const float bz = V.rp - fmaf(P1[0], V.nf[0], fmaf(P1[1], V.nf[1], P1[2] * V.nf[2]));
  const float cz = V.rp - fmaf(P2[0], V.nf[0], fmaf(P2[1], V.nf[1], P2[2] * V.nf[2]));
  const float zmin = fminf(az, fminf(bz, cz)), zmax = fmaxf(az, fmaxf(bz, cz));
  t->zlo = zmin - V.e_z;
  t->zhi = zmax + V.e_z;
  t->e0x = t->bx - t->ax; t->e0y = t->by - t->ay;
  t->e1x = t->cx - t->bx; t->e1y = t->cy - t->by;
  t->e2x = t->ax - t->cx; t->e2y = t->ay - t->cy;
  const float area2 = t->e0x * (t->cy - t->ay) - t->e0y * (t->cx - t->ax);
  const float m0 = 2.0f * V.e_p + V.d0;
  const float Dq = ((t->x1 - t->x0) + (t->y1 - t->y0)) - 4.0f * m0;  // L1 size of the unpadded box (screen_setup's Dq)
  int flags = 0;
  if (!(t->zhi < zo - V.e_z || t->zlo > zo + V.two_r + V.e_z)) flags |= 1;
  // same SURE eligibility as screen_setup, with Dq rounded up
  const float Dq_up = fmaxf(Dq, 0.0f) * 1.000001f + 4.0f * V.e_p;
  const float amin = fabsf(area2) - 8.0f * V.e_p * Dq_up;
  // V.enabled (a guard of the LBVH screening for radii that are tiny against the mesh) is not needed here: the exact
  // test's plane rejection |n.d| <= 1e-6 |n.(P1 - o)| cannot fire for a triangle that passes this test whatever the
  // radius, because its depth range lies inside the segment: |n.(P1 - o)| = |n| cos(theta) z_hit <= |n| 2R cos(theta)
  // while |n.d| = |n| 2R cos(theta) -- a ratio of at most 1 against the 1e-6 of the rejection
  if ((zmax - zmin + 4.0f * V.e_z) * Dq_up <= 500.0f * amin && t->zlo > zo + V.e_z && t->zhi < zo + V.two_r - V.e_z)
    flags |= 2;
  // >= |r|_1 from any vertex to any pixel raster_project lets through: the padded box plus the slack of its
  // pixel range (0.01 pixel per side and the rounding of the index arithmetic, < 0.02 pixel in all)
  const float R1 = ((t->x1 - t->x0) + (t->y1 - t->y0) + 0.04f * (fabsf(V.qdx) + fabsf(V.qdy))) * 1.000001f;
  const float l0 = fabsf(t->e0x) + fabsf(t->e0y), l1 = fabsf(t->e1x) + fabsf(t->e1y), l2 = fabsf(t->e2x) + fabsf(t->e2y);
  // kr = 2^-20: 2^-22 bounds the rounding of raster_pixel's own products and differences; the linear form below
  // (raster_linear_*) adds < 2^-22 l_k R1 of its own; the sum is doubled like every other margin.  Against k2 R1 the
  // term is ~0.5 % for the triangles of a 100k-2M surface.
  const float k2 = 2.0f * V.e_p, kd = k2 + V.d0, kr = 9.5367431640625e-07f;
  t->g0 = fmaf(k2, R1, fmaf(kd, l0, kr * l0 * R1)) * 1.000001f;
  t->g1 = fmaf(k2, R1, fmaf(kd, l1, kr * l1 * R1)) * 1.000001f;
  t->g2 = fmaf(k2, R1, fmaf(kd, l2, kr * l2 * R1)) * 1.000001f;
  t->G = (t->g0 + t->g1 + t->g2) * 1.000001f;
  t->flags = flags;
}

// q must be a pixel of the range raster_project returned (that is what bounds |r| in the margins)
FB_HD int raster_pixel(const TriR& t, float qx, float qy) {
  const float r0x = qx - t.ax, r0y = qy - t.ay, r1x = qx - t.bx, r1y = qy - t.by, r2x = qx - t.cx, r2y = qy - t.cy;
  const float p0a = t.e0x * r0y, p0b = t.e0y * r0x, p1a = t.e1x * r1y, p1b = t.e1y * r1x, p2a = t.e2x * r2y, p2b = t.e2y * r2x;
  const float E0 = p0a - p0b, E1 = p1a - p1b, E2 = p2a - p2b;
  const float area2 = (E0 + E1) + E2;
  if (fabsf(area2) <= 2.0f * t.G) {  // (nearly) edge-on: strip test, see screen_pixel
    const float lim = fabsf(area2) + t.G;
    if (fabsf(E0) > lim + t.g0 || fabsf(E1) > lim + t.g1 || fabsf(E2) > lim + t.g2) return SC_MISS;
    return SC_MAYBE;
  }
  const float s = area2 > 0.f ? 1.0f : -1.0f;
  const float v0 = s * E0, v1 = s * E1, v2 = s * E2;
  if (v0 < -t.g0 || v1 < -t.g1 || v2 < -t.g2) return SC_MISS;
  if (!(v0 > t.g0 && v1 > t.g1 && v2 > t.g2)) return SC_MAYBE;
  return (t.flags & 2) ? SC_SURE : SC_MAYBE;
}

// ---- linear form of raster_pixel: the three edge functions are affine in the pixel index, so the per-pixel work is
// two fused multiply-adds per edge instead of the differences and products of raster_pixel:
//     E_k(i0 + di, j0 + dj) = C_k + A_k di + B_k dj,   A_k = -e_ky qdx,  B_k = e_kx qdy,
// C_k = E_k at the corner pixel (i0, j0) of the triangle's pixel range, evaluated exactly as raster_pixel evaluates it
// there.  What the linear form adds to raster_pixel's error at a pixel of the range (|di qdx| + |dj qdy| <= R1):
//   rounding of A_k, B_k (one product each):  2^-24 (|e_ky| |di qdx| + |e_kx| |dj qdy|)  <=  2^-24 l_k R1
//   the two fmaf roundings:                   2 * 2^-24 max|E_k|                          <=  2^-23 l_k R1 (1 + 2^-20)
// i.e. < 2^-22 l_k R1, which raster_setup's kr carries.  The position the form evaluates at is (float corner) +
// (di qdx, dj qdy) in exact arithmetic: one fmaf rounding away from the lattice q0 + i qd -- the same distance
// raster_pixel's fmaf(i, qd, q0) has, and what e_q (viewf_setup) budgets for.  The twice-area is the corner pixel's
// (any pixel's value is within G of the true one, which is all the orientation / strip argument uses); its sign is
// folded into the coefficients.  Decisions: exactly those of raster_pixel with E_k replaced -- MISS / SURE / MAYBE
// keep their meaning, audited pixel by pixel against the exact test by tests/hostsim.
struct TriL {
  float A0, A1, A2, B0, B1, B2, C0, C1, C2;  // sign-folded (orientation +1) unless strip
  float g0, g1, g2;
  float lim;   // strip: |area2| + G
  int flags;   // 1 strip test (nearly edge-on), 2 SURE-eligible
};

FB_HD void raster_linear_setup(const ViewF& V, const TriR& t, int i0, int j0, TriL* L) {
  const float qx = fmaf((float)i0, V.qdx, V.q0x), qy = fmaf((float)j0, V.qdy, V.q0y);
  const float r0x = qx - t.ax, r0y = qy - t.ay, r1x = qx - t.bx, r1y = qy - t.by, r2x = qx - t.cx, r2y = qy - t.cy;
  const float p0a = t.e0x * r0y, p0b = t.e0y * r0x, p1a = t.e1x * r1y, p1b = t.e1y * r1x, p2a = t.e2x * r2y, p2b = t.e2y * r2x;
  const float E0 = p0a - p0b, E1 = p1a - p1b, E2 = p2a - p2b;
  const float area2 = (E0 + E1) + E2;
  const bool strip = fabsf(area2) <= 2.0f * t.G;
  const float s = (strip || area2 > 0.f) ? 1.0f : -1.0f;
  L->A0 = -s * (t.e0y * V.qdx); L->A1 = -s * (t.e1y * V.qdx); L->A2 = -s * (t.e2y * V.qdx);
  L->B0 = s * (t.e0x * V.qdy); L->B1 = s * (t.e1x * V.qdy); L->B2 = s * (t.e2x * V.qdy);
  L->C0 = s * E0; L->C1 = s * E1; L->C2 = s * E2;
  L->g0 = t.g0; L->g1 = t.g1; L->g2 = t.g2;
  L->lim = fabsf(area2) + t.G;
  L->flags = (strip ? 1 : 0) | (t.flags & 2);
}

// (di, dj) = pixel offsets from the corner as floats (exact: small integers).  The two variants are the two branches
// of raster_pixel; raster_kernel picks the first for a whole warp when none of its triangles is edge-on.
FB_HD int raster_linear_pixel_plain(const TriL& L, float di, float dj) {
  const float E0 = fmaf(L.A0, di, fmaf(L.B0, dj, L.C0));
  const float E1 = fmaf(L.A1, di, fmaf(L.B1, dj, L.C1));
  const float E2 = fmaf(L.A2, di, fmaf(L.B2, dj, L.C2));
  // (no short-circuit: the device code is three comparisons per line into one predicate, no branches)
  const bool miss = (E0 < -L.g0) | (E1 < -L.g1) | (E2 < -L.g2);
  const bool inside = (E0 > L.g0) & (E1 > L.g1) & (E2 > L.g2) & ((L.flags & 2) != 0);
  return miss ? SC_MISS : (inside ? SC_SURE : SC_MAYBE);
}
FB_HD int raster_linear_pixel_strip(const TriL& L, float di, float dj) {
  const float E0 = fmaf(L.A0, di, fmaf(L.B0, dj, L.C0));
  const float E1 = fmaf(L.A1, di, fmaf(L.B1, dj, L.C1));
  const float E2 = fmaf(L.A2, di, fmaf(L.B2, dj, L.C2));
  if (fabsf(E0) > L.lim + L.g0 || fabsf(E1) > L.lim + L.g1 || fabsf(E2) > L.lim + L.g2) return SC_MISS;
  return SC_MAYBE;
}
FB_HD int raster_linear_pixel(const TriL& L, float di, float dj) {
  return (L.flags & 1) ? raster_linear_pixel_strip(L, di, dj) : raster_linear_pixel_plain(L, di, dj);
}

// Work items of raster_kernel: a box of w x h pixels is cut into ipr = ceil(w / U) items per row; item k of the box is
// row k / ipr.  The division is a multiply-shift by rt_magic(ipr) = floor(2^16 / ipr) + 1 (the float quotient is within
// 0.004 of 2^16 / ipr, whose distance from the next integer below is >= 1 / ipr unless it is one itself):
// (k * magic) >> 16 == k / ipr for every k < 2^9 and ipr <= 2^7 -- boxes have at most 96 pixels (tests/hostsim checks all).
FB_HD uint32_t rt_magic(int ipr) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)(65536.0f * __frcp_rn((float)ipr)) + 1u;
#else
  return (uint32_t)(65536.0f * (1.0f / (float)ipr)) + 1u;  // (1 / x in float is correctly rounded, as __frcp_rn)
#endif
}

// The flat enumeration of raster_kernel, lane by lane (tests/hostsim replays a warp with loops and these functions).
// Records (the warp's visible triangles, compacted) own consecutive runs of work items; S = first item of a record.
//   rt_start_bit   what record `lane` contributes to the OR of the window [base, base + 32): the bit of its first item
//   rt_owner       the record that owns item base + lane: records started before the window + starts at or below the lane
//   rt_item        item k of a box of width w, U pixels per item: row and first column
FB_HD uint32_t rt_start_bit(int S, int base) {
  const uint32_t rel = (uint32_t)(S - base);
  return rel < 32u ? 1u << rel : 0u;
}
FB_HD int rt_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
FB_HD uint32_t rt_lemask(int lane) { return 0xffffffffu >> (31 - lane); }  // bits 0..lane
FB_HD int rt_owner(uint32_t starts, int started, uint32_t lemask) { return started + rt_popc(starts & lemask) - 1; }
FB_HD void rt_item(int k, int w, uint32_t magic, int U, int* row, int* col0) {
  *row = (int)(((uint32_t)k * magic) >> 16);
  *col0 = (k - *row * ((w + U - 1) / U)) * U;
}

// ---- front-buffer key of the single-pass scan conversion (raster.cu):
//   [63:32] far depth bound zhi as an order-preserving integer   [31:27] thickness code   [26:0] slot
// The thickness code c bounds the triangle's depth extent: zhi - zlo < 2^(c - 24) (c = 31: no bound), so that
// whoever displaces a front entry can tell, without fetching the triangle, whether it can still be a candidate.
FB_HD uint32_t ford_bits(float f) {
#if defined(__CUDA_ARCH__)
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
FB_HD float unford_bits(uint32_t u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
FB_HD uint64_t front_key(float zlo, float zhi, int slot) {
  const float thick = zhi - zlo;  // > 0: the interval is widened by e_z on both sides
#if defined(__CUDA_ARCH__)
  uint32_t tb = __float_as_uint(thick);
#else
  uint32_t tb;
  memcpy(&tb, &thick, 4);
#endif
  int c = (int)((tb >> 23) & 0xffu) - 127 + 25;  // thick < 2^(e + 1) = 2^(c - 24)
  if (!(thick > 0.0f) || c > 31) c = 31;
  if (c < 0) c = 0;
  return ((uint64_t)ford_bits(zhi) << 32) | ((uint64_t)c << 27) | (uint32_t)slot;
}
// lower bound of the near depth bound of the triangle a key belongs to (twice the coded extent: the spare
// factor absorbs the rounding of the subtraction)
FB_HD float front_key_zlo_lower(uint64_t key) {
  const int c = (int)((key >> 27) & 31u);
  if (c == 31) return -INFINITY;
  const float zhi = unford_bits((uint32_t)(key >> 32));
#if defined(__CUDA_ARCH__)
  return zhi - __int_as_float((c + 104) << 23);  // 2^(c - 23), built from its exponent field
#else
  return zhi - ldexpf(1.0f, c - 23);
#endif
}
#define FB_FRONT_SLOT(key) ((int)((uint32_t)(key) & 0x07ffffffu))

// Stackless traversal: a 64-bit trail holds one "far sibling still pending" bit
// per level, parent / sibling links walk back up.  Depth <= 62 is guaranteed by
// the 30+32-bit key length of the LBVH.  NodeFetch: node(i) -> Node by value,
// parent(i), sibling(i) -> links of node i.
// Tie-break: smallest t, then lowest original cell id.
template <class NodeFetch, class Counter>
FB_HD Hit traverse(const NodeFetch& fetch, const float4* __restrict__ tv0,
                   const float4* __restrict__ tv1, const float4* __restrict__ tv2, const double* o,
                   const double* d, const float* dirf, const float* invf, float slack, float tslack,
                   double tol2, Counter& cnt, const ViewF* vf = nullptr, float sqx = 0.f, float sqy = 0.f,
                   float szo = 0.f) {
  // vf != nullptr: the float32 screening test runs first and the exact test is skipped on a certain miss
  (void)dirf;
  Hit best;
  best.t = DBL_MAX;
  best.slot = -1;
  best.cell = 0x7fffffff;
  float of[3] = {(float)o[0], (float)o[1], (float)o[2]};
  float tcull = 1.0f + tslack;
  uint64_t trail = 1;
  int node = 0;
  for (;;) {
    const Node N = fetch.node(node);
    cnt.node();
    float tn0, tn1;
    bool h0 = slab(N.box, of, invf, slack, tcull, &tn0);
    bool h1 = slab(N.box + 6, of, invf, slack, tcull, &tn1);
    int c0 = N.c0, c1 = N.c1;
#pragma unroll
    for (int side = 0; side < 2; side++) {
      bool h = side ? h1 : h0;
      int c = side ? c1 : c0;
      if (h && c < 0) {
        const int first = leaf_first(c), count = leaf_count(c);
        for (int slot = first; slot < first + count; slot++) {
          float4 a = tv0[slot], b = tv1[slot], e = tv2[slot];
          cnt.tri();
          float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
          if (vf) {
            float zl, zh;
            if (screen_tri(*vf, sqx, sqy, szo, P0, P1, P2, &zl, &zh) == SC_MISS) continue;
          }
          double t, x[3], w[3];
          if (tri_exact(o, d, P0, P1, P2, best.t, tol2, &t, x, w)) {
#if defined(__CUDA_ARCH__)
            int cell = __float_as_int(a.w);
#else
            int cell;
            memcpy(&cell, &a.w, 4);
#endif
            if (t < best.t || (t == best.t && cell < best.cell)) {
              best.t = t;
              best.slot = slot;
              best.cell = cell;
              tcull = fminf(tcull, FB_F2F_RU(t) + tslack);
            }
          }
        }
      }
    }
    h0 = h0 && c0 >= 0 && tn0 <= tcull;
    h1 = h1 && c1 >= 0 && tn1 <= tcull;
    if (h0 || h1) {
      bool both = h0 && h1;
      node = both ? (tn1 < tn0 ? c1 : c0) : (h0 ? c0 : c1);
      trail = (trail << 1) | (both ? 1u : 0u);
      continue;
    }
    // climb until a level with a pending sibling (or the sentinel)
    for (;;) {
      if (trail & 1) break;
      trail >>= 1;
      node = fetch.parent(node);
    }
    if (trail == 1) break;
    trail ^= 1;
    node = fetch.sibling(node);
  }
  if (best.slot < 0) best.cell = -1;
  return best;
}

FB_HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}

struct NoCount {
  FB_HD void node() {}
  FB_HD void tri() {}
};

// ---------------------------------------------------------------- shading
// src/app/fly_by_features.cxx:766-807 (barycentric normal, z, pixel) and
// src/py/fly_by_features.py:124-150 (use_z / split_z channel assembly).
// shade_z: the same with the depth supplied by the caller (z already scaled); enc 3 = the 8-bit colour an
// OpenGL read-back returns (round(255 c), clamped), kept as float
FB_HD void shade_z(const double* w, const float* n0, const float* n1, const float* n2, double z, int use_z,
                   int split_z, int enc, float* px) {
  double nn[3];
  for (int k = 0; k < 3; k++) {
    double s = FB_ADD(0.0, FB_MUL((double)n0[k], w[0]));
    s = FB_ADD(s, FB_MUL((double)n1[k], w[1]));
    nn[k] = FB_ADD(s, FB_MUL((double)n2[k], w[2]));
  }
  double e[3];
  for (int k = 0; k < 3; k++) {
    if (enc == 2) e[k] = nn[k];
    else if (enc == 1) e[k] = FB_MUL(FB_ADD(FB_MUL(nn[k], 0.5), 0.5), 255.0);
    else if (enc == 3) {
      double q = floor(FB_ADD(FB_MUL(FB_ADD(FB_MUL(nn[k], 0.5), 0.5), 255.0), 0.5));
      e[k] = q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q);
    } else e[k] = FB_ADD(FB_MUL(nn[k], 0.5), 0.5);
  }
  if (!use_z) {
    for (int k = 0; k < 3; k++) px[k] = (float)e[k];
  } else if (split_z) {
    for (int k = 0; k < 3; k++) px[k] = (float)e[k];
    px[3] = (float)z;
  } else {
    for (int k = 0; k < 3; k++) px[k] = (float)FB_MUL(e[k], z);
  }
}

FB_HD void shade(const double* o, const double* x, const double* w, const float* n0,
                 const float* n1, const float* n2, int use_z, int split_z, int enc, double zscale,
                 float* px) {
  double dd[3] = {FB_SUB(o[0], x[0]), FB_SUB(o[1], x[1]), FB_SUB(o[2], x[2])};
  shade_z(w, n0, n1, n2, FB_MUL(norm3(dd), zscale), use_z, split_z, enc, px);
}

// ---------------------------------------------------------------- tolerance shading (opt-in)
// The exact chain above (tri_exact + shade: ~200 dependent double-precision instructions, five IEEE divisions and a
// square root) exists to make the values written to the tensor bit-equal to the reference arithmetic; the task's
// bar for depth and normals is 1e-5 relative.  For a pixel whose ONLY candidate is a certain hit (SC_SURE: the exact
// test is known to accept it, so the cell id needs no arithmetic at all) shade_approx computes
//   * the hit parameter, the depth and the hit point in double (fused multiply-adds, one Newton reciprocal; the ray
//     origin stays the exact float32-rounded vtkPlaneSource point), error ~1e-15 / cos(incidence);
//   * the barycentric weights and the interpolated normal in float32, from the hit point taken RELATIVE TO VERTEX 0
//     (so the float32 rounding is relative to the triangle's size, not to the coordinates' magnitude).
// Returns false (caller runs the exact chain) for slivers, where float32 weights would lose more than ~1e-6.
// Measured against the exact chain: depth <= 1e-7 relative, normal channels <= ~1e-6 absolute (tests, bench line).
// Cell ids are not affected.  Selected per render with FLYBY_SHADE_TOLERANCE in flyby_render_opts.mode.
FB_HD double rcp_newton(double a) {  // 1 / a to ~2 ulp for |a| in the float range
  double r = (double)FB_FRCP((float)a);
  r = FB_FMA(r, FB_FMA(-a, r, 1.0), r);
  r = FB_FMA(r, FB_FMA(-a, r, 1.0), r);
  return r;
}
FB_HD bool shade_approx(const double* o, const double* d, double dlen, const float* P0, const float* P1, const float* P2,
                        const float* n0, const float* n1, const float* n2, int use_z, int split_z, int enc,
                        double zscale, float* px) {
  if (enc == 3) return false;  // 8-bit colours: a float32 floor() could land one code off
  const double p0[3] = {(double)P0[0], (double)P0[1], (double)P0[2]};
  const double a[3] = {(double)P1[0] - p0[0], (double)P1[1] - p0[1], (double)P1[2] - p0[2]};  // edge P0 -> P1
  const double b[3] = {(double)P2[0] - p0[0], (double)P2[1] - p0[1], (double)P2[2] - p0[2]};  // edge P0 -> P2
  const double n[3] = {FB_FMA(a[1], b[2], -(a[2] * b[1])), FB_FMA(a[2], b[0], -(a[0] * b[2])),
                       FB_FMA(a[0], b[1], -(a[1] * b[0]))};
  const double q[3] = {p0[0] - o[0], p0[1] - o[1], p0[2] - o[2]};
  const double num = FB_FMA(n[0], q[0], FB_FMA(n[1], q[1], n[2] * q[2]));
  const double den = FB_FMA(n[0], d[0], FB_FMA(n[1], d[1], n[2] * d[2]));
  const float aden = (float)fabs(den);
  if (!(aden > 1e-30f) || !(aden < 1e30f)) return false;
  const double rden = rcp_newton(den);
  double t = num * rden;
  t = FB_FMA(FB_FMA(-t, den, num), rden, t);
  if (!(t >= 0.0 && t <= 1.0)) return false;
  // hit point relative to vertex 0: r = (o - p0) + t d
  const float r[3] = {(float)FB_FMA(t, d[0], -q[0]), (float)FB_FMA(t, d[1], -q[1]), (float)FB_FMA(t, d[2], -q[2])};
  const float af[3] = {(float)a[0], (float)a[1], (float)a[2]}, bf[3] = {(float)b[0], (float)b[1], (float)b[2]};
  const float f0 = (float)fabs(n[0]), f1 = (float)fabs(n[1]), f2 = (float)fabs(n[2]);
  // drop the dominant axis of n; the two kept components are selected without indexed arrays
  const bool d0 = f0 >= f1 && f0 >= f2, d2 = !d0 && f2 > f1;  // dominant axis 0 / 2 (else 1)
  const float aa = d0 ? af[1] : af[0], ab = d2 ? af[1] : af[2];
  const float ba = d0 ? bf[1] : bf[0], bb = d2 ? bf[1] : bf[2];
  const float ra = d0 ? r[1] : r[0], rb = d2 ? r[1] : r[2];
  const float det = fmaf(aa, bb, -(ba * ab));
  const float scale = (fabsf(aa) + fabsf(ab)) * (fabsf(ba) + fabsf(bb));
  if (!(fabsf(det) > 0.02f * scale)) return false;  // sliver: the weights would amplify float32 rounding by > 50
  const float rdet = 1.0f / det;
  const float w1 = fmaf(ra, bb, -(ba * rb)) * rdet;  // weight of P1
  const float w2 = fmaf(aa, rb, -(ra * ab)) * rdet;  // weight of P2
  const float w0 = 1.0f - w1 - w2;
  const float z = (float)(t * dlen * zscale);
  float e[3];
  for (int k = 0; k < 3; k++) {
    const float nn = fmaf(w0, n0[k], fmaf(w1, n1[k], w2 * n2[k]));
    if (enc == 2) e[k] = nn;
    else {
      e[k] = fmaf(nn, 0.5f, 0.5f);
      if (enc == 1) e[k] *= 255.0f;
    }
  }
  if (!use_z) { px[0] = e[0]; px[1] = e[1]; px[2] = e[2]; }
  else if (split_z) { px[0] = e[0]; px[1] = e[1]; px[2] = e[2]; px[3] = z; }
  else { px[0] = e[0] * z; px[1] = e[1] * z; px[2] = e[2] * z; }
  return true;
}

// ---------------------------------------------------------------- pinhole camera (raster-compatible views)
// What FlyByGenerator.getFlyBy sets up for OpenGL (src/py/fly_by_features.py:85-107): position = the
// viewpoint, focal point = origin, view-up (0,0,-1), or (1,0,0) / (-1,0,0) exactly at the north / south
// pole; vtkCamera's default view angle is 30 degrees.  f = direction of projection, s = image x (right),
// u = image y (up, orthogonalised as vtkCamera does); rows run bottom to top as vtkWindowToImageFilter
// returns them.  A pixel's ray goes through its centre: d = f + sx tan(a/2) s + sy tan(a/2) u, so that the
// parameter along d is the depth along the viewing direction (the linearised z-buffer of :127-128).
struct Camera {
  double eye[3], f[3], s[3], u[3];
};
FB_HD void camera_setup(const float* spf, Camera* c) {
  double len = FB_SQRT(FB_ADD(FB_ADD(FB_MUL((double)spf[0], (double)spf[0]), FB_MUL((double)spf[1], (double)spf[1])),
                              FB_MUL((double)spf[2], (double)spf[2])));
  double nz = FB_DIV((double)spf[2], len);
  double up[3] = {0.0, 0.0, -1.0};
  if (nz == 1.0) { up[0] = 1.0; up[2] = 0.0; }
  else if (nz == -1.0) { up[0] = -1.0; up[2] = 0.0; }
  for (int k = 0; k < 3; k++) {
    c->eye[k] = (double)spf[k];
    c->f[k] = FB_DIV(-(double)spf[k], len);
  }
  double s[3];
  cross3(c->f, up, s);
  double sl = norm3(s);
  for (int k = 0; k < 3; k++) c->s[k] = FB_DIV(s[k], sl);
  cross3(c->s, c->f, c->u);
}
// segment of pixel (i, j): from depth z_near to depth z_far along the pixel's direction
FB_HD void camera_ray(const Camera& c, int res, int i, int j, double tan_half, double z_near, double z_far,
                      double* o, double* d) {
  double sx = FB_MUL(FB_SUB(FB_MUL(FB_DIV(FB_ADD((double)i, 0.5), (double)res), 2.0), 1.0), tan_half);
  double sy = FB_MUL(FB_SUB(FB_MUL(FB_DIV(FB_ADD((double)j, 0.5), (double)res), 2.0), 1.0), tan_half);
  double span = FB_SUB(z_far, z_near);
  for (int k = 0; k < 3; k++) {
    double dir = FB_ADD(FB_ADD(c.f[k], FB_MUL(sx, c.s[k])), FB_MUL(sy, c.u[k]));
    o[k] = FB_ADD(c.eye[k], FB_MUL(dir, z_near));
    d[k] = FB_MUL(dir, span);
  }
}

}  // namespace fb
