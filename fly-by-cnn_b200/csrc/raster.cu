// raster.cu -- the default render path: triangle-centric candidate search + exact resolve.
//
// The rays of a view are parallel and sit on a regular pixel grid, and a triangle of a
// 100k-2M triangle surface covers only a few pixels.  Instead of walking the LBVH once per
// ray packet, every triangle is projected into the view frame once and tested -- with the
// float32 screening predicate of flyby_core.cuh, which never decides a result -- against the
// handful of pixels inside its bounding box:
//   raster    front[pixel] = min over its SURE triangles of (far depth bound, slot): one 64-bit atomicMin
//             whose return value tells the triangle whom it displaced / lost to; every triangle that is not
//             a certain miss and can still be at or in front of the front joins extra[pixel] (settle_pixel)
//   resolve1  one thread per pixel in raster order: background / the exact test on the front
//             triangle when it is the only candidate (the common case) / queue the pixel
//   big       triangles whose box exceeds RT_BIG pixels are queued by `raster` and walked here in row bands
//             spread over the whole grid (coarse meshes, close-ups), with the same settle_pixel
//   resolve2  queued pixels: exact test on the front triangle and the extras, tie-break (t, cell id)
//   spill     pixels with more than 8 extras (edge-on piles at silhouettes): the rest of their candidates
//             sits in one global (pixel, slot) list; three small kernels take the minimum of (t, cell id)
//   resolve3  last resort, only when that list itself was full: exact LBVH re-trace of the pixel
// The candidate set is a superset of what the exact test can accept at or in front of the
// nearest certain hit, so the result equals the LBVH traversal and the oracle bit for bit; the
// order in which atomics land does not matter.
//
// Reference replaced: the per-view / per-plane-point loops of
// src/app/fly_by_features.cxx:652-852 (search = IntersectWithLine :750-758, shading :766-807)
// and the channel assembly of src/py/fly_by_features.py:124-150.
#include <stdlib.h>

#include "trace_common.cuh"

namespace fb {

#ifndef RT_THREADS
#define RT_THREADS 128
#endif
#ifndef RT_MINB
#define RT_MINB 12
#endif
#define RT_BIG 96    // bounding boxes with more pixels than this go to raster_big_kernel (row bands over the whole grid)
#define RT_EXTRA 8    // extra candidates kept per pixel (two int4)
#define RT_BAND 4096  // pixels of a big triangle's box handled by one warp of raster_big_kernel
#define RV_THREADS 128
#ifndef RV_SHARE
#define RV_SHARE 4         // FLYBY_SHARE_GPU: persistent CTAs of raster_resolve1_kernel per SM (exact / tolerance variant)
#endif
#ifndef RV_SHARE_APPROX
#define RV_SHARE_APPROX 6
#endif
#ifndef RV_MINB
#define RV_MINB 5
#endif

struct RasterParams {
  RenderParams r;
  unsigned long long* front;  // [n_pix] front_key() of the nearest SURE triangle: far depth bound | thickness code | slot
  int* cnt;                   // [n_pix] extra candidates found
  int* extra;                 // [n_pix, 4] extra slots 0..3
  int* extra2;                // [n_pix, 4] extra slots 4..7 (touched only by pixels that need them)
  uint32_t* queue;            // [2, n_pix] pixels for resolve2, pixels for resolve3
  unsigned int* n_queue;      // [2]
  // spill: candidates beyond RT_EXTRA of a pixel (edge-on piles at silhouettes) go to one global pair list
  uint2* pairs;               // [pair_cap] (pixel, slot)
  double* pair_t;             // [pair_cap] exact t of the pair, < 0 = no hit
  unsigned int* n_spill;      // [0] pairs appended, [1] overflow pixels numbered
  uint32_t* ov_pix;           // [pair_cap] pixel of overflow record k
  unsigned long long *ov_prov_t, *ov_prov_cs;  // [pair_cap] best of the front + stored extras: t bits, (cell << 32) | slot
  unsigned long long *ov_best_t, *ov_best_cs;  // [pair_cap] running minimum over everything / over the spilled pairs
  unsigned int pair_cap;
  // triangles of more than RT_BIG pixels are queued and rasterised in row bands by the whole grid (raster_big_kernel)
  int4* big;                        // [big_cap] slot, local view, first band, rows per band
  unsigned long long* big_counter;  // (entries << 32) | bands, advanced by one 64-bit atomicAdd per entry
  unsigned int big_cap;
  int64_t n_pix;
  int T;
};
#define RT_LOST 0x40000000  // cnt flag: a spilled pair found the list full, the pixel is re-traced through the LBVH

__device__ __forceinline__ void append_extra(const RasterParams& p, uint32_t pix, int slot) {
  const int k = atomicAdd(p.cnt + pix, 1) & (RT_LOST - 1);
  if (k < 4) p.extra[(size_t)pix * 4 + k] = slot;
  else if (k < RT_EXTRA) p.extra2[(size_t)pix * 4 + (k - 4)] = slot;
  else {
    const unsigned int i2 = atomicAdd(p.n_spill, 1u);
    if (i2 < p.pair_cap) p.pairs[i2] = make_uint2(pix, (unsigned int)slot);
    else atomicOr(p.cnt + pix, RT_LOST);
  }
}

// What a triangle does at a pixel where it is not a certain miss, given the front entry it met there
// (`old`: the value its atomicMin replaced or kept if it is a certain hit, else the value it read):
//   certain hit, became the front: the entry it displaced stays a candidate if its depth range can reach this one
//   certain hit, did not:          it is a candidate if its near depth bound is not behind the front it met
//   undecided:                     likewise
// The front only ever moves nearer, so every test against an intermediate front is weaker than the test
// against the final one: the extras are a superset of what pass-by-pass scan conversion would list, for any
// order of arrival.
__device__ __forceinline__ void settle_pixel(const RasterParams& p, uint32_t pix, int cls, unsigned long long key,
                                             unsigned long long old, unsigned int zlo_ord, float zhi) {
  if (cls == SC_SURE && key < old) {
    if (old != ~0ull && front_key_zlo_lower(old) <= zhi) append_extra(p, pix, FB_FRONT_SLOT(old));
  } else if (zlo_ord <= (unsigned int)(old >> 32)) {
    append_extra(p, pix, FB_FRONT_SLOT(key));
  }
}

// ---- raster_kernel: one warp = 32 consecutive triangle slots of one view, in two phases.
//   set-up   one lane per triangle: projection, pixel range, depth interval, margins, and the linear form of the
//            three edge functions about the corner pixel of the range (raster_linear_setup); the lanes whose triangle
//            can be seen write an 80-byte record to the warp's shared memory
//   pixels   the boxes of the warp's triangles are cut into row quads (4 pixels of a row) and the quads of ALL 32
//            triangles are enumerated flat over the lanes -- a lane is busy whatever the size of "its" triangle was.
//            The owner of quad q is found with one warp-wide OR: every record marks the bit of its first quad in
//            the current window of 32, owner = records started before the window + popcount below the lane.
//            A pixel test is two fused multiply-adds per edge and six comparisons (raster_linear_pixel).
//   settle   pixels that are not certain misses are compacted into a per-warp queue; whenever 32 have gathered,
//            the warp does their front atomics / loads and settle_pixel with every lane busy.
// Boxes of more than RT_BIG pixels go to raster_big_kernel as before; if its queue is full, the warp walks them
// itself, one at a time.
#ifndef RT_U
#define RT_U 4   // pixels of a row per work item (A/B on the bench workload: 1 -> 0.710 ms, 2 -> 0.674, 4 -> 0.656)
#endif
struct __align__(16) RtRec {
  float A0, A1, A2, B0;
  float B1, B2, C0, C1;
  float C2, g0, g1, g2;
  float lim; uint32_t ij0; uint32_t wf; uint32_t magic;  // corner pixel i0 | j0 << 16; width | TriL flags << 16; ceil(2^16 / items per row)
  uint32_t key_lo, key_hi, zlo_ord; float zhi;
};
#define RT_WARPS (RT_THREADS / 32)
#define RT_QCAP (32 + 32 * RT_U)  // settle queue per warp: flushed down to < 32 after every window, which adds at most 32 RT_U

// 32 (or the last n) queued pixels: entry = record | SURE << 5 | row << 6 | column << 13, both relative to the corner
__device__ __forceinline__ void rt_settle32(const RasterParams& p, const RtRec* rec, const uint32_t* q, int at, int n, int lane,
                                            uint32_t view_base, uint32_t res) {
  if (lane < n) {
    const uint32_t e = q[at + lane];
    const RtRec* r = rec + (e & 31u);
    const uint4 kz = *reinterpret_cast<const uint4*>(&r->key_lo);
    const uint32_t ij0 = r->ij0;
    const uint32_t pix = view_base + ((ij0 >> 16) + ((e >> 6) & 127u)) * res + (ij0 & 0xffffu) + (e >> 13);
    const unsigned long long key = ((unsigned long long)kz.y << 32) | kz.x;
    const int cls = (e & 32u) ? SC_SURE : SC_MAYBE;
    const unsigned long long old = cls == SC_SURE ? atomicMin(p.front + pix, key) : __ldcg(p.front + pix);
    settle_pixel(p, pix, cls, key, old, kz.z, __uint_as_float(kz.w));
  }
}

template <bool STRIP>
__device__ __forceinline__ void rt_window(const TriL& L, float di0, float dj, int nvalid, uint32_t ent0, uint32_t* q, int& qn, unsigned ltmask) {
#pragma unroll
  for (int u = 0; u < RT_U; u++) {
    const int c = STRIP ? raster_linear_pixel(L, di0 + (float)u, dj) : raster_linear_pixel_plain(L, di0 + (float)u, dj);
    const bool keep = (c != SC_MISS) & (u < nvalid);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    uint32_t* at = q + (qn + __popc(m & ltmask));  // computed by every lane so that the store is a single predicated instruction
    const uint32_t ent = ent0 + ((uint32_t)u << 13) + (c == SC_SURE ? 32u : 0u);
    if (keep) *at = ent;
    qn += __popc(m);
  }
}

__global__ void __launch_bounds__(RT_THREADS, RT_MINB) raster_kernel(const RasterParams p) {
  __shared__ ViewF s_vf;
  __shared__ RtRec s_rec[RT_WARPS][32];
  __shared__ int s_pre[RT_WARPS][32];
  __shared__ uint32_t s_q[RT_WARPS][RT_QCAP];
  const int view = blockIdx.y;  // local view index of this render call
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int slot = blockIdx.x * RT_THREADS + threadIdx.x;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.r.viewsf + p.r.view_begin + view);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_vf);
    if (threadIdx.x < (int)(sizeof(ViewF) / 4)) dst[threadIdx.x] = __ldg(src + threadIdx.x);
  }
  __syncthreads();
  const ViewF& V = s_vf;
  const int res = p.r.res;
  const uint32_t view_base = (uint32_t)view * (uint32_t)res * (uint32_t)res;  // a render call has < 2^32 pixels
  RtRec* rec = s_rec[wid];
  uint32_t* q = s_q[wid];
  const unsigned ltmask = (1u << lane) - 1u;

  // ---- set-up
  int npx = 0;
  {
    TriR t;
    PixRange r;
    r.i0 = r.i1 = r.j0 = r.j1 = 0;
    if (slot < p.T) {
      const float4 a = __ldg(p.r.tv0 + slot), b = __ldg(p.r.tv1 + slot), e = __ldg(p.r.tv2 + slot);
      const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
      if (raster_project(V, P0, P1, P2, res, &t, &r)) {  // triangles outside a close-up view end here
        npx = (r.i1 - r.i0 + 1) * (r.j1 - r.j0 + 1);
        raster_setup(V, 0.0f, P0, P1, P2, &t);  // ray origins lie in the view plane: depth 0 (+- e_z)
        if (!(t.flags & 1)) npx = 0;
      }
    }
    // large triangles (coarse meshes, close-ups) are queued for raster_big_kernel, which spreads them over the grid
    // in bands of about RT_BAND pixels
    if (npx > RT_BIG) {
      const int w = r.i1 - r.i0 + 1, rows = r.j1 - r.j0 + 1;
      int rpb = RT_BAND / w;
      rpb = rpb < 1 ? 1 : (rpb > rows ? rows : rpb);
      const unsigned int nb = (unsigned int)((rows + rpb - 1) / rpb);
      const unsigned long long at = atomicAdd(p.big_counter, (1ull << 32) | nb);
      const unsigned int e = (unsigned int)(at >> 32);
      if (e < p.big_cap) {
        const bool fits = (at & 0xffffffffull) + nb < 0x80000000ull;  // the band counter must never carry into the entry count
        p.big[e] = make_int4(slot, view, (int)(unsigned int)(at & 0xffffffffull), fits ? rpb : 0);
        if (fits) npx = 0;
      }
    }
    const bool small = npx > 0 && npx <= RT_BIG;
    const unsigned alive = __ballot_sync(0xffffffffu, small);
    if (alive) {
      const int w = r.i1 - r.i0 + 1, ipr = (w + RT_U - 1) / RT_U;  // work items per row
      const int n_mine = small ? ipr * (r.j1 - r.j0 + 1) : 0;
      int incl = n_mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      bool strip = false;
      if (small) {
        TriL L;
        raster_linear_setup(V, t, r.i0, r.j0, &L);
        strip = L.flags & 1;
        const unsigned long long key = front_key(t.zlo, t.zhi, slot);
        RtRec* o = rec + __popc(alive & ltmask);
        const uint32_t magic = rt_magic(ipr);  // (k magic) >> 16 == k / ipr
        *reinterpret_cast<float4*>(&o->A0) = make_float4(L.A0, L.A1, L.A2, L.B0);
        *reinterpret_cast<float4*>(&o->B1) = make_float4(L.B1, L.B2, L.C0, L.C1);
        *reinterpret_cast<float4*>(&o->C2) = make_float4(L.C2, L.g0, L.g1, L.g2);
        *reinterpret_cast<uint4*>(&o->lim) = make_uint4(__float_as_uint(L.lim), (uint32_t)r.i0 | ((uint32_t)r.j0 << 16),
                                                        (uint32_t)w | ((uint32_t)L.flags << 16), magic);
        *reinterpret_cast<uint4*>(&o->key_lo) = make_uint4((uint32_t)key, (uint32_t)(key >> 32), ford_bits(t.zlo), __float_as_uint(t.zhi));
        s_pre[wid][__popc(alive & ltmask)] = incl - n_mine;
      }
      const int n_items = __shfl_sync(0xffffffffu, incl, 31);
      const bool any_strip = __any_sync(0xffffffffu, strip);
      __syncwarp();
      // ---- pixels
      const int n_rec = __popc(alive);
      const int S = lane < n_rec ? s_pre[wid][lane] : 0x7fffffff;  // first work item of record `lane`
      const unsigned lemask = rt_lemask(lane);
      int started = 0;  // records that start before the current window
      int qn = 0;       // entries in the settle queue
      for (int base = 0; base < n_items; base += 32) {
        const unsigned starts = __reduce_or_sync(0xffffffffu, rt_start_bit(S, base));
        const int item = base + lane;
        const bool valid = item < n_items;
        const int own = valid ? rt_owner(starts, started, lemask) : 0;
        started += __popc(starts);
        const int k = item - __shfl_sync(0xffffffffu, S, own);
        const RtRec* o = rec + own;
        const float4 f0 = *reinterpret_cast<const float4*>(&o->A0), f1 = *reinterpret_cast<const float4*>(&o->B1);
        const float4 f2 = *reinterpret_cast<const float4*>(&o->C2);
        const uint4 f3 = *reinterpret_cast<const uint4*>(&o->lim);
        TriL L;
        L.A0 = f0.x; L.A1 = f0.y; L.A2 = f0.z; L.B0 = f0.w; L.B1 = f1.x; L.B2 = f1.y; L.C0 = f1.z; L.C1 = f1.w;
        L.C2 = f2.x; L.g0 = f2.y; L.g1 = f2.z; L.g2 = f2.w; L.lim = __uint_as_float(f3.x);
        L.flags = (int)(f3.z >> 16);
        const int w = (int)(f3.z & 0xffffu);
        int row, col0;
        rt_item(k, w, f3.w, RT_U, &row, &col0);
        const int nvalid = valid ? w - col0 : 0;
        const uint32_t ent0 = (uint32_t)own | ((uint32_t)row << 6) | ((uint32_t)col0 << 13);
        if (any_strip) rt_window<true>(L, (float)col0, (float)row, nvalid, ent0, q, qn, ltmask);
        else rt_window<false>(L, (float)col0, (float)row, nvalid, ent0, q, qn, ltmask);
        __syncwarp();
        while (qn >= 32) {
          qn -= 32;
          rt_settle32(p, rec, q, qn, 32, lane, view_base, (uint32_t)res);
        }
        __syncwarp();
      }
      if (qn) rt_settle32(p, rec, q, 0, qn, lane, view_base, (uint32_t)res);
    }
  }
  // a big triangle that found the queue full: the warp walks its box itself (set-up redone by every lane)
  unsigned big = __ballot_sync(0xffffffffu, npx > RT_BIG);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1;
    const int uslot = __shfl_sync(0xffffffffu, slot, src);
    const float4 a = __ldg(p.r.tv0 + uslot), b = __ldg(p.r.tv1 + uslot), e = __ldg(p.r.tv2 + uslot);
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
    TriR u;
    PixRange r;
    if (!raster_project(V, P0, P1, P2, res, &u, &r)) continue;
    raster_setup(V, 0.0f, P0, P1, P2, &u);
    const unsigned long long key = front_key(u.zlo, u.zhi, uslot);
    const unsigned int zlo_ord = ford_bits(u.zlo);
    const int w = r.i1 - r.i0 + 1;
    const long long n = (long long)w * (r.j1 - r.j0 + 1);
    for (long long k = lane; k < n; k += 32) {
      const int i = r.i0 + (int)(k % w), j = r.j0 + (int)(k / w);
      const int cls = raster_pixel(u, fmaf((float)i, V.qdx, V.q0x), fmaf((float)j, V.qdy, V.q0y));
      if (cls == SC_MISS) continue;
      const uint32_t pix = view_base + (uint32_t)j * (uint32_t)res + (uint32_t)i;
      const unsigned long long old = cls == SC_SURE ? atomicMin(p.front + pix, key) : __ldcg(p.front + pix);
      settle_pixel(p, pix, cls, key, old, zlo_ord, u.zhi);
    }
  }
}

// ---- big triangles, band by band.  Entry e covers the bands [big[e].z, big[e + 1].z); a warp finds the entry of
// its band by bisection (the bases grow with the entry index: both come from one atomic), redoes the triangle's
// set-up (cheap next to a band of pixels) and walks the band 32 pixels at a time.
__global__ void __launch_bounds__(128) raster_big_kernel(const RasterParams p) {
  const unsigned long long total = *p.big_counter;
  unsigned int n_ent = (unsigned int)(total >> 32);
  if (n_ent > p.big_cap) n_ent = p.big_cap;  // entries beyond the capacity were rasterised inline by raster_kernel
  const unsigned int n_bands = (unsigned int)(total & 0xffffffffull);
  if (n_ent == 0) return;
  const int lane = threadIdx.x & 31;
  const unsigned int warps = gridDim.x * (blockDim.x >> 5);
  const int res = p.r.res;
  for (unsigned int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < n_bands; b += warps) {
    unsigned int lo = 0, hi = n_ent - 1;  // the entry with the largest first band <= b (entry 0 starts at band 0)
    while (lo < hi) {
      const unsigned int mid = (lo + hi + 1) >> 1;
      if ((unsigned int)p.big[mid].z <= b) lo = mid; else hi = mid - 1;
    }
    const int4 ent = p.big[lo];
    const int slot = ent.x, view = ent.y, rpb = ent.w;
    if (rpb <= 0) continue;  // an entry that was not accepted
    const ViewF& V = p.r.viewsf[p.r.view_begin + view];
    const float4 a = __ldg(p.r.tv0 + slot), bb = __ldg(p.r.tv1 + slot), e = __ldg(p.r.tv2 + slot);
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
    TriR u;
    PixRange r;
    if (!raster_project(V, P0, P1, P2, res, &u, &r)) continue;
    raster_setup(V, 0.0f, P0, P1, P2, &u);
    const int jb = r.j0 + (int)(b - (unsigned int)ent.z) * rpb;
    if (jb > r.j1) continue;  // a band of an entry that lies behind the stored ones
    const int je = min(r.j1, jb + rpb - 1);
    const unsigned long long key = front_key(u.zlo, u.zhi, slot);
    const unsigned int zlo_ord = ford_bits(u.zlo);
    const uint32_t view_base = (uint32_t)view * (uint32_t)res * (uint32_t)res;
    const float qdx = V.qdx, q0x = V.q0x, qdy = V.qdy, q0y = V.q0y;
    const int w = r.i1 - r.i0 + 1;
    const int n = w * (je - jb + 1);
    for (int k = lane; k < n; k += 32) {
      const int i = r.i0 + k % w, j = jb + k / w;
      const int cls = raster_pixel(u, fmaf((float)i, qdx, q0x), fmaf((float)j, qdy, q0y));
      if (cls == SC_MISS) continue;
      const uint32_t pix = view_base + (uint32_t)j * (uint32_t)res + (uint32_t)i;
      const unsigned long long old = cls == SC_SURE ? atomicMin(p.front + pix, key) : __ldcg(p.front + pix);
      settle_pixel(p, pix, cls, key, old, zlo_ord, u.zhi);
    }
  }
}

__global__ void raster_init_kernel(unsigned long long* front, int* cnt, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    front[i] = ~0ull;
    cnt[i] = 0;
  }
}

// ---- resolve1: every pixel, raster order
#define RV_STAGE 64  // per-warp staging of queued pixels: one global atomic per 32+ entries instead of one per iteration
#ifndef RV_MINB_APPROX
#define RV_MINB_APPROX 8  // 64 registers: 0.60 ms against 0.64 (6) and 0.78 (10) on the bench workload (profiles/r02_ab_series.txt)
#endif
// per-warp staging of queued pixels, flushed with one global atomic per 32+ entries
__device__ __forceinline__ void stage_pixels(const RasterParams& rp, uint32_t* stage, int& n_stage, bool mine, uint32_t pix,
                                             int lane) {
  const unsigned mq = __ballot_sync(0xffffffffu, mine);
  if (!mq) return;
  if (mine) stage[n_stage + __popc(mq & ((1u << lane) - 1u))] = pix;
  n_stage += __popc(mq);
  __syncwarp();
  if (n_stage > RV_STAGE - 32) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(rp.n_queue, (unsigned int)n_stage);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < n_stage; k += 32) rp.queue[base + k] = stage[k];
    n_stage = 0;
    __syncwarp();
  }
}

// APPROX (FLYBY_SHADE_TOLERANCE): a pixel whose only candidate is a certain hit is shaded by shade_approx; the exact
// chain is not part of this variant at all (fewer registers, more resident warps) -- a pixel it declines (sliver
// triangle) joins the queue of resolve2, which runs the exact chain on the front candidate.
template <bool APPROX>
__global__ void __launch_bounds__(RV_THREADS, APPROX ? RV_MINB_APPROX : RV_MINB) raster_resolve1_kernel(const RasterParams rp) {
  const RenderParams& p = rp.r;
  __shared__ uint32_t s_stage[RV_THREADS / 32][RV_STAGE];
  const int lane = threadIdx.x & 31;
  uint32_t* stage = s_stage[threadIdx.x >> 5];
  int n_stage = 0;  // warp-uniform
  unsigned int hits = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // the loop bound is rounded up to whole warps so that the votes below see all 32 lanes
  const int64_t n_round = (rp.n_pix + 31) & ~(int64_t)31;
  for (int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; pix < n_round; pix += stride) {
    const bool valid = pix < rp.n_pix;
    unsigned long long f = ~0ull;
    int nc = 0;
    if (valid) {
      f = __ldcs(rp.front + pix);
      nc = __ldcs(rp.cnt + pix);
    }
    if (!APPROX) stage_pixels(rp, stage, n_stage, nc > 0, (uint32_t)pix, lane);  // more than the front candidate: resolve2
    bool declined = false;
    if (valid && nc == 0) {
      if (f == ~0ull) {
        const float zero[4] = {0.f, 0.f, 0.f, 0.f};
        write_pixel(p, pix, zero, -1, -1.0);
      } else {
        const int slot = FB_FRONT_SLOT(f);
        const float4* rec = p.trec + 4 * (int64_t)slot;
        const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2), g = __ldg(rec + 3);
        const PixelRay r = pixel_ray(p, pix);
        const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
        const int pid0 = __float_as_int(b.w), pid1 = __float_as_int(e.w), pid2 = __float_as_int(g.x), cell = __float_as_int(a.w);
        if (APPROX) {
          const float4 n0 = __ldg(p.normals + pid0), n1 = __ldg(p.normals + pid1), n2 = __ldg(p.normals + pid2);
          const float N0[3] = {n0.x, n0.y, n0.z}, N1[3] = {n1.x, n1.y, n1.z}, N2[3] = {n2.x, n2.y, n2.z};
          float px[4] = {0.f, 0.f, 0.f, 0.f};
          if (shade_approx(r.o, r.d, r.dlen, P0, P1, P2, N0, N1, N2, p.use_z, p.split_z, p.enc, p.zscale, px)) {
            hits++;
            write_pixel(p, pix, px, cell, -1.0);
            write_hit_scalars(p, pix, nullptr, pid0, pid1, pid2, cell);
          } else {
            declined = true;
          }
        } else {
          double t, x[3], w[3];
          if (tri_exact(r.o, r.d, P0, P1, P2, DBL_MAX, p.tol2, &t, x, w)) {
            hits++;
            shade_and_write(p, pix, r, t, w, pid0, pid1, pid2, cell);
          } else {
            const float zero[4] = {0.f, 0.f, 0.f, 0.f};
            write_pixel(p, pix, zero, -1, -1.0);
          }
        }
      }
    }
    if (APPROX) stage_pixels(rp, stage, n_stage, nc > 0 || declined, (uint32_t)pix, lane);
  }
  __syncwarp();
  if (n_stage) {
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(rp.n_queue, (unsigned int)n_stage);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < n_stage; k += 32) rp.queue[base + k] = stage[k];
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o2);
  if (lane == 0 && hits) atomicAdd(p.counters + 3, (unsigned long long)hits);
}

// Shading of a decided pixel outside resolve1.  In tolerance mode (rp.r.approx) the winner is shaded by shade_approx
// here as well, so that every hit pixel is the same function of (pixel, winning triangle) whichever kernel finishes
// it -- which pixels carry extra candidates depends on the order the atomics of raster_kernel land in.
static __device__ __forceinline__ void shade_decided(const RenderParams& p, int64_t pix, const PixelRay& r, double t,
                                                   double w0, double w1, double w2, int slot, int cell) {
  const float4* rec = p.trec + 4 * (int64_t)slot;
  const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2), g = __ldg(rec + 3);
  const int pid0 = __float_as_int(b.w), pid1 = __float_as_int(e.w), pid2 = __float_as_int(g.x);
  if (p.approx) {
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
    const float4 n0 = __ldg(p.normals + pid0), n1 = __ldg(p.normals + pid1), n2 = __ldg(p.normals + pid2);
    const float N0[3] = {n0.x, n0.y, n0.z}, N1[3] = {n1.x, n1.y, n1.z}, N2[3] = {n2.x, n2.y, n2.z};
    float px[4] = {0.f, 0.f, 0.f, 0.f};
    if (shade_approx(r.o, r.d, r.dlen, P0, P1, P2, N0, N1, N2, p.use_z, p.split_z, p.enc, p.zscale, px)) {
      write_pixel(p, pix, px, cell, -1.0);
      write_hit_scalars(p, pix, nullptr, pid0, pid1, pid2, cell);
      return;
    }
  }
  const double w[3] = {w0, w1, w2};
  shade_and_write(p, pix, r, t, w, pid0, pid1, pid2, cell);
}

// ---- resolve2: queued pixels (near edges, silhouettes, layered geometry)
template <bool APPROX>
__global__ void __launch_bounds__(RV_THREADS, 4) raster_resolve2_kernel(const RasterParams rp) {
  const RenderParams& p = rp.r;
  const unsigned int n = *rp.n_queue;
  unsigned int hits = 0;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t pix = rp.queue[k];
    const int nc_all = __ldcs(rp.cnt + pix);
    if (nc_all & RT_LOST) {  // part of its candidates were lost: resolve3 re-traces it
      rp.queue[rp.n_pix + atomicAdd(rp.n_queue + 1, 1u)] = (uint32_t)pix;
      continue;
    }
    const int nc = min(nc_all, RT_EXTRA);
    const unsigned long long f = __ldcs(rp.front + pix);
    const int4 c0 = __ldcs(reinterpret_cast<const int4*>(rp.extra) + pix);
    int4 c1 = make_int4(-1, -1, -1, -1);
    if (nc > 4) c1 = __ldcs(reinterpret_cast<const int4*>(rp.extra2) + pix);
    const PixelRay r = pixel_ray(p, pix);
    double best_t = DBL_MAX, bw[3] = {0.0, 0.0, 0.0};
    int best_slot = -1, best_cell = 0x7fffffff;
    for (int i = (f == ~0ull ? 0 : -1); i < nc; i++) {
      const int4 cc = i < 4 ? c0 : c1;
      const int j = i & 3;
      const int slot = i < 0 ? FB_FRONT_SLOT(f) : (j == 0 ? cc.x : (j == 1 ? cc.y : (j == 2 ? cc.z : cc.w)));
      const float4* rec = p.trec + 4 * (int64_t)slot;
      const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2), g = __ldg(rec + 3);
      const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
      double tt, x[3], w[3];
      if (tri_exact(r.o, r.d, P0, P1, P2, best_t, p.tol2, &tt, x, w)) {
        const int c = __float_as_int(a.w);
        if (tt < best_t || (tt == best_t && c < best_cell)) {
          best_t = tt; best_slot = slot; best_cell = c;
          bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2];
        }
      }
    }
    if (nc_all > RT_EXTRA) {
      // more candidates wait in the pair list: park the best so far in overflow record k, the spill kernels finish it
      const unsigned int k2 = atomicAdd(rp.n_spill + 1, 1u);
      const unsigned long long tb = (unsigned long long)__double_as_longlong(best_t + 0.0);
      rp.ov_pix[k2] = (uint32_t)pix;
      rp.ov_prov_t[k2] = tb;
      rp.ov_prov_cs[k2] = best_slot >= 0 ? (((unsigned long long)(unsigned int)best_cell << 32) | (unsigned int)best_slot) : ~0ull;
      rp.ov_best_t[k2] = tb;
      rp.ov_best_cs[k2] = ~0ull;
      rp.cnt[pix] = -(int)(k2 + 1);
      continue;
    }
    if (best_slot >= 0) {
      hits++;
      if (APPROX) {
        shade_decided(p, pix, r, best_t, bw[0], bw[1], bw[2], best_slot, best_cell);
      } else {  // the exact chain's weights are at hand: shade in line
        const float4* rec = p.trec + 4 * (int64_t)best_slot;
        shade_and_write(p, pix, r, best_t, bw, __float_as_int(__ldg(rec + 1).w), __float_as_int(__ldg(rec + 2).w),
                        __float_as_int(__ldg(rec + 3).x), best_cell);
      }
    } else {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      write_pixel(p, pix, zero, -1, -1.0);
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o2);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd(p.counters + 3, (unsigned long long)hits);
}

// ---- spill kernels: one thread per spilled (pixel, slot) pair.  The nearest hit of an overflow pixel is the
// lexicographic minimum of (t, cell id) over its parked best and its pairs: pass 1 takes the minimum of t
// (the bits of a non-negative double order like the value), pass 2 the minimum cell id among the pairs that
// reach it, pass 3 shades.  Every step is a minimum, so the order of the atomics does not matter.
__global__ void __launch_bounds__(RV_THREADS) spill_t_kernel(const RasterParams rp) {
  const RenderParams& p = rp.r;
  const unsigned int n = min(rp.n_spill[0], rp.pair_cap);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint2 pr = rp.pairs[i];
    const int k = -rp.cnt[pr.x] - 1;
    double tt = -1.0;
    if (k >= 0) {  // (k < 0: the pixel lost pairs and is re-traced)
      const float4* rec = p.trec + 4 * (int64_t)pr.y;
      const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2);
      const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
      const PixelRay r = pixel_ray(p, pr.x);
      double x[3], w[3], t;
      if (tri_exact(r.o, r.d, P0, P1, P2, __longlong_as_double((long long)rp.ov_prov_t[k]), p.tol2, &t, x, w)) {
        tt = t + 0.0;
        atomicMin(rp.ov_best_t + k, (unsigned long long)__double_as_longlong(tt));
      }
    }
    rp.pair_t[i] = tt;
  }
}

__global__ void __launch_bounds__(RV_THREADS) spill_cell_kernel(const RasterParams rp) {
  const RenderParams& p = rp.r;
  const unsigned int n = min(rp.n_spill[0], rp.pair_cap);
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double tt = rp.pair_t[i];
    if (tt < 0.0) continue;
    const uint2 pr = rp.pairs[i];
    const int k = -rp.cnt[pr.x] - 1;
    if ((unsigned long long)__double_as_longlong(tt) != rp.ov_best_t[k]) continue;
    const int cell = __float_as_int(__ldg(p.trec + 4 * (int64_t)pr.y).w);
    atomicMin(rp.ov_best_cs + k, ((unsigned long long)(unsigned int)cell << 32) | pr.y);
  }
}

template <bool APPROX>
__global__ void __launch_bounds__(RV_THREADS) spill_shade_kernel(const RasterParams rp) {
  const RenderParams& p = rp.r;
  const unsigned int n = rp.n_spill[1];
  unsigned int hits = 0;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t pix = rp.ov_pix[k];
    unsigned long long cs = rp.ov_best_cs[k];
    if (rp.ov_prov_t[k] == rp.ov_best_t[k] && rp.ov_prov_cs[k] < cs) cs = rp.ov_prov_cs[k];
    if (cs == ~0ull) {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      write_pixel(p, pix, zero, -1, -1.0);
      continue;
    }
    const int slot = (int)(unsigned int)cs;
    const float4* rec = p.trec + 4 * (int64_t)slot;
    const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2), g = __ldg(rec + 3);
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
    const PixelRay r = pixel_ray(p, pix);
    double t, x[3], w[3];
    tri_exact(r.o, r.d, P0, P1, P2, DBL_MAX, p.tol2, &t, x, w);  // hit by construction; recomputed for the weights
    hits++;
    if (APPROX) shade_decided(p, pix, r, t, w[0], w[1], w[2], slot, __float_as_int(a.w));
    else shade_and_write(p, pix, r, t, w, __float_as_int(b.w), __float_as_int(e.w), __float_as_int(g.x), __float_as_int(a.w));
  }
  if (hits) atomicAdd(p.counters + 3, (unsigned long long)hits);
  if (n && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.counters + 4, (unsigned long long)n);  // pixels finished from the spill list
}

// ---- resolve3: the few pixels with more candidates than the lists hold: one thread each re-traces
// its ray through the LBVH (screening test first, exact test on what is left)
template <bool APPROX>
__global__ void __launch_bounds__(64) raster_resolve3_kernel(const RasterParams rp) {
  const RenderParams& p = rp.r;
  const unsigned int n = rp.n_queue[1];
  const int64_t per_view = (int64_t)p.res * p.res;
  unsigned int hits = 0;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t pix = rp.queue[rp.n_pix + k];
    const int v = (int)(pix / per_view), loc = (int)(pix % per_view);
    const ViewF* vf = p.viewsf + p.view_begin + v;
    const float qx = fmaf((float)(loc % p.res), vf->qdx, vf->q0x), qy = fmaf((float)(loc / p.res), vf->qdy, vf->q0y);
    const PixelRay r = pixel_ray(p, pix);
    const ExactHit h = resolve_overflow(p, p.view_begin + v, r.o[0], r.o[1], r.o[2], vf->enabled ? vf : nullptr, qx, qy);
    if (h.slot >= 0) {
      hits++;
      if (APPROX) {
        shade_decided(p, pix, r, h.t, h.w0, h.w1, h.w2, h.slot, h.cell);
      } else {
        const int4 vid = p.tvid[h.slot];
        const double w[3] = {h.w0, h.w1, h.w2};
        shade_and_write(p, pix, r, h.t, w, vid.x, vid.y, vid.z, h.cell);
      }
    } else {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      write_pixel(p, pix, zero, -1, -1.0);
    }
  }
  if (hits) atomicAdd(p.counters + 3, (unsigned long long)hits);
  if (n && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.counters + 5, (unsigned long long)n);  // rays re-traced (rare)
}

int raster_render(flyby_ctx* c, const RenderParams& rp) {
  cudaStream_t st = c->stream;
  const int64_t n_pix = (int64_t)rp.n_views * rp.res * rp.res;
  RasterParams p;
  p.r = rp;
  // tolerance shading applies when no exact hit parameter is asked for and no scalar needs the exact weights
  const bool approx = rp.approx && !rp.tt && !(rp.n_scal && rp.scal_mode == 1);
  p.r.approx = approx ? 1 : 0;
  p.n_pix = n_pix;
  p.T = (int)c->mesh.T;
  FB_CUDA(c, c->cand.reserve(sizeof(int4) * (size_t)n_pix));
  FB_CUDA(c, c->cand2.reserve(sizeof(int4) * (size_t)n_pix));
  {
    const size_t cap0 = c->raster_tmp.cap;
    FB_CUDA(c, c->raster_tmp.reserve(12 * (size_t)n_pix));  // (growing it frees the old buffer: cudaFree waits for the device)
    if (c->raster_tmp.cap != cap0) c->front_clean = 0;      // a new allocation: nothing is initialised
  }
  FB_CUDA(c, c->pix_lists.reserve(8 * (size_t)n_pix + 16));
  // the counts sit behind the fronts of the whole allocation, so that both keep their place when n_pix changes
  const size_t cap_pix = c->raster_tmp.cap / 12;
  p.front = c->raster_tmp.as<unsigned long long>();
  p.cnt = reinterpret_cast<int*>(p.front + cap_pix);
  p.extra = c->cand.as<int>();
  p.extra2 = c->cand2.as<int>();
  p.queue = c->pix_lists.as<uint32_t>();
  p.n_queue = reinterpret_cast<unsigned int*>(c->counters.as<unsigned long long>() + 14);
  p.n_spill = reinterpret_cast<unsigned int*>(c->counters.as<unsigned long long>() + 13);
  p.big_counter = c->counters.as<unsigned long long>() + 12;
  FB_CUDA(c, cudaMemsetAsync(p.big_counter, 0, 24, st));  // big_counter, n_spill[2], n_queue[2]
  p.big_cap = 1u << 20;
  {  // FLYBY_BIG_CAP: test hook that shrinks the queue so that the inline fallback of raster_kernel gets exercised
    static long env_big = -1;
    if (env_big < 0) {
      const char* e = getenv("FLYBY_BIG_CAP");
      env_big = e ? atol(e) : 0;
    }
    if (env_big > 0 && env_big < (long)p.big_cap) p.big_cap = (unsigned int)env_big;
  }
  FB_CUDA(c, c->raster_big.reserve(sizeof(int4) * (size_t)p.big_cap));
  p.big = c->raster_big.as<int4>();
  size_t cap = (size_t)(n_pix / 8 > (1 << 20) ? n_pix / 8 : (1 << 20));
  {  // FLYBY_PAIR_CAP: test hook that shrinks the spill list so that the LBVH re-trace of resolve3 gets exercised
    static long env_cap = -1;
    if (env_cap < 0) {
      const char* e = getenv("FLYBY_PAIR_CAP");
      env_cap = e ? atol(e) : 0;
    }
    if (env_cap > 0) cap = (size_t)env_cap;
  }
  p.pair_cap = (unsigned int)cap;
  FB_CUDA(c, c->raster_ovf.reserve(cap * (8 + 8 + 4 + 4 * 8)));
  p.pairs = c->raster_ovf.as<uint2>();
  p.pair_t = reinterpret_cast<double*>(p.pairs + cap);
  p.ov_prov_t = reinterpret_cast<unsigned long long*>(p.pair_t + cap);
  p.ov_prov_cs = p.ov_prov_t + cap;
  p.ov_best_t = p.ov_prov_cs + cap;
  p.ov_best_cs = p.ov_best_t + cap;
  p.ov_pix = reinterpret_cast<uint32_t*>(p.ov_best_cs + cap);
  // front / count: what the side stream re-initialised behind the previous render is taken as it is (its event
  // is waited for in any case: the buffer must not be written while that kernel runs), the rest is set here
  const int64_t clean = c->front_clean;
  c->front_clean = 0;  // dirty from here on
  // (no-op before the first record; inside a captured step only the prologue's record may be waited for)
  if (c->ev_front && (c->capturing ? c->cap_front_forked : c->ev_front_live)) FB_CUDA(c, cudaStreamWaitEvent(st, c->ev_front, 0));
  if (n_pix > clean) {
    raster_init_kernel<<<(unsigned)cdiv(n_pix - clean, 256), 256, 0, st>>>(p.front + clean, p.cnt + clean, n_pix - clean);
    FB_LAUNCH_CHECK(c);
  }
  { int rc = slab_mark(c, 0); if (rc) return rc; }
  dim3 grid((unsigned)cdiv(p.T, RT_THREADS), (unsigned)rp.n_views);
  raster_kernel<<<grid, RT_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  raster_big_kernel<<<(unsigned)c->sm_count * 8, 128, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  { int rc = slab_mark(c, 1); if (rc) return rc; }  // between the candidate search and the exact resolve
  {  // the shading needs the point normals, the last-resort fallback the LBVH: both are finished on the side stream
    int rc = wait_tree(c);
    if (rc) return rc;
  }
  long long rgrid = (long long)c->sm_count * (approx ? RV_MINB_APPROX : RV_MINB) * 4;  // CTAs resident per SM, 4 rounds for balance
  {  // FLYBY_SHARE_GPU: exactly this many persistent resolve1 CTAs per SM (4 x 96 registers x 128 threads leave room for
     // three CTAs of another context's raster_kernel; 2 / 3 / 4 / 5 measured 1.77 / 1.79 / 1.50 / 1.66 ms per mesh for two
     // alternating contexts).  FLYBY_RV_CTAS: A/B hook
    static int env_rv = -1;
    if (env_rv < 0) { const char* e = getenv("FLYBY_RV_CTAS"); env_rv = e ? atoi(e) : 0; }
    if (env_rv > 0 || rp.share) rgrid = (long long)c->sm_count * (env_rv > 0 ? env_rv : (approx ? RV_SHARE_APPROX : RV_SHARE));
  }
  const long long want = cdiv(n_pix, RV_THREADS);
  if (rgrid > want) rgrid = want;
  if (approx)
    raster_resolve1_kernel<true><<<(unsigned)rgrid, RV_THREADS, 0, st>>>(p);
  else
    raster_resolve1_kernel<false><<<(unsigned)rgrid, RV_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  if (approx) raster_resolve2_kernel<true><<<(unsigned)rgrid, RV_THREADS, 0, st>>>(p);
  else raster_resolve2_kernel<false><<<(unsigned)rgrid, RV_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  const unsigned sgrid = (unsigned)c->sm_count * 4;
  spill_t_kernel<<<sgrid, RV_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  spill_cell_kernel<<<sgrid, RV_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  if (approx) spill_shade_kernel<true><<<sgrid, RV_THREADS, 0, st>>>(p);
  else spill_shade_kernel<false><<<sgrid, RV_THREADS, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  if (approx) raster_resolve3_kernel<true><<<(unsigned)c->sm_count, 64, 0, st>>>(p);
  else raster_resolve3_kernel<false><<<(unsigned)c->sm_count, 64, 0, st>>>(p);
  FB_LAUNCH_CHECK(c);
  { int rc = slab_mark(c, 2); if (rc) return rc; }
  c->last_front_pix = n_pix;
  if (c->side_stream && c->ev_front && !c->capturing) {
    // the next render's initialisation, off the critical path: the side stream runs it beside the upload / unit
    // surface / sort of the next mesh (or right away, when the next render follows at once)
    FB_CUDA(c, cudaEventRecord(c->ev_front_free, st));
    FB_CUDA(c, cudaStreamWaitEvent(c->side_stream, c->ev_front_free, 0));
    const int64_t n_init = clean > n_pix ? clean : n_pix;  // (pixels beyond n_pix are still clean)
    raster_init_kernel<<<(unsigned)cdiv(n_pix, 256), 256, 0, c->side_stream>>>(p.front, p.cnt, n_pix);
    FB_LAUNCH_CHECK(c);
    FB_CUDA(c, cudaEventRecord(c->ev_front, c->side_stream));
    c->ev_front_live = true;
    c->front_clean = n_init;
  }
  return FLYBY_OK;
}

// A captured step cannot lean on what happened before it: the front / count words its render will use are
// initialised at its start, on the side stream (a fork of the captured graph) beside upload / unit surface / sort,
// and the render joins it through ev_front as usual.  The post-render initialisation is left out while capturing.
int raster_capture_prologue(flyby_ctx* c) {
  c->front_clean = 0;
  const int64_t n = c->last_front_pix;
  if (n <= 0 || c->raster_tmp.cap < 12 * (size_t)n) return FLYBY_OK;  // the render initialises on its own stream
  const size_t cap_pix = c->raster_tmp.cap / 12;
  unsigned long long* front = c->raster_tmp.as<unsigned long long>();
  int* cnt = reinterpret_cast<int*>(front + cap_pix);
  cudaStream_t s = c->stream;
  if (c->side_stream && c->ev_front) {
    FB_CUDA(c, cudaEventRecord(c->ev_front_free, c->stream));
    FB_SIDE_WAIT(c, c->ev_front_free);
    s = c->side_stream;
  }
  raster_init_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(front, cnt, n);
  FB_LAUNCH_CHECK(c);
  if (s != c->stream) {
    FB_CUDA(c, cudaEventRecord(c->ev_front, s));
    c->cap_front_forked = true;
  }
  c->front_clean = n;
  return FLYBY_OK;
}

}  // namespace fb
