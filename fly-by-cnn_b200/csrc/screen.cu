// screen.cu -- the default render path, split by arithmetic type:
//
//   trace_kernel    float32 only.  Warp-cooperative stackless packet traversal of the
//                   LBVH (see trace.cu) with the float32 screening test of
//                   flyby_core.cuh at the leaves.  It decides nothing: per ray it emits
//                   the short list of triangle slots that can still be the nearest hit
//                   (every candidate the exact test could accept whose depth interval
//                   is not behind a certain hit).  No double precision in the loop keeps
//                   it at 64 registers, i.e. twice the resident warps of the fused kernel,
//                   which is what hides the dependent node-fetch latency.
//   resolve_kernel  double precision.  One thread per pixel in raster order: the exact
//                   VTK-semantics test on the 1-2 listed slots (tie-break smallest t, then
//                   lowest cell id), shading and the coalesced tensor write.  A ray whose
//                   list overflowed is re-traced exactly, per ray, right here.
//
// Reference replaced: the per-view / per-plane-point loops of
// src/app/fly_by_features.cxx:652-852 and the channel assembly of
// src/py/fly_by_features.py:124-150.
#include "trace_common.cuh"

namespace fb {

#define TK_THREADS 512
#define TK_WARPS (TK_THREADS / 32)
#define TK_CAND 4       // candidate slots kept per ray
#define TK_RING 16
#define TK_OVERFLOW (-2)  // cand.w: the list overflowed, resolve re-traces the ray exactly
#define RK_MAX_STEPS (1u << 24)  // dead ends of one packet before the watchdog gives up

struct TraceParams {
  RenderParams r;
  int4* cand;  // [n_pix] candidate slots of every ray (-1 = empty)
};

template <bool USE_TOP, bool COUNT>
__global__ void __launch_bounds__(TK_THREADS, 2) trace_kernel(const TraceParams tp) {
  const RenderParams& p = tp.r;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Node* top = reinterpret_cast<Node*>(smem_raw);
  uint32_t* queues = reinterpret_cast<uint32_t*>(smem_raw + (size_t)p.top_cap * sizeof(Node));
  __shared__ float s_ubox[6];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int32_t s_ring[TK_WARPS][TK_RING];
  __shared__ int32_t s_cslot[TK_CAND][TK_THREADS];
  __shared__ float s_clo[TK_CAND][TK_THREADS];
  __shared__ unsigned int s_steps[TK_WARPS];  // dead ends of the current packet (watchdog)
  __shared__ ViewF s_vf[TK_WARPS];  // screening frame of the view each warp is working on

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  int ntop = 0;
  if (USE_TOP) {
    ntop = *p.n_top_dev;
    if (ntop > p.top_cap) ntop = p.top_cap;
  }
  if (threadIdx.x < 6) s_ubox[threadIdx.x] = p.ubox[threadIdx.x];
  stage_top_of_tree(top, p.nodes, USE_TOP ? ntop : 0, &s_bar);

  uint32_t* q = queues + warp * RK_QCAP;
  int32_t* ring = s_ring[warp];
  int qn = 0;              // warp-uniform
  bool tiles_left = true;  // warp-uniform
  const int res = p.res;
  const uint32_t per_view = (uint32_t)res * (uint32_t)res;
  const long long tiles_per_view = (long long)p.tiles_x * p.tiles_y;
  const int groups_x = p.tiles_x >> 1;
  unsigned long long n_nodes = 0, n_tris = 0, n_visits = 0, n_packets = 0, n_lanes = 0, n_leaf = 0;
  TopFetch tf{p.nodes, top, ntop};

  for (;;) {
    // ---- refill the queue until a full warp of live rays is available (ray compaction by vote)
    while (qn < 32 && tiles_left) {
      long long tile = 0;
      if (lane == 0) tile = (long long)atomicAdd(p.counters, 1ull);
      tile = __shfl_sync(0xffffffffu, tile, 0);
      if (tile >= p.n_tiles) { tiles_left = false; break; }
      const int v = (int)(tile / tiles_per_view);
      const int tl = (int)(tile % tiles_per_view);
      // 8 consecutive tiles form a 16x16-pixel block (2 tiles wide, 4 high)
      const int g = tl >> 3, r = tl & 7;
      const int tx = (g % groups_x) * 2 + (r & 1);
      const int ty = (g / groups_x) * 4 + (r >> 1);
      const int px_ = tx * 8 + (lane & 7), py_ = ty * 4 + (lane >> 3);
      const bool valid = px_ < res && py_ < res;
      bool live = false;
      uint32_t npid = 0;
      if (valid) {
        const View& vw = p.views[p.view_begin + v];
        double o[3];
        view_ray(vw, res, px_, py_, p.spacing, o);
        live = hits_bounds(s_ubox, o, vw);
        npid = (uint32_t)v * per_view + (uint32_t)py_ * (uint32_t)res + (uint32_t)px_;
        if (!live) tp.cand[npid] = make_int4(-1, -1, -1, -1);  // background: nothing to resolve
      }
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) q[qn + __popc(m & lt_mask)] = npid;
      qn += __popc(m);
      __syncwarp();
    }
    if (qn == 0) break;
    // ---- take the leading entries of one view (at most 32)
    const int view = (int)(q[0] / per_view);
    const int look = qn < 32 ? qn : 32;
    const bool same = lane < look && (int)(q[lane] / per_view) == view;
    const unsigned msame = __ballot_sync(0xffffffffu, same);
    const int take = (msame == 0xffffffffu) ? 32 : __ffs(~msame) - 1;
    const bool active = lane < take;
    const uint32_t pid = active ? q[lane] : 0u;
    __syncwarp();
    {
      uint32_t m0 = (lane + take < qn) ? q[lane + take] : 0u;
      uint32_t m1 = (lane + 32 + take < qn) ? q[lane + 32 + take] : 0u;
      __syncwarp();
      if (lane + take < qn) q[lane] = m0;
      if (lane + 32 + take < qn) q[lane + 32] = m1;
      __syncwarp();
    }
    qn -= take;

    if (COUNT) { n_packets++; n_lanes += take; }
    // ---- per-lane ray in float32 (the double origin is only needed to derive it)
    const View& vw = p.views[p.view_begin + view];
    {
      // warp-uniform, so it lives in shared memory instead of 18 registers per thread
      const uint32_t* src = reinterpret_cast<const uint32_t*>(p.viewsf + p.view_begin + view);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&s_vf[warp]);
      if (lane < (int)(sizeof(ViewF) / 4)) dst[lane] = __ldg(src + lane);
      __syncwarp();
    }
    const ViewF& vf = s_vf[warp];
    const float tslack = vw.tslack;
    const float dfx = vw.dirf[0], dfy = vw.dirf[1], dfz = vw.dirf[2];
    const bool negx = dfx < 0.f, negy = dfy < 0.f, negz = dfz < 0.f;
    SlabRay sr;
    float qx = 0.f, qy = 0.f, zo = 0.f;
    {
      double o[3] = {0.0, 0.0, 0.0};
      const uint32_t pix = pid % per_view;
      if (active) view_ray(vw, res, (int)(pix % (uint32_t)res), (int)(pix / (uint32_t)res), p.spacing, o);
      float of[3] = {(float)o[0], (float)o[1], (float)o[2]};
      float inv[3] = {clamp_inv(dfx), clamp_inv(dfy), clamp_inv(dfz)};
      slab_ray_setup(of, inv, vw.slack, &sr);
      screen_ray(vw, o, &qx, &qy, &zo);
    }
    const float inv_two_r = 1.000001f / vf.two_r;
    float tcull = active ? 1.0f + tslack : -1.0f;  // an idle lane never hits a box
    float uz = INFINITY;                            // depth bound of the nearest certain hit
    int nc = 0;                                     // candidates listed (TK_OVERFLOW once lost)

    // ---- packet traversal: node and trail are warp-uniform
    int node = 0;
    unsigned long long trail = 1;
    int ring_head = 0, ring_count = 0;
    s_steps[warp] = 0;
    for (;;) {
      const Node N = tf.node(node);
      float tn0, tn1;
      const bool h0 = slab_fma(N.box, sr, negx, negy, negz, tcull, &tn0);
      const bool h1 = slab_fma(N.box + 6, sr, negx, negy, negz, tcull, &tn1);
      const unsigned m0 = __ballot_sync(0xffffffffu, h0);
      const unsigned m1 = __ballot_sync(0xffffffffu, h1);
      const int c0 = N.c0, c1 = N.c1;
      if (COUNT) { n_nodes += __popc(m0 | m1); n_visits++; }
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const unsigned m = side ? m1 : m0;
        const int c = side ? c1 : c0;
        if (m != 0u && c < 0) {
          if (COUNT) n_leaf++;
          const bool h = side ? h1 : h0;
          const int first = leaf_first(c), last = first + leaf_count(c);
          for (int slot = first; slot < last; slot++) {
            const float4 a = __ldg(p.tv0 + slot), b = __ldg(p.tv1 + slot), e = __ldg(p.tv2 + slot);
            if (h) {
              if (COUNT) n_tris++;
              const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
              float zlo, zhi;
              const int cls = screen_tri(vf, qx, qy, zo, P0, P1, P2, &zlo, &zhi);
              if (cls != SC_MISS && zlo <= uz && nc != TK_OVERFLOW) {
                if (cls == SC_SURE && zhi < uz) {
                  uz = zhi;
                  tcull = fminf(tcull, fmaf(uz - zo + 2.0f * vf.e_z, inv_two_r, tslack));
                }
                if (nc == TK_CAND) {
                  // list full: drop entries that a nearer certain hit has ruled out since
                  int k = 0;
                  for (int i = 0; i < TK_CAND; i++) {
                    const float lo = s_clo[i][threadIdx.x];
                    const int sl = s_cslot[i][threadIdx.x];
                    if (lo <= uz) { s_clo[k][threadIdx.x] = lo; s_cslot[k][threadIdx.x] = sl; k++; }
                  }
                  nc = k;
                }
                if (nc == TK_CAND) {
                  nc = TK_OVERFLOW;  // resolve_kernel re-traces this ray exactly
                } else {
                  s_cslot[nc][threadIdx.x] = slot;
                  s_clo[nc][threadIdx.x] = zlo;
                  nc++;
                }
              }
            }
          }
        }
      }
      // internal children still wanted by at least one lane (cull distances may have shrunk)
      const bool w0 = h0 && c0 >= 0 && tn0 <= tcull;
      const bool w1 = h1 && c1 >= 0 && tn1 <= tcull;
      const unsigned d0 = __ballot_sync(0xffffffffu, w0);
      const unsigned d1 = __ballot_sync(0xffffffffu, w1);
      if ((d0 | d1) != 0u) {
        const bool both = d0 != 0u && d1 != 0u;
        // near child first: majority of the lanes that want child 1 earlier (or only child 1)
        const unsigned pref1 = __ballot_sync(0xffffffffu, w1 && (!w0 || tn1 < tn0));
        const bool swap = both && 2 * __popc(pref1) > __popc(d0 | d1);
        trail = (trail << 1) | (both ? 1ull : 0ull);
        if (both) {
          ring_head = (ring_head + 1) & (TK_RING - 1);
          ring[ring_head] = swap ? c0 : c1;  // every lane stores the same word: no branch
          ring_count = ring_count < TK_RING ? ring_count + 1 : TK_RING;
        }
        node = both ? (swap ? c1 : c0) : (d0 != 0u ? c0 : c1);
        continue;
      }
      // dead end: climb to the nearest level with a pending far sibling
      if (++s_steps[warp] > RK_MAX_STEPS) {  // watchdog: a traversal bug must end as an error, not a hung GPU
        if (lane == 0) atomicAdd(p.counters + 6, 1ull);
        break;
      }
      const int k = __ffsll((long long)trail) - 1;
      trail >>= k;
      if (trail == 1ull) break;
      trail ^= 1ull;
      if (ring_count > 0) {
        __syncwarp();
        node = ring[ring_head];
        ring_head = (ring_head - 1) & (TK_RING - 1);
        ring_count--;
      } else {
        int nd = node;
        for (int i = 0; i < k; i++) nd = tf.parent(nd);
        node = tf.sibling(nd);
      }
    }

    // ---- emit the candidates that are not behind the nearest certain hit
    if (active) {
      int4 out = make_int4(-1, -1, -1, -1);
      if (nc == TK_OVERFLOW) {
        out.w = TK_OVERFLOW;
      } else {
        int k = 0;
        for (int i = 0; i < TK_CAND; i++) {
          if (i < nc && s_clo[i][threadIdx.x] <= uz) {
            const int sl = s_cslot[i][threadIdx.x];
            if (k == 0) out.x = sl; else if (k == 1) out.y = sl; else if (k == 2) out.z = sl; else out.w = sl;
            k++;
          }
        }
      }
      tp.cand[pid] = out;
    }
    __syncwarp();
  }
  if (COUNT) {
    for (int o2 = 16; o2 > 0; o2 >>= 1) n_tris += __shfl_xor_sync(0xffffffffu, n_tris, o2);
    if (lane == 0) {
      atomicAdd(p.counters + 1, n_nodes);
      atomicAdd(p.counters + 2, n_tris);
      // debug-only traversal statistics (count_work): warp node visits, packets, live lanes, leaf visits
      atomicAdd(p.counters + 8, n_visits);
      atomicAdd(p.counters + 9, n_packets);
      atomicAdd(p.counters + 10, n_lanes);
      atomicAdd(p.counters + 11, n_leaf);
    }
  }
}

// ---------------------------------------------------------------- resolve
#define RS_THREADS 128

// ---- classify: bin the pixels by their number of candidates so that the exact kernels run dense.
//   0 candidates -> background, written here;  1 -> list1 (front of `lists`);  >= 2 -> list2 (back).
// cnt == nullptr: candidates come from trace_kernel (empty slots are -1, cand.w == TK_OVERFLOW flags an
// overflowed list); cnt != nullptr: they come from the scan-conversion passes (cnt[pixel] valid slots,
// more than 2 * TK_CAND means overflow).
__global__ void __launch_bounds__(1024) classify_kernel(const TraceParams tp, const int* __restrict__ cnt, int64_t n_pix,
                                                       uint32_t* __restrict__ lists, unsigned int* __restrict__ n_lists) {
  const RenderParams& p = tp.r;
  const int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int n = 0;
  if (pix < n_pix) {
    if (cnt) {
      n = __ldcs(cnt + pix);
    } else {
      const int4 cd = __ldcs(tp.cand + pix);
      n = cd.w == TK_OVERFLOW ? 2 * TK_CAND + 1 : (cd.x >= 0) + (cd.y >= 0) + (cd.z >= 0) + (cd.w >= 0);
    }
    if (n == 0) {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      write_pixel(p, pix, zero, -1, -1.0);
    }
  }
  // block-aggregated appends (two atomics per block) keep the lists in raster order inside a block, so the
  // later writes stay coalesced, and keep the single pair of counters out of the critical path
  __shared__ unsigned int s_cnt[2][32], s_base[2];
  const unsigned m1 = __ballot_sync(0xffffffffu, n == 1), m2 = __ballot_sync(0xffffffffu, n >= 2);
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (lane == 0) { s_cnt[0][warp] = __popc(m1); s_cnt[1][warp] = __popc(m2); }
  __syncthreads();
  if (threadIdx.x < 2) {
    unsigned int tot = 0;
    for (int w = 0; w < nwarps; w++) { const unsigned int c = s_cnt[threadIdx.x][w]; s_cnt[threadIdx.x][w] = tot; tot += c; }
    s_base[threadIdx.x] = tot ? atomicAdd(n_lists + threadIdx.x, tot) : 0u;
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  if (n == 1) lists[s_base[0] + s_cnt[0][warp] + __popc(m1 & lt)] = (uint32_t)pix;
  if (n >= 2) lists[n_pix - 1 - (s_base[1] + s_cnt[1][warp] + __popc(m2 & lt))] = (uint32_t)pix;
}

// ---- pixels with exactly one candidate (the common case): straight-line exact test, all lanes busy
__global__ void __launch_bounds__(RS_THREADS, 4) resolve1_kernel(const TraceParams tp, const uint32_t* __restrict__ list,
                                                                  const unsigned int* __restrict__ n_list) {
  const RenderParams& p = tp.r;
  const unsigned int n = *n_list;
  unsigned int hits = 0;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t pix_id = list[k];
    const int slot = __ldcs(reinterpret_cast<const int*>(tp.cand + pix_id));
    const float4* rec = p.trec + 4 * (int64_t)slot;
    const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2), g = __ldg(rec + 3);
    const PixelRay r = pixel_ray(p, pix_id);
    const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
    double t, x[3], w[3];
    if (tri_exact(r.o, r.d, P0, P1, P2, DBL_MAX, p.tol2, &t, x, w)) {
      hits++;
      shade_and_write(p, pix_id, r, t, w, __float_as_int(b.w), __float_as_int(e.w), __float_as_int(g.x), __float_as_int(a.w));
    } else {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      write_pixel(p, pix_id, zero, -1, -1.0);
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o2);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd(p.counters + 3, (unsigned long long)hits);
}

// ---- pixels with several candidates (near edges, silhouettes, layered geometry) or an overflowed list
__global__ void __launch_bounds__(RS_THREADS, 4) resolve2_kernel(const TraceParams tp, const int* __restrict__ cnt,
                                                                  const int4* __restrict__ cand2,
                                                                  const uint32_t* __restrict__ list_end,
                                                                  const unsigned int* __restrict__ n_list) {
  const RenderParams& p = tp.r;
  const unsigned int n = *n_list;
  unsigned int hits = 0, n_over = 0;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const int64_t pix_id = *(list_end - k);
    int4 c0 = __ldcs(tp.cand + pix_id), c1 = make_int4(-1, -1, -1, -1);
    int nc;
    if (cnt) {
      nc = __ldcs(cnt + pix_id);
      if (nc > TK_CAND && nc <= 2 * TK_CAND) c1 = __ldcs(cand2 + pix_id);
    } else {
      nc = c0.w == TK_OVERFLOW ? 2 * TK_CAND + 1 : (c0.x >= 0) + (c0.y >= 0) + (c0.z >= 0) + (c0.w >= 0);
    }
    const PixelRay r = pixel_ray(p, pix_id);
    double best_t = DBL_MAX, bw[3] = {0.0, 0.0, 0.0};
    int best_slot = -1, best_cell = 0x7fffffff, bp0 = 0, bp1 = 0, bp2 = 0;
    if (nc > 2 * TK_CAND) {
      n_over++;
      const int64_t per_view = (int64_t)p.res * p.res;
      const ExactHit h = resolve_overflow(p, p.view_begin + (int)(pix_id / per_view), r.o[0], r.o[1], r.o[2]);
      best_t = h.t; best_slot = h.slot; best_cell = h.cell;
      bw[0] = h.w0; bw[1] = h.w1; bw[2] = h.w2;
      if (h.slot >= 0) {
        const int4 vid = p.tvid[h.slot];
        bp0 = vid.x; bp1 = vid.y; bp2 = vid.z;
      }
    } else {
      for (int i = 0; i < nc; i++) {
        const int4 cc = i < TK_CAND ? c0 : c1;
        const int j = i & (TK_CAND - 1);
        const int slot = j == 0 ? cc.x : (j == 1 ? cc.y : (j == 2 ? cc.z : cc.w));
        const float4* rec = p.trec + 4 * (int64_t)slot;
        const float4 a = __ldg(rec), b = __ldg(rec + 1), e = __ldg(rec + 2), g = __ldg(rec + 3);
        const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
        double tt, x[3], w[3];
        if (tri_exact(r.o, r.d, P0, P1, P2, best_t, p.tol2, &tt, x, w)) {
          const int c = __float_as_int(a.w);
          if (tt < best_t || (tt == best_t && c < best_cell)) {
            best_t = tt; best_slot = slot; best_cell = c;
            bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2];
            bp0 = __float_as_int(b.w); bp1 = __float_as_int(e.w); bp2 = __float_as_int(g.x);
          }
        }
      }
    }
    if (best_slot >= 0) {
      hits++;
      shade_and_write(p, pix_id, r, best_t, bw, bp0, bp1, bp2, best_cell);
    } else {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      write_pixel(p, pix_id, zero, -1, -1.0);
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o2);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd(p.counters + 3, (unsigned long long)hits);
  if (n_over) atomicAdd(p.counters + 5, (unsigned long long)n_over);  // rays re-traced exactly (rare)
}

int resolve_launch(flyby_ctx* c, const RenderParams& rp, const int4* cand, const int* cnt, const int4* cand2);

int screen_render(flyby_ctx* c, const RenderParams& rp) {
  cudaStream_t st = c->stream;
  const int64_t n_pix = (int64_t)rp.n_views * rp.res * rp.res;
  FB_CUDA(c, c->cand.reserve(sizeof(int4) * (size_t)n_pix));
  TraceParams tp;
  tp.r = rp;
  tp.cand = c->cand.as<int4>();
  const size_t smem = (size_t)rp.top_cap * sizeof(Node) + TK_WARPS * RK_QCAP * sizeof(uint32_t);
  auto launch = [&](auto kern) -> int {
    FB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    FB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TK_THREADS, smem));
    if (occ < 1) occ = 1;
    long long want = cdiv(rp.n_tiles, TK_WARPS);
    long long grid = (long long)c->sm_count * occ;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, TK_THREADS, smem, st>>>(tp);
    FB_LAUNCH_CHECK(c);
    return FLYBY_OK;
  };
  int rc;
  if ((rc = slab_mark(c, 0))) return rc;
  const int variant = (rp.top_cap > 0 ? 2 : 0) | (c->count_work ? 1 : 0);
  switch (variant) {
    case 0: rc = launch(trace_kernel<false, false>); break;
    case 1: rc = launch(trace_kernel<false, true>); break;
    case 2: rc = launch(trace_kernel<true, false>); break;
    default: rc = launch(trace_kernel<true, true>); break;
  }
  if (rc) return rc;
  if ((rc = slab_mark(c, 1))) return rc;  // between the two kernels
  if ((rc = resolve_launch(c, rp, c->cand.as<int4>(), nullptr, nullptr))) return rc;
  return slab_mark(c, 2);
}

int resolve_launch(flyby_ctx* c, const RenderParams& rp, const int4* cand, const int* cnt, const int4* cand2) {
  const int64_t n_pix = (int64_t)rp.n_views * rp.res * rp.res;
  cudaStream_t st = c->stream;
  TraceParams tp;
  tp.r = rp;
  tp.cand = const_cast<int4*>(cand);
  FB_CUDA(c, c->pix_lists.reserve(4 * (size_t)n_pix + 16));
  uint32_t* lists = c->pix_lists.as<uint32_t>();
  unsigned int* n_lists = reinterpret_cast<unsigned int*>(c->counters.as<unsigned long long>() + 14);
  FB_CUDA(c, cudaMemsetAsync(n_lists, 0, 8, st));
  classify_kernel<<<(unsigned)cdiv(n_pix, 1024), 1024, 0, st>>>(tp, cnt, n_pix, lists, n_lists);
  FB_LAUNCH_CHECK(c);
  long long grid = (long long)c->sm_count * 4 * 4;  // 4 CTAs resident per SM, 4 rounds for balance
  const long long want = cdiv(n_pix, RS_THREADS);
  if (grid > want) grid = want;
  resolve1_kernel<<<(unsigned)grid, RS_THREADS, 0, st>>>(tp, lists, n_lists);
  FB_LAUNCH_CHECK(c);
  resolve2_kernel<<<(unsigned)grid, RS_THREADS, 0, st>>>(tp, cnt, cand2, lists + n_pix - 1, n_lists + 1);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

}  // namespace fb
