// trace.cu -- view setup, ray generation, LBVH traversal, shading and tensor write.
//
// Replaces the hot loops of the reference's ray-cast engine
// (src/app/fly_by_features.cxx:652-852: per view, per plane point,
// vtkCellLocator::IntersectWithLine + barycentric normal + depth) and the
// channel assembly of FlyByGenerator.getFlyBy (src/py/fly_by_features.py:124-150).
//
// render_kernel is a persistent kernel: one CTA slot per SM x occupancy, each
// warp pulls 8x4-pixel tiles from a global counter.  Rays that miss the mesh
// bounds are written as background at once; the survivors are compacted into a
// per-warp queue with warp votes so that every traversal runs 32 live rays.
// The top of the LBVH (breadth-first prefix, <= 1024 nodes = 64 KB) is staged in
// shared memory with one bulk async copy (cp.async.bulk + mbarrier).
// Traversal is stackless (flyby_core.cuh); the triangle decision and the output
// values are exact double arithmetic.
//
// Roofline: algorithmic bytes per ray = 4*C + 4 (+8 with t); the
// kernel is bound by L1/L2 node-fetch latency, not HBM.
#include "trace_common.cuh"

namespace fb {

// ---------------------------------------------------------------- view setup
__global__ void view_setup_kernel(const float* __restrict__ pts, int n, double radius, double plane_scale,
                                  double spacing, double tol, int res, const float* __restrict__ ubox, View* __restrict__ views,
                                  ViewF* __restrict__ viewsf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  View v;
  float sp[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  view_setup(sp, radius, plane_scale, spacing, &v, tol);
  views[i] = v;
  // largest |coordinate| of the unit surface bounds the float32 screening errors
  float max_abs_coord = 0.f;
  for (int k = 0; k < 6; k++) max_abs_coord = fmaxf(max_abs_coord, fabsf(ubox[k]));
  ViewF f;
  viewf_setup(v, radius, spacing, max_abs_coord, tol, &f);
  viewf_pixels(v, spacing, res, &f);
  viewsf[i] = f;
}

__global__ void tc_table_kernel(int res, double* __restrict__ tc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < res) tc[i] = (res > 1) ? FB_DIV((double)i, (double)(res - 1)) : 0.0;  // same expression as view_ray
}

int views_setup(flyby_ctx* c, const flyby_render_opts* o, int res) {
  if (c->n_views <= 0) return fail(c, FLYBY_ERR_INVALID, "no viewpoints set");
  FB_CUDA(c, c->views.reserve(sizeof(View) * (size_t)c->n_views));
  FB_CUDA(c, c->viewsf.reserve(sizeof(ViewF) * (size_t)c->n_views));
  view_setup_kernel<<<(unsigned)cdiv(c->n_views, 64), 64, 0, c->stream>>>(
      c->view_pts.as<float>(), c->n_views, c->radius, o->plane_scale, o->plane_spacing, o->tolerance, res, prep_unit_box(c),
      c->views.as<View>(), c->viewsf.as<ViewF>());
  FB_LAUNCH_CHECK(c);
  FB_CUDA(c, c->tc_table.reserve(sizeof(double) * (size_t)res));
  tc_table_kernel<<<(unsigned)cdiv(res, 256), 256, 0, c->stream>>>(res, c->tc_table.as<double>());
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// ---------------------------------------------------------------- raster-compatible (pinhole) views
// the images FlyByGenerator.getFlyBy reads back from OpenGL (src/py/fly_by_features.py:85-152)
// -- perspective camera at the viewpoint, one ray through every pixel centre, depth along the viewing
// direction, 8-bit colours with normal_encoding 3 -- produced by the ray caster.  Rays are not parallel, so
// this goes through the per-ray stackless LBVH traversal (as flyby_cast), not the scan conversion.
__global__ void __launch_bounds__(128) persp_kernel(const Node* __restrict__ nodes, const float4* __restrict__ tv0,
                                                    const float4* __restrict__ tv1, const float4* __restrict__ tv2,
                                                    const int4* __restrict__ tvid, const float4* __restrict__ normals,
                                                    int has_tris, const float* __restrict__ view_pts, int view_begin,
                                                    int64_t n_pix, int res, double tan_half, double z_near, double z_far,
                                                    int use_z, int split_z, int enc, int channels, double zscale,
                                                    double tol2, float* feat, int32_t* ids, double* depth) {
  const int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (pix >= n_pix) return;
  const int64_t per_view = (int64_t)res * res;
  const int v = (int)(pix / per_view), loc = (int)(pix % per_view);
  Camera cam;
  camera_setup(view_pts + 3 * (int64_t)(view_begin + v), &cam);
  double o[3], d[3];
  camera_ray(cam, res, loc % res, loc / res, tan_half, z_near, z_far, o, d);
  float px[4] = {0.f, 0.f, 0.f, 0.f};
  int cell = -1;
  double z = -1.0;
  if (has_tris) {
    float dirf[3], invf[3];
    double amax = 1.0;
    for (int k = 0; k < 3; k++) {
      dirf[k] = (float)d[k];
      invf[k] = 1.0f / dirf[k];
      const double reach = fabs(o[k]) + fabs(d[k]);
      amax = reach > amax ? reach : amax;
    }
    const float slack = (float)(amax * 64.0 * 5.9604644775390625e-08 + 2.0 * sqrt(tol2));  // + the test's own tolerance
    const double dl = norm3(d);
    const float tslack = (float)(dl > 0.0 ? 4.0 * (double)slack / dl : 0.0);
    GlobalFetch f{nodes};
    NoCount nc;
    const Hit h = traverse(f, tv0, tv1, tv2, o, d, dirf, invf, slack, tslack, tol2, nc);
    if (h.slot >= 0) {
      const float4 a = tv0[h.slot], b = tv1[h.slot], e = tv2[h.slot];
      const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
      double t, x[3], w[3];
      tri_exact(o, d, P0, P1, P2, DBL_MAX, tol2, &t, x, w);
      cell = h.cell;
      z = FB_ADD(z_near, FB_MUL(t, FB_SUB(z_far, z_near)));
      if (feat) {
        const int4 vid = tvid[h.slot];
        const float4 n0 = __ldg(normals + vid.x), n1 = __ldg(normals + vid.y), n2 = __ldg(normals + vid.z);
        const float N0[3] = {n0.x, n0.y, n0.z}, N1[3] = {n1.x, n1.y, n1.z}, N2[3] = {n2.x, n2.y, n2.z};
        shade_z(w, N0, N1, N2, FB_MUL(z, zscale), use_z, split_z, enc, px);
      }
    }
  }
  if (feat)
    for (int k = 0; k < channels; k++) feat[pix * channels + k] = px[k];
  if (ids) ids[pix] = cell;
  if (depth) depth[pix] = z;
}

int trace_render_perspective(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, double tan_half,
                             double z_near, double z_far, float* feat, int32_t* ids, double* depth) {
  MeshDev& m = c->mesh;
  const int64_t n_pix = (int64_t)(ve - vb) * res * res;
  if (n_pix == 0) return FLYBY_OK;
  const int channels = (o->use_z && o->split_z) ? 4 : 3;
  { int rc = wait_tree(c); if (rc) return rc; }
  persp_kernel<<<(unsigned)cdiv(n_pix, 128), 128, 0, c->stream>>>(
      m.nodes.as<Node>(), m.tv0.as<float4>(), m.tv1.as<float4>(), m.tv2.as<float4>(), m.tvid.as<int4>(),
      m.normals.as<float4>(), m.T > 0 ? 1 : 0, c->view_pts.as<float>(), vb, n_pix, res, tan_half, z_near, z_far,
      o->use_z, o->split_z, o->normal_encoding, channels, o->z_scale, o->tolerance * o->tolerance, feat, ids, depth);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// ---------------------------------------------------------------- barycentric lookup at the hits
// out[pixel, f] = sum_k w_k feats[point_k(cell), f] with the exact test's weights of the cell that
// flyby_render reported for the pixel (fly_by_features.cxx:766-790 interpolates its normals this way;
// the vertex-colour shading of challenge-teeth/teeth_nets.py:121-153 is the same lookup).  The weights are
// recomputed with tri_exact on the same inputs, i.e. they are the values the render used.
__global__ void interpolate_kernel(RenderParams p, const float* __restrict__ verts, const int32_t* __restrict__ tris,
                                   const int32_t* __restrict__ cell_ids, const float* __restrict__ feats, int F,
                                   float zero, float* __restrict__ out, int64_t n_pix, int64_t T) {
  const int64_t pix = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (pix >= n_pix) return;
  const int32_t cell = cell_ids[pix];
  float* o = out + pix * F;
  if (cell < 0 || cell >= T) {  // miss, or an id that does not belong to this mesh: never reaches a gather
    for (int f = 0; f < F; f++) o[f] = zero;
    return;
  }
  const int32_t i0 = tris[3 * (int64_t)cell], i1 = tris[3 * (int64_t)cell + 1], i2 = tris[3 * (int64_t)cell + 2];
  const float P0[3] = {verts[3 * (int64_t)i0], verts[3 * (int64_t)i0 + 1], verts[3 * (int64_t)i0 + 2]};
  const float P1[3] = {verts[3 * (int64_t)i1], verts[3 * (int64_t)i1 + 1], verts[3 * (int64_t)i1 + 2]};
  const float P2[3] = {verts[3 * (int64_t)i2], verts[3 * (int64_t)i2 + 1], verts[3 * (int64_t)i2 + 2]};
  const PixelRay r = pixel_ray(p, pix);
  double t, x[3], w[3];
  if (!tri_exact(r.o, r.d, P0, P1, P2, DBL_MAX, p.tol2, &t, x, w)) {  // not a hit of this ray: ids from another set-up
    for (int f = 0; f < F; f++) o[f] = zero;
    return;
  }
  for (int f = 0; f < F; f++) {
    const double a = FB_MUL(w[0], (double)feats[(int64_t)i0 * F + f]);
    const double b = FB_MUL(w[1], (double)feats[(int64_t)i1 * F + f]);
    const double e = FB_MUL(w[2], (double)feats[(int64_t)i2 * F + f]);
    o[f] = (float)FB_ADD(FB_ADD(a, b), e);
  }
}

int trace_interpolate(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, const int32_t* cell_ids,
                      const float* feats, int F, float zero, float* out) {
  MeshDev& m = c->mesh;
  const int64_t n_pix = (int64_t)(ve - vb) * res * res;
  if (n_pix >= (int64_t)1 << 32) return fail(c, FLYBY_ERR_INVALID, "more than 2^32 rays in one call");
  if (n_pix == 0) return FLYBY_OK;
  RenderParams p;
  memset(&p, 0, sizeof p);
  p.tc = c->tc_table.as<double>();
  p.views = c->views.as<View>();
  p.view_begin = vb; p.n_views = ve - vb; p.res = res;
  p.div_view = make_fastdiv((uint32_t)res * (uint32_t)res);
  p.div_res = make_fastdiv((uint32_t)res);
  p.spacing = o->plane_spacing;
  p.tol2 = o->tolerance * o->tolerance;
  interpolate_kernel<<<(unsigned)cdiv(n_pix, 128), 128, 0, c->stream>>>(p, m.verts.as<float>(), m.tris.as<int32_t>(),
                                                                          cell_ids, feats, F, zero, out, n_pix, m.T);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// ---------------------------------------------------------------- ray dump (parity of a5)
__global__ void rays_kernel(const View* __restrict__ views, int view_begin, int n_views, int res,
                            double spacing, double* __restrict__ out) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t per = (int64_t)res * res;
  if (p >= per * n_views) return;
  int v = (int)(p / per);
  int pix = (int)(p % per);
  const View& vw = views[view_begin + v];
  double o[3];
  view_ray(vw, res, pix % res, pix / res, spacing, o);
  double* r = out + 6 * p;
  for (int k = 0; k < 3; k++) {
    r[k] = o[k];
    r[3 + k] = FB_ADD(o[k], vw.dir[k]);
  }
}

int trace_rays(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, double* rays) {
  int64_t n = (int64_t)(ve - vb) * res * res;
  rays_kernel<<<(unsigned)cdiv(n, 256), 256, 0, c->stream>>>(c->views.as<View>(), vb, ve - vb, res,
                                                             o->plane_spacing, rays);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// Warp-cooperative packet traversal.  A warp owns a batch of up to 32 compacted
// rays of ONE view (parallel rays from neighbouring pixels) and walks the LBVH as
// a unit: node index and bit-trail are warp-uniform, every lane tests its own ray
// against the two child boxes and warp votes decide where to go -- descend if ANY
// lane hits, near child first by majority.  Control flow never diverges during
// traversal, node fetches are broadcast loads, and climbing back through the
// parent / sibling links costs one lane's worth of work instead of 32.  A small
// per-warp ring in shared memory remembers recent far siblings so that most pops
// need no link walk (the trail + links stay authoritative: stackless).
// Each lane still decides its own hit exactly (double precision) and keeps its
// own cull distance, so the result equals the per-ray traversal bit for bit.
#define RK_RING 16
#define RK_MAX_STEPS (1u << 24)  // node visits of one packet; a tree has < 2^31 nodes, a packet visits far fewer

template <bool USE_TOP, bool COUNT>
__global__ void __launch_bounds__(RK_THREADS, RK_MINB) render_kernel(const RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Node* top = reinterpret_cast<Node*>(smem_raw);
  uint32_t* queues = reinterpret_cast<uint32_t*>(smem_raw + (size_t)p.top_cap * sizeof(Node));
  __shared__ float s_ubox[6];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int32_t s_ring[RK_WARPS][RK_RING];

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
  unsigned long long n_nodes = 0, n_tris = 0;
  unsigned int n_hits = 0;
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
        if (!live) {
          const float zero[4] = {0.f, 0.f, 0.f, 0.f};
          write_pixel(p, (int64_t)npid, zero, -1, -1.0);
        }
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

    // ---- per-lane ray, warp-uniform view constants
    const View& vw = p.views[p.view_begin + view];
    const double d[3] = {vw.dir[0], vw.dir[1], vw.dir[2]};
    const float slack = vw.slack, tslack = vw.tslack;
    const float dfx = vw.dirf[0], dfy = vw.dirf[1], dfz = vw.dirf[2];
    const bool negx = dfx < 0.f, negy = dfy < 0.f, negz = dfz < 0.f;
    double o[3] = {0.0, 0.0, 0.0};
    SlabRay sr;
    {
      const uint32_t pix = pid % per_view;
      if (active) view_ray(vw, res, (int)(pix % (uint32_t)res), (int)(pix / (uint32_t)res), p.spacing, o);
      float of[3] = {(float)o[0], (float)o[1], (float)o[2]};
      float inv[3] = {clamp_inv(dfx), clamp_inv(dfy), clamp_inv(dfz)};
      slab_ray_setup(of, inv, slack, &sr);
    }
    double best_t = DBL_MAX;
    int best_slot = -1, best_cell = 0x7fffffff;
    float tcull = active ? 1.0f + tslack : -1.0f;  // an idle lane never hits a box

    // exact nearest hit of this lane so far (double), with its barycentric weights
    double bw0 = 0.0, bw1 = 0.0, bw2 = 0.0;
    auto exact_test = [&](int slot) {
      const float4 a = __ldg(p.tv0 + slot), b = __ldg(p.tv1 + slot), e = __ldg(p.tv2 + slot);
      if (COUNT) n_tris++;
      const float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
      double t, x[3], w[3];
      if (tri_exact(o, d, P0, P1, P2, best_t, p.tol2, &t, x, w)) {
        const int cell = __float_as_int(a.w);
        if (t < best_t || (t == best_t && cell < best_cell)) {
          best_t = t;
          best_slot = slot;
          best_cell = cell;
          bw0 = w[0]; bw1 = w[1]; bw2 = w[2];
          return true;
        }
      }
      return false;
    };

    // ---- packet traversal (everything but the leaf tests is warp-uniform)
    int node = 0;
    unsigned long long trail = 1;
    int ring_head = 0, ring_count = 0;
    unsigned int steps = 0;
    for (;;) {
      const Node N = tf.node(node);
      float tn0, tn1;
      const bool h0 = slab_fma(N.box, sr, negx, negy, negz, tcull, &tn0);
      const bool h1 = slab_fma(N.box + 6, sr, negx, negy, negz, tcull, &tn1);
      const unsigned m0 = __ballot_sync(0xffffffffu, h0);
      const unsigned m1 = __ballot_sync(0xffffffffu, h1);
      const int c0 = N.c0, c1 = N.c1;
      if (COUNT) n_nodes += __popc(m0 | m1);
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const unsigned m = side ? m1 : m0;
        const int c = side ? c1 : c0;
        if (m != 0u && c < 0) {
          const bool h = side ? h1 : h0;
          const int first = leaf_first(c), last = first + leaf_count(c);
          // every lane whose ray enters the leaf box runs the exact test on the leaf's triangles
          for (int slot = first; slot < last; slot++)
            if (h && exact_test(slot)) tcull = fminf(tcull, FB_F2F_RU(best_t) + tslack);
        }
      }
      // internal children still wanted by at least one lane (cull distances may have shrunk)
      const bool w0 = h0 && c0 >= 0 && tn0 <= tcull;
      const bool w1 = h1 && c1 >= 0 && tn1 <= tcull;
      const unsigned d0 = __ballot_sync(0xffffffffu, w0);
      const unsigned d1 = __ballot_sync(0xffffffffu, w1);
      if (d0 != 0u || d1 != 0u) {
        const bool both = d0 != 0u && d1 != 0u;
        // near child first: majority of the lanes that want child 1 earlier (or only child 1)
        const unsigned pref1 = __ballot_sync(0xffffffffu, w1 && (!w0 || tn1 < tn0));
        const bool swap = both && 2 * __popc(pref1) > __popc(d0 | d1);
        trail = (trail << 1) | (both ? 1ull : 0ull);
        if (both) {
          ring_head = (ring_head + 1) & (RK_RING - 1);
          if (lane == 0) ring[ring_head] = swap ? c0 : c1;
          ring_count = ring_count < RK_RING ? ring_count + 1 : RK_RING;
        }
        node = both ? (swap ? c1 : c0) : (d0 != 0u ? c0 : c1);
        continue;
      }
      // dead end: climb to the nearest level with a pending far sibling
      if (++steps > RK_MAX_STEPS) {  // watchdog: a traversal bug must end as an error, not a hung GPU
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
        ring_head = (ring_head - 1) & (RK_RING - 1);
        ring_count--;
      } else {
        int nd = node;
        for (int i = 0; i < k; i++) nd = tf.parent(nd);
        node = tf.sibling(nd);
      }
    }

    // ---- shade + write (x is recomputed from t exactly as the test computed it)
    if (active) {
      float px[4] = {0.f, 0.f, 0.f, 0.f};
      double t = -1.0;
      if (best_slot >= 0) {
        n_hits++;
        t = best_t;
        if (p.feat) {
          const double x[3] = {FB_ADD(o[0], FB_MUL(t, d[0])), FB_ADD(o[1], FB_MUL(t, d[1])), FB_ADD(o[2], FB_MUL(t, d[2]))};
          const double w[3] = {bw0, bw1, bw2};
          const int4 vid = p.tvid[best_slot];
          const float4 n0 = __ldg(p.normals + vid.x), n1 = __ldg(p.normals + vid.y), n2 = __ldg(p.normals + vid.z);
          const float N0[3] = {n0.x, n0.y, n0.z}, N1[3] = {n1.x, n1.y, n1.z}, N2[3] = {n2.x, n2.y, n2.z};
          shade(o, x, w, N0, N1, N2, p.use_z, p.split_z, p.enc, p.zscale, px);
        }
      }
      write_pixel(p, (int64_t)pid, px, best_slot >= 0 ? best_cell : -1, t);
      if (best_slot >= 0 && p.n_scal) {
        const double w[3] = {bw0, bw1, bw2};
        const int4 vid = p.tvid[best_slot];
        write_hit_scalars(p, (int64_t)pid, w, vid.x, vid.y, vid.z, best_cell);
      }
    }
    __syncwarp();
  }
  if (COUNT) {
    for (int o2 = 16; o2 > 0; o2 >>= 1) n_tris += __shfl_xor_sync(0xffffffffu, n_tris, o2);
    if (lane == 0) {
      atomicAdd(p.counters + 1, n_nodes);
      atomicAdd(p.counters + 2, n_tris);
    }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) n_hits += __shfl_xor_sync(0xffffffffu, n_hits, o2);
  if (lane == 0 && n_hits) atomicAdd(p.counters + 3, (unsigned long long)n_hits);
}

__global__ void fill_miss_kernel(float* feat, int32_t* ids, double* tt, int64_t n_pix, int channels, int n_scal,
                                 float scal_zero) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  if (feat) {
    float* o = feat + i * (channels + n_scal);
    for (int k = 0; k < channels; k++) o[k] = 0.f;
    for (int k = 0; k < n_scal; k++) o[channels + k] = scal_zero;
  }
  if (ids) ids[i] = -1;
  if (tt) tt[i] = -1.0;
}

static int g_screen = 1;  // FLYBY_NO_SCREEN=1: fused kernel, exact test on every leaf candidate (A/B, parity checks)
int screen_render(flyby_ctx* c, const RenderParams& p);  // screen.cu: LBVH trace kernel + exact resolve kernel
int raster_render(flyby_ctx* c, const RenderParams& p);  // raster.cu: scan-conversion passes + exact resolve kernel
static int g_mode = 2;  // what flyby_render_opts.mode 0 means; FLYBY_MODE (A/B runs): 2 raster, 1 bvh, 0 fused
static int g_use_top = 1;  // FLYBY_NO_SMEM_TOP=1 disables the shared-memory top (A/B measurements)

// ---- flyby_render_packed: float feature tensor of one view chunk -> uint8 colours + float32 depth
// 4 pixels per thread when the chunk is a multiple of 4 pixels (12 colour bytes = three aligned words), else one
template <int C>
__global__ void __launch_bounds__(256) pack_rgb8_kernel(const float* __restrict__ feat, int64_t n_pix, uint8_t* __restrict__ rgb,
                                                         float* __restrict__ depth, int vec) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (vec) {
    if (4 * i >= n_pix) return;
    uint32_t b[12];
    float z[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float r, g, bl;
      if (C == 4) {
        const float4 f = __ldcs(reinterpret_cast<const float4*>(feat) + 4 * i + k);
        r = f.x; g = f.y; bl = f.z; z[k] = f.w;
      } else {
        const float* f = feat + 3 * (4 * i + k);
        r = __ldcs(f); g = __ldcs(f + 1); bl = __ldcs(f + 2); z[k] = 0.f;
      }
      b[3 * k] = (uint32_t)(int)r & 255u; b[3 * k + 1] = (uint32_t)(int)g & 255u; b[3 * k + 2] = (uint32_t)(int)bl & 255u;
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(rgb) + 3 * i;
    __stcs(o, b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24));
    __stcs(o + 1, b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24));
    __stcs(o + 2, b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24));
    if (C == 4 && depth) __stcs(reinterpret_cast<float4*>(depth) + i, make_float4(z[0], z[1], z[2], z[3]));
  } else {
    if (i >= n_pix) return;
    const float* f = feat + C * i;
    rgb[3 * i] = (uint8_t)(int)f[0]; rgb[3 * i + 1] = (uint8_t)(int)f[1]; rgb[3 * i + 2] = (uint8_t)(int)f[2];
    if (C == 4 && depth) depth[i] = f[3];
  }
}

int pack_rgb8(flyby_ctx* c, const float* feat, int channels, int64_t n_pix, uint8_t* rgb, float* depth, int vec) {
  const unsigned grid = (unsigned)((vec ? n_pix / 4 : n_pix) / 256 + 1);
  if (channels == 4) pack_rgb8_kernel<4><<<grid, 256, 0, c->stream>>>(feat, n_pix, rgb, depth, vec);
  else pack_rgb8_kernel<3><<<grid, 256, 0, c->stream>>>(feat, n_pix, rgb, nullptr, vec);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

// first_chunk / last_chunk: a render call may be split into view chunks (api.cu overlaps the
// copy-back of one chunk with the render of the next); stats and events span all chunks.
int trace_render(flyby_ctx* c, int vb, int ve, int res, const flyby_render_opts* o, float* feat,
                 int32_t* ids, double* tt, bool first_chunk, bool last_chunk) {
  MeshDev& m = c->mesh;
  cudaStream_t st = c->stream;
  const int nv = ve - vb;
  const int64_t n_pix = (int64_t)nv * res * res;
  if (n_pix >= (int64_t)1 << 32) return fail(c, FLYBY_ERR_INVALID, "more than 2^32 rays in one render call");
  const int channels = (o->use_z && o->split_z) ? 4 : 3;
  if (first_chunk) {
    c->n_slab_ev = 0;
    FB_STAT_EVENT(c, c->ev[5], st);
    FB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, 7 * sizeof(unsigned long long), st));
    FB_CUDA(c, cudaMemsetAsync(c->counters.as<unsigned long long>() + 8, 0, 8 * sizeof(unsigned long long), st));
    c->stats.n_rays = 0;
  }
  c->stats.n_rays += n_pix;
  if (m.T == 0) {
    fill_miss_kernel<<<(unsigned)cdiv(n_pix, 256), 256, 0, st>>>(feat, ids, tt, n_pix, channels, m.n_scalars, m.scalar_zero);
    FB_LAUNCH_CHECK(c);
    if (last_chunk) FB_STAT_EVENT(c, c->ev[6], st);
    return FLYBY_OK;
  }
  static bool env_read = false;
  if (!env_read) {
    const char* e = getenv("FLYBY_NO_SMEM_TOP");
    if (e && e[0] == '1') g_use_top = 0;
    e = getenv("FLYBY_NO_SCREEN");
    if (e && e[0] == '1') g_screen = 0;
    e = getenv("FLYBY_MODE");
    if (e && !strcmp(e, "bvh")) g_mode = 1;
    if (e && !strcmp(e, "fused")) g_mode = 0;
    if (!g_screen) g_mode = 0;
    env_read = true;
  }
  RenderParams p;
  p.nodes = m.nodes.as<Node>();
  p.n_top_dev = m.n_top_dev;
  p.tv0 = m.tv0.as<float4>(); p.tv1 = m.tv1.as<float4>(); p.tv2 = m.tv2.as<float4>();
  p.tvid = m.tvid.as<int4>();
  p.trec = m.trec.as<float4>();
  p.tc = c->tc_table.as<double>();
  p.normals = m.normals.as<float4>();
  p.views = c->views.as<View>();
  p.ubox = prep_unit_box(c);
  p.view_begin = vb; p.n_views = nv; p.res = res;
  p.div_view = make_fastdiv((uint32_t)res * (uint32_t)res);  // (res <= 16384: res * res < 2^32)
  p.div_res = make_fastdiv((uint32_t)res);
  p.use_z = o->use_z; p.split_z = o->split_z; p.enc = o->normal_encoding; p.channels = channels;
  p.zscale = o->z_scale; p.spacing = o->plane_spacing;
  p.tol2 = o->tolerance * o->tolerance;
  p.feat = feat; p.ids = ids; p.tt = tt;
  p.counters = c->counters.as<unsigned long long>();
  int tx = (int)cdiv(res, 8), ty = (int)cdiv(res, 4);
  p.tiles_x = (tx + 1) & ~1;
  p.tiles_y = (ty + 3) & ~3;
  p.n_tiles = (long long)p.tiles_x * p.tiles_y * nv;
  p.top_cap = g_use_top ? 1024 : 0;
  p.viewsf = c->viewsf.as<ViewF>();
  p.scal = m.n_scalars > 0 ? m.scalars.as<float>() : nullptr;
  p.n_scal = m.n_scalars; p.scal_mode = m.scalar_mode; p.scal_zero = m.scalar_zero;
  p.stride = channels + m.n_scalars;
  p.flip = m.n_flipped > 0 ? m.flip.as<uint8_t>() : nullptr;
  // flyby_render_opts.mode: 0 default (g_mode: raster unless FLYBY_MODE says otherwise), 1 raster, 2 bvh, 3 fused
  int mode = g_mode;
  const int engine = o->mode & 0xff;
  if (engine == 1) mode = 2;
  else if (engine == 2) mode = 1;
  else if (engine == 3) mode = 0;
  else if (engine != 0) return fail(c, FLYBY_ERR_INVALID, "render: unknown mode %d", o->mode);
  if (o->mode & ~(0xff | FLYBY_SHADE_TOLERANCE | FLYBY_SHARE_GPU)) return fail(c, FLYBY_ERR_INVALID, "render: unknown mode flags %#x", o->mode);
  p.approx = (o->mode & FLYBY_SHADE_TOLERANCE) ? 1 : 0;  // honoured by the raster engine
  p.share = (o->mode & FLYBY_SHARE_GPU) ? 1 : 0;
  FB_CUDA(c, cudaMemsetAsync(c->counters.p, 0, sizeof(unsigned long long), st));  // tile counter
  const size_t smem = (size_t)p.top_cap * sizeof(Node) + RK_WARPS * RK_QCAP * sizeof(uint32_t);
  auto launch = [&](auto kern) -> int {
    FB_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    FB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RK_THREADS, smem));
    if (occ < 1) occ = 1;
    long long want = cdiv(p.n_tiles, RK_WARPS);
    long long grid = (long long)c->sm_count * occ;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, RK_THREADS, smem, st>>>(p);
    FB_LAUNCH_CHECK(c);
    return FLYBY_OK;
  };
  int rc;
  // the scan conversion needs a regular pixel grid with a positive pitch
  const bool grid_ok = res > 1 && o->plane_scale * o->plane_spacing > 0.0;
  if (mode == 2 && grid_ok) {
    // the per-pixel scratch of the scan conversion (~58 B / pixel) is sized for at most 2^26 pixels at a time:
    // long view lists are rendered in slabs of whole views, each straight into its part of the outputs
    const int64_t per_view = (int64_t)res * res;
    static int64_t slab_pix = 0;
    if (!slab_pix) {
      const char* e = getenv("FLYBY_SLAB_PIXELS");  // A/B hook
      slab_pix = e ? atoll(e) : ((int64_t)1 << 26);
      if (slab_pix < 1) slab_pix = (int64_t)1 << 26;
    }
    int slab = (int)(slab_pix / per_view);
    if (slab < 1) slab = 1;
    if (slab > 65535) slab = 65535;  // gridDim.y of raster_kernel
    rc = FLYBY_OK;
    for (int v0 = 0; v0 < nv && rc == FLYBY_OK; v0 += slab) {
      RenderParams q = p;
      q.view_begin = vb + v0;
      q.n_views = nv - v0 < slab ? nv - v0 : slab;
      const int64_t off = (int64_t)v0 * per_view;
      if (feat) q.feat = feat + off * p.stride;
      if (ids) q.ids = ids + off;
      if (tt) q.tt = tt + off;
      rc = raster_render(c, q);
    }
  } else if (mode >= 1) {
    if ((rc = wait_tree(c))) return rc;
    rc = screen_render(c, p);
  } else {
    if ((rc = wait_tree(c))) return rc;
    if ((rc = slab_mark(c, 0))) return rc;
    const int variant = (g_use_top ? 2 : 0) | (c->count_work ? 1 : 0);
    switch (variant) {
      case 0: rc = launch(render_kernel<false, false>); break;
      case 1: rc = launch(render_kernel<false, true>); break;
      case 2: rc = launch(render_kernel<true, false>); break;
      default: rc = launch(render_kernel<true, true>); break;
    }
    if (!rc) rc = slab_mark(c, 1);  // one kernel searches and resolves: all of it counts as the search
    if (!rc) rc = slab_mark(c, 2);
  }
  if (rc) return rc;
  if (last_chunk) FB_STAT_EVENT(c, c->ev[6], st);
  return FLYBY_OK;
}

// ---------------------------------------------------------------- generic segments
// vtkCellLocator::IntersectWithLine for caller-supplied segments (predict.py:128)
__global__ void cast_kernel(const Node* __restrict__ nodes, const float4* __restrict__ tv0,
                            const float4* __restrict__ tv1, const float4* __restrict__ tv2, int has_tris,
                            const double* __restrict__ rays, int64_t n, double tol2, int32_t* ids, double* tt,
                            double* ww) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  double o[3] = {rays[6 * r], rays[6 * r + 1], rays[6 * r + 2]};
  double d[3] = {FB_SUB(rays[6 * r + 3], o[0]), FB_SUB(rays[6 * r + 4], o[1]), FB_SUB(rays[6 * r + 5], o[2])};
  Hit h;
  h.slot = -1; h.cell = -1; h.t = DBL_MAX;
  double w[3] = {0, 0, 0}, t = -1.0;
  if (has_tris) {
    float dirf[3], invf[3];
    double amax = 1.0;
    for (int k = 0; k < 3; k++) {
      dirf[k] = (float)d[k];
      invf[k] = 1.0f / dirf[k];
      double reach = fabs(o[k]) + fabs(d[k]);
      amax = reach > amax ? reach : amax;
    }
    float slack = (float)(amax * 64.0 * 5.9604644775390625e-08 + 2.0 * sqrt(tol2));  // + the test's own tolerance
    double dl = norm3(d);
    float tslack = (float)(dl > 0.0 ? 4.0 * (double)slack / dl : 0.0);
    GlobalFetch f{nodes};
    NoCount nc;
    h = traverse(f, tv0, tv1, tv2, o, d, dirf, invf, slack, tslack, tol2, nc);
    if (h.slot >= 0) {
      float4 a = tv0[h.slot], b = tv1[h.slot], e = tv2[h.slot];
      float P0[3] = {a.x, a.y, a.z}, P1[3] = {b.x, b.y, b.z}, P2[3] = {e.x, e.y, e.z};
      double x[3];
      tri_exact(o, d, P0, P1, P2, DBL_MAX, tol2, &t, x, w);
    }
  }
  if (ids) ids[r] = h.cell;
  if (tt) tt[r] = t;
  if (ww) for (int k = 0; k < 3; k++) ww[3 * r + k] = w[k];
}

int trace_cast(flyby_ctx* c, const double* rays, int64_t n, double tol, int32_t* ids, double* tt, double* w) {
  MeshDev& m = c->mesh;
  if (n <= 0) return FLYBY_OK;
  { int rc = wait_tree(c); if (rc) return rc; }
  cast_kernel<<<(unsigned)cdiv(n, 128), 128, 0, c->stream>>>(
      m.nodes.as<Node>(), m.tv0.as<float4>(), m.tv1.as<float4>(), m.tv2.as<float4>(), m.T > 0 ? 1 : 0,
      rays, n, tol * tol, ids, tt, w);
  FB_LAUNCH_CHECK(c);
  return FLYBY_OK;
}

}  // namespace fb
