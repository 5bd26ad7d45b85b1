// hostsim.cpp -- TEST-ONLY host build of fly-by-cnn_b200/csrc/flyby_core.cuh.
//
// Compiles the very same per-element device logic (Morton codes, Karras
// topology, conservative slab test, exact triangle test, stackless traversal,
// shading) with g++ and drives it with serial loops, so that logic errors are
// found on the CPU-only build box.  Never loaded by the product package.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../fly-by-cnn_b200/csrc/flyby_core.cuh"

using namespace fb;

struct SimFetch {
  const Node* n;
  Node node(int i) const { return n[i]; }
  int parent(int i) const { return n[i].parent; }
  int sibling(int i) const { return n[i].sibling; }
};
struct SimCount {
  long long nodes = 0, tris = 0;
  void node() { nodes++; }
  void tri() { tris++; }
};

static int g_leaf_max = FB_LEAF_MAX;
extern "C" __attribute__((visibility("default"))) void sim_set_leaf_max(int k) { g_leaf_max = k; }

struct SimBvh {
  std::vector<Node> nodes;
  std::vector<float4> tv0, tv1, tv2;
  std::vector<int32_t> vid;  // 4 per slot
  int max_depth = 0;
};

static void box_of(const SimBvh& b, const std::vector<float>& leaf_box, const std::vector<float>& node_box,
                   int child, float* out) {
  const float* s = child < 0 ? &leaf_box[6 * (size_t)(~child)] : &node_box[6 * (size_t)child];
  memcpy(out, s, 24);
  (void)b;
}

static void build(const float* v, int64_t V, const int32_t* tris, int64_t T, SimBvh& b) {
  float ubox[6] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = 0; i < V; i++)
    for (int k = 0; k < 3; k++) {
      ubox[k] = fminf(ubox[k], v[3 * i + k]);
      ubox[3 + k] = fmaxf(ubox[3 + k], v[3 * i + k]);
    }
  std::vector<uint32_t> codes(T), idx(T);
  for (int64_t c = 0; c < T; c++) {
    float q[3];
    for (int k = 0; k < 3; k++) {
      float a = v[3 * tris[3 * c] + k], bb = v[3 * tris[3 * c + 1] + k], e = v[3 * tris[3 * c + 2] + k];
      float mn = fminf(a, fminf(bb, e)), mx = fmaxf(a, fmaxf(bb, e));
      float ext = ubox[3 + k] - ubox[k];
      q[k] = ext > 0.f ? (0.5f * (mn + mx) - ubox[k]) / ext : 0.5f;
    }
    codes[c] = morton30(q[0], q[1], q[2]);
    idx[c] = (uint32_t)c;
  }
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t c) { return codes[a] < codes[c]; });
  std::vector<uint32_t> sc(T);
  for (int64_t s = 0; s < T; s++) sc[s] = codes[idx[s]];
  b.tv0.resize(T); b.tv1.resize(T); b.tv2.resize(T); b.vid.resize(4 * T);
  std::vector<float> leaf_box(6 * T);
  for (int64_t s = 0; s < T; s++) {
    int cell = (int)idx[s];
    const float* p0 = v + 3 * tris[3 * cell];
    const float* p1 = v + 3 * tris[3 * cell + 1];
    const float* p2 = v + 3 * tris[3 * cell + 2];
    for (int k = 0; k < 3; k++) {
      leaf_box[6 * s + k] = fminf(p0[k], fminf(p1[k], p2[k]));
      leaf_box[6 * s + 3 + k] = fmaxf(p0[k], fmaxf(p1[k], p2[k]));
    }
    float w;
    memcpy(&w, &cell, 4);
    b.tv0[s] = {p0[0], p0[1], p0[2], w};
    b.tv1[s] = {p1[0], p1[1], p1[2], 0.f};
    b.tv2[s] = {p2[0], p2[1], p2[2], 0.f};
    b.vid[4 * s] = tris[3 * cell]; b.vid[4 * s + 1] = tris[3 * cell + 1];
    b.vid[4 * s + 2] = tris[3 * cell + 2]; b.vid[4 * s + 3] = cell;
  }
  int n = (int)T;
  if (n == 1) {
    Node nd;
    for (int k = 0; k < 6; k++) { nd.box[k] = leaf_box[k]; nd.box[6 + k] = leaf_box[k]; }
    nd.c0 = ~0; nd.c1 = ~0; nd.parent = -1; nd.sibling = -1;
    b.nodes.assign(1, nd);
    return;
  }
  int ni = n - 1;
  std::vector<int> c0(ni), c1(ni), par(ni, -1), rfirst(ni), rcount(ni);
  for (int i = 0; i < ni; i++) {
    int first, last, split;
    karras_node(sc.data(), n, i, &first, &last, &split);
    rfirst[i] = first; rcount[i] = last - first + 1;
    c0[i] = (split == first) ? ~split : split;
    c1[i] = (split + 1 == last) ? ~(split + 1) : split + 1;
    if (c0[i] >= 0) par[c0[i]] = i;
    if (c1[i] >= 0) par[c1[i]] = i;
  }
  // boxes bottom-up: process nodes by decreasing depth (iterative post-order)
  std::vector<float> node_box(6 * (size_t)ni);
  std::vector<int> order, depth(ni, 0), stack{0};
  while (!stack.empty()) {
    int i = stack.back(); stack.pop_back();
    order.push_back(i);
    if (c0[i] >= 0) { depth[c0[i]] = depth[i] + 1; stack.push_back(c0[i]); }
    if (c1[i] >= 0) { depth[c1[i]] = depth[i] + 1; stack.push_back(c1[i]); }
    b.max_depth = std::max(b.max_depth, depth[i] + 1);
  }
  if ((int)order.size() != ni) { b.nodes.clear(); return; }  // topology error
  for (int q = ni - 1; q >= 0; q--) {
    int i = order[q];
    float b0[6], b1[6];
    box_of(b, leaf_box, node_box, c0[i], b0);
    box_of(b, leaf_box, node_box, c1[i], b1);
    for (int k = 0; k < 3; k++) {
      node_box[6 * (size_t)i + k] = fminf(b0[k], b1[k]);
      node_box[6 * (size_t)i + 3 + k] = fmaxf(b0[k + 3], b1[k + 3]);
    }
  }
  b.nodes.resize(ni);
  for (int i = 0; i < ni; i++) {
    Node nd;
    box_of(b, leaf_box, node_box, c0[i], nd.box);
    box_of(b, leaf_box, node_box, c1[i], nd.box + 6);
    auto fin = [&](int k) { return (k >= 0 && rcount[k] <= g_leaf_max) ? leaf_encode(rfirst[k], rcount[k]) : k; };
    nd.c0 = fin(c0[i]); nd.c1 = fin(c1[i]); nd.parent = par[i];
    nd.sibling = -1;
    if (par[i] >= 0) nd.sibling = fin((c0[par[i]] == i) ? c1[par[i]] : c0[par[i]]);
    b.nodes[i] = nd;
  }
}

extern "C" __attribute__((visibility("default")))
int sim_render(const float* verts, int64_t V, const int32_t* tris, int64_t T, const float* normals,
               const float* views, int n_views, double radius, int res, double plane_scale, double spacing,
               int use_z, int split_z, int enc, double zscale, float* out_feat, int32_t* out_ids,
               double* out_t, long long* work /* nodes, tris, max_depth */, int mode /* 0 traverse(), 1 lane steps */) {
  SimBvh b;
  build(verts, V, tris, T, b);
  if (b.nodes.empty()) return 1;
  int C = (use_z && split_z) ? 4 : 3;
  SimCount cnt;
  SimFetch f{b.nodes.data()};
  for (int v = 0; v < n_views; v++) {
    View vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw, 1e-10);
    for (int j = 0; j < res; j++)
      for (int i = 0; i < res; i++) {
        double o[3];
        view_ray(vw, res, i, j, spacing, o);
        (void)mode;
        Hit h = traverse(f, b.tv0.data(), b.tv1.data(), b.tv2.data(), o, vw.dir, vw.dirf, vw.invf, vw.slack,
                         vw.tslack, 1e-20, cnt);
        size_t pix = ((size_t)v * res + j) * res + i;
        float px[4] = {0, 0, 0, 0};
        double t = -1.0;
        if (h.slot >= 0) {
          float4 a = b.tv0[h.slot], bb = b.tv1[h.slot], e = b.tv2[h.slot];
          float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
          double x[3], w[3];
          tri_exact(o, vw.dir, P0, P1, P2, DBL_MAX, 1e-20, &t, x, w);
          const int32_t* vid = &b.vid[4 * (size_t)h.slot];
          shade(o, x, w, normals + 3 * vid[0], normals + 3 * vid[1], normals + 3 * vid[2], use_z, split_z, enc,
                zscale, px);
        }
        if (out_feat) memcpy(out_feat + pix * C, px, 4 * C);
        if (out_ids) out_ids[pix] = h.cell;
        if (out_t) out_t[pix] = t;
      }
  }
  if (work) { work[0] = cnt.nodes; work[1] = cnt.tris; work[2] = b.max_depth; }
  return 0;
}

// ---- emulation of the render kernel's packet traversal (trace.cu): 32 lanes in
// lockstep, ballots as loops.  Checks the vote / trail / ring / link logic and
// the FMA slab test on the CPU.
static void packet_trace(const SimBvh& b, const View& vw, const double (*o)[3], int n_active, int ring_cap,
                         Hit* out, SimCount& cnt) {
  const Node* nodes = b.nodes.data();
  SlabRay sr[32];
  double best_t[32]; int best_slot[32], best_cell[32]; float tcull[32];
  bool negx = vw.dirf[0] < 0.f, negy = vw.dirf[1] < 0.f, negz = vw.dirf[2] < 0.f;
  for (int l = 0; l < 32; l++) {
    float of[3] = {0, 0, 0};
    if (l < n_active) for (int k = 0; k < 3; k++) of[k] = (float)o[l][k];
    float inv[3] = {clamp_inv(vw.dirf[0]), clamp_inv(vw.dirf[1]), clamp_inv(vw.dirf[2])};
    slab_ray_setup(of, inv, vw.slack, &sr[l]);
    best_t[l] = DBL_MAX; best_slot[l] = -1; best_cell[l] = 0x7fffffff;
    tcull[l] = l < n_active ? 1.0f + vw.tslack : -1.0f;
  }
  int node = 0; uint64_t trail = 1;
  std::vector<int> ring(ring_cap); int ring_head = 0, ring_count = 0;
  for (;;) {
    const Node N = nodes[node];
    cnt.node();
    bool h0[32], h1[32]; float tn0[32], tn1[32];
    bool any0 = false, any1 = false;
    for (int l = 0; l < 32; l++) {
      h0[l] = slab_fma(N.box, sr[l], negx, negy, negz, tcull[l], &tn0[l]);
      h1[l] = slab_fma(N.box + 6, sr[l], negx, negy, negz, tcull[l], &tn1[l]);
      any0 |= h0[l]; any1 |= h1[l];
    }
    for (int side = 0; side < 2; side++) {
      int c = side ? N.c1 : N.c0;
      if ((side ? any1 : any0) && c < 0)
       for (int slot = leaf_first(c); slot < leaf_first(c) + leaf_count(c); slot++) {
        float4 a = b.tv0[slot], bb = b.tv1[slot], e = b.tv2[slot];
        float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
        int cell; memcpy(&cell, &a.w, 4);
        for (int l = 0; l < 32; l++) {
          if (!(side ? h1[l] : h0[l])) continue;
          cnt.tri();
          double t, x[3], w[3];
          if (tri_exact(o[l], vw.dir, P0, P1, P2, best_t[l], 1e-20, &t, x, w) &&
              (t < best_t[l] || (t == best_t[l] && cell < best_cell[l]))) {
            best_t[l] = t; best_slot[l] = slot; best_cell[l] = cell;
            tcull[l] = fminf(tcull[l], FB_F2F_RU(t) + vw.tslack);
          }
        }
      }
    }
    int nd0 = 0, nd1 = 0, npref1 = 0, nany = 0;
    for (int l = 0; l < 32; l++) {
      bool w0 = h0[l] && N.c0 >= 0 && tn0[l] <= tcull[l];
      bool w1 = h1[l] && N.c1 >= 0 && tn1[l] <= tcull[l];
      nd0 += w0; nd1 += w1; nany += (w0 || w1);
      npref1 += (w1 && (!w0 || tn1[l] < tn0[l]));
    }
    if (nd0 || nd1) {
      bool both = nd0 && nd1;
      bool swap = both && 2 * npref1 > nany;
      trail = (trail << 1) | (both ? 1ull : 0ull);
      if (both) {
        ring_head = (ring_head + 1) % ring_cap;
        ring[ring_head] = swap ? N.c0 : N.c1;
        ring_count = ring_count < ring_cap ? ring_count + 1 : ring_cap;
      }
      node = both ? (swap ? N.c1 : N.c0) : (nd0 ? N.c0 : N.c1);
      continue;
    }
    int k = ctz64(trail);
    trail >>= k;
    if (trail == 1ull) break;
    trail ^= 1ull;
    if (ring_count > 0) {
      node = ring[ring_head];
      ring_head = (ring_head - 1 + ring_cap) % ring_cap;
      ring_count--;
    } else {
      int nd = node;
      for (int i = 0; i < k; i++) nd = nodes[nd].parent;
      node = nodes[nd].sibling;
    }
  }
  for (int l = 0; l < n_active; l++) {
    out[l].t = best_t[l]; out[l].slot = best_slot[l]; out[l].cell = best_slot[l] >= 0 ? best_cell[l] : -1;
  }
}

extern "C" __attribute__((visibility("default")))
int sim_render_packet(const float* verts, int64_t V, const int32_t* tris, int64_t T, const float* views, int n_views,
                      double radius, int res, double plane_scale, double spacing, int ring_cap, int32_t* out_ids,
                      double* out_t, long long* work) {
  SimBvh b;
  build(verts, V, tris, T, b);
  if (b.nodes.empty()) return 1;
  SimCount cnt;
  for (int v = 0; v < n_views; v++) {
    View vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw, 1e-10);
    // 8x4 tiles like the kernel (no bounds compaction: partial tiles run with idle lanes)
    for (int ty = 0; ty < (res + 3) / 4; ty++)
      for (int tx = 0; tx < (res + 7) / 8; tx++) {
        double o[32][3]; int pix[32]; int n = 0;
        for (int l = 0; l < 32; l++) {
          int px = tx * 8 + (l & 7), py = ty * 4 + (l >> 3);
          if (px < res && py < res) { view_ray(vw, res, px, py, spacing, o[n]); pix[n] = py * res + px; n++; }
        }
        Hit h[32];
        packet_trace(b, vw, o, n, ring_cap, h, cnt);
        for (int l = 0; l < n; l++) {
          size_t id = (size_t)v * res * res + pix[l];
          double t = -1.0;
          if (h[l].slot >= 0) t = h[l].t;
          out_ids[id] = h[l].cell; out_t[id] = t;
        }
      }
  }
  if (work) { work[0] = cnt.nodes; work[1] = cnt.tris; work[2] = b.max_depth; }
  return 0;
}

// ---- packet traversal with float32 screening (render_kernel<.., SCREEN = true>): candidates are
// classified by screen_tri, kept in a short per-lane list and settled by the exact test in batches.
static void packet_trace_screen(const SimBvh& b, const View& vw, const ViewF& vf, const double (*o)[3],
                                int n_active, int pend_cap, Hit* out, SimCount& cnt, long long* cls_count) {
  // pend_cap > 0: fused kernel with in-loop resolution; pend_cap < 0: split design (screen.cu): the
  // trace pass only lists candidates (|pend_cap| slots, overflow -> exact per-ray traversal in resolve)
  const bool split = pend_cap < 0;
  if (split) pend_cap = -pend_cap;
  bool overflow[32] = {false};
  const Node* nodes = b.nodes.data();
  SlabRay sr[32];
  double best_t[32]; int best_slot[32], best_cell[32]; float tcull[32];
  float qx[32], qy[32], zo[32], uz[32]; int nc[32];
  std::vector<int> cslot(32 * pend_cap); std::vector<float> clo(32 * pend_cap);
  bool negx = vw.dirf[0] < 0.f, negy = vw.dirf[1] < 0.f, negz = vw.dirf[2] < 0.f;
  const float inv_two_r = 1.000001f / vf.two_r;
  for (int l = 0; l < 32; l++) {
    float of[3] = {0, 0, 0};
    double oo[3] = {0, 0, 0};
    if (l < n_active) for (int k = 0; k < 3; k++) { of[k] = (float)o[l][k]; oo[k] = o[l][k]; }
    float inv[3] = {clamp_inv(vw.dirf[0]), clamp_inv(vw.dirf[1]), clamp_inv(vw.dirf[2])};
    slab_ray_setup(of, inv, vw.slack, &sr[l]);
    best_t[l] = DBL_MAX; best_slot[l] = -1; best_cell[l] = 0x7fffffff;
    tcull[l] = l < n_active ? 1.0f + vw.tslack : -1.0f;
    screen_ray(vw, oo, &qx[l], &qy[l], &zo[l]);
    uz[l] = INFINITY; nc[l] = 0;
  }
  auto exact_test = [&](int l, int slot) {
    float4 a = b.tv0[slot], bb = b.tv1[slot], e = b.tv2[slot];
    float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
    int cell; memcpy(&cell, &a.w, 4);
    cnt.tri();
    double t, x[3], w[3];
    if (tri_exact(o[l], vw.dir, P0, P1, P2, best_t[l], 1e-20, &t, x, w) &&
        (t < best_t[l] || (t == best_t[l] && cell < best_cell[l]))) {
      best_t[l] = t; best_slot[l] = slot; best_cell[l] = cell;
      return true;
    }
    return false;
  };
  auto resolve = [&]() {
    for (int l = 0; l < 32; l++)
      while (nc[l] > 0) {
        nc[l]--;
        int slot = cslot[l * pend_cap + nc[l]];
        if (clo[l * pend_cap + nc[l]] <= uz[l] && exact_test(l, slot)) {
          float zx = fmaf((float)best_t[l], vf.two_r, zo[l]) + 2.0f * vf.e_z;
          uz[l] = fminf(uz[l], zx);
          tcull[l] = fminf(tcull[l], FB_F2F_RU(best_t[l]) + vw.tslack);
        }
      }
  };
  int node = 0; uint64_t trail = 1;
  std::vector<int> ring(16); int ring_head = 0, ring_count = 0;
  for (;;) {
    const Node N = nodes[node];
    cnt.node();
    bool h0[32], h1[32]; float tn0[32], tn1[32];
    bool any0 = false, any1 = false;
    for (int l = 0; l < 32; l++) {
      h0[l] = slab_fma(N.box, sr[l], negx, negy, negz, tcull[l], &tn0[l]);
      h1[l] = slab_fma(N.box + 6, sr[l], negx, negy, negz, tcull[l], &tn1[l]);
      any0 |= h0[l]; any1 |= h1[l];
    }
    for (int side = 0; side < 2; side++) {
      int c = side ? N.c1 : N.c0;
      if ((side ? any1 : any0) && c < 0)
       for (int slot = leaf_first(c); slot < leaf_first(c) + leaf_count(c); slot++) {
        float4 a = b.tv0[slot], bb = b.tv1[slot], e = b.tv2[slot];
        float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
        bool full = false;
        for (int l = 0; l < 32; l++) {
          if (!(side ? h1[l] : h0[l])) continue;
          float zlo, zhi;
          int cls = screen_tri(vf, qx[l], qy[l], zo[l], P0, P1, P2, &zlo, &zhi);
          if (cls_count) cls_count[cls]++;
          if (cls != SC_MISS && zlo <= uz[l] && !overflow[l]) {
            if (cls == SC_SURE && zhi < uz[l]) {
              uz[l] = zhi;
              tcull[l] = fminf(tcull[l], fmaf(uz[l] - zo[l] + 2.0f * vf.e_z, inv_two_r, vw.tslack));
            }
            if (split && nc[l] == pend_cap) {  // compact, then give up on the list
              int k = 0;
              for (int i = 0; i < pend_cap; i++)
                if (clo[l * pend_cap + i] <= uz[l]) { clo[l * pend_cap + k] = clo[l * pend_cap + i]; cslot[l * pend_cap + k] = cslot[l * pend_cap + i]; k++; }
              nc[l] = k;
            }
            if (split && nc[l] == pend_cap) overflow[l] = true;
            else { cslot[l * pend_cap + nc[l]] = slot; clo[l * pend_cap + nc[l]] = zlo; nc[l]++; }
          }
          if (!split) full |= nc[l] == pend_cap;
        }
        if (full) resolve();
      }
    }
    int nd0 = 0, nd1 = 0, npref1 = 0, nany = 0;
    for (int l = 0; l < 32; l++) {
      bool w0 = h0[l] && N.c0 >= 0 && tn0[l] <= tcull[l];
      bool w1 = h1[l] && N.c1 >= 0 && tn1[l] <= tcull[l];
      nd0 += w0; nd1 += w1; nany += (w0 || w1);
      npref1 += (w1 && (!w0 || tn1[l] < tn0[l]));
    }
    if (nd0 || nd1) {
      bool both = nd0 && nd1;
      bool swap = both && 2 * npref1 > nany;
      trail = (trail << 1) | (both ? 1ull : 0ull);
      if (both) {
        ring_head = (ring_head + 1) % 16;
        ring[ring_head] = swap ? N.c0 : N.c1;
        ring_count = ring_count < 16 ? ring_count + 1 : 16;
      }
      node = both ? (swap ? N.c1 : N.c0) : (nd0 ? N.c0 : N.c1);
      continue;
    }
    int k = ctz64(trail);
    trail >>= k;
    if (trail == 1ull) break;
    trail ^= 1ull;
    if (ring_count > 0) {
      node = ring[ring_head];
      ring_head = (ring_head + 15) % 16;
      ring_count--;
    } else {
      int nd = node;
      for (int i = 0; i < k; i++) nd = nodes[nd].parent;
      node = nodes[nd].sibling;
    }
  }
  resolve();
  for (int l = 0; l < n_active; l++) {
    if (overflow[l]) {
      SimFetch f{nodes};
      out[l] = traverse(f, b.tv0.data(), b.tv1.data(), b.tv2.data(), o[l], vw.dir, vw.dirf, vw.invf, vw.slack, vw.tslack,
                        1e-20, cnt);
      if (cls_count) cls_count[1] += 1000000;  // make overflows visible in the class counters
      continue;
    }
    out[l].t = best_t[l]; out[l].slot = best_slot[l]; out[l].cell = best_slot[l] >= 0 ? best_cell[l] : -1;
  }
}

extern "C" __attribute__((visibility("default")))
int sim_render_packet_screen(const float* verts, int64_t V, const int32_t* tris, int64_t T, const float* views,
                             int n_views, double radius, int res, double plane_scale, double spacing, int pend_cap,
                             int32_t* out_ids, double* out_t, long long* work /* nodes, exact tests, depth, miss, maybe, sure */) {
  SimBvh b;
  build(verts, V, tris, T, b);
  if (b.nodes.empty()) return 1;
  float maxc = 0.f;
  for (int64_t i = 0; i < 3 * V; i++) maxc = fmaxf(maxc, fabsf(verts[i]));
  SimCount cnt;
  long long cls[3] = {0, 0, 0};
  for (int v = 0; v < n_views; v++) {
    View vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw, 1e-10);
    ViewF vf;
    viewf_setup(vw, radius, spacing, maxc, 1e-10, &vf);
    viewf_pixels(vw, spacing, res, &vf);
    for (int ty = 0; ty < (res + 3) / 4; ty++)
      for (int tx = 0; tx < (res + 7) / 8; tx++) {
        double o[32][3]; int pix[32]; int n = 0;
        for (int l = 0; l < 32; l++) {
          int px = tx * 8 + (l & 7), py = ty * 4 + (l >> 3);
          if (px < res && py < res) { view_ray(vw, res, px, py, spacing, o[n]); pix[n] = py * res + px; n++; }
        }
        Hit h[32];
        packet_trace_screen(b, vw, vf, o, n, pend_cap, h, cnt, cls);
        for (int l = 0; l < n; l++) {
          size_t id = (size_t)v * res * res + pix[l];
          out_ids[id] = h[l].cell; out_t[id] = h[l].slot >= 0 ? h[l].t : -1.0;
        }
      }
  }
  if (work) { work[0] = cnt.nodes; work[1] = cnt.tris; work[2] = b.max_depth; work[3] = cls[0]; work[4] = cls[1]; work[5] = cls[2]; }
  return 0;
}

// ---- emulation of the scan-conversion path (raster.cu + resolve_kernel): pass A nearest certain depth,
// pass B candidate collection, then the exact decision per pixel (overflow -> exact per-ray traversal).
extern "C" __attribute__((visibility("default")))
int sim_render_raster(const float* verts, int64_t V, const int32_t* tris, int64_t T, const float* views,
                      int n_views, double radius, int res, double plane_scale, double spacing, int cand_cap,
                      int32_t* out_ids, double* out_t, long long* work /* [7]: pixel tests, exact tests, overflow pixels, in: 1 = audit every decision (out: 0), in: order seed, in: predicate / out: unsound decisions, MAYBE box pixels */) {
  SimBvh b;
  build(verts, V, tris, T, b);
  if (b.nodes.empty()) return 1;
  float maxc = 0.f;
  for (int64_t i = 0; i < 3 * V; i++) maxc = fmaxf(maxc, fabsf(verts[i]));
  long long n_tests = 0, n_exact = 0, n_over = 0, n_unsound = 0, n_maybe = 0;
  const bool audit = work && work[3] == 1;  // every MISS / SURE decision of the predicate against the exact test
  // in: work[5] selects the per-pixel predicate: 0 as raster.cu (linear form for boxes up to RT_BIG pixels, raster_pixel
  // for the larger ones, which raster_big_kernel and raster_kernel's inline fallback walk), 1 the linear form for every
  // box, 2 raster_pixel for every box
  const long long pred_mode = work ? work[5] : 0;
  // in: work[6] = margins of the predicate in percent of their value (0 = 100): the audit with thinner margins measures
  // how much headroom the rounding analysis leaves (the results stay exact either way only while no decision is wrong)
  const float margin_scale = work && work[6] > 0 ? (float)work[6] * 0.01f : 1.0f;
  const long long order_seed = work ? work[4] : 0;  // in: work[4] != 0 shuffles the arrival order of the triangles
  SimCount cnt;
  SimFetch f{b.nodes.data()};
  for (int v = 0; v < n_views; v++) {
    View vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw, 1e-10);
    ViewF vf;
    viewf_setup(vw, radius, spacing, maxc, 1e-10, &vf);
    viewf_pixels(vw, spacing, res, &vf);
    // single pass, as raster_kernel: front[pixel] = smallest front_key over the SURE triangles; what a triangle
    // learns from the entry it displaced / lost to / read decides who joins the extras (settle_pixel).
    // Triangles arrive in slot order, or in a seeded random order to stand in for the GPU's arbitrary one.
    std::vector<uint64_t> front((size_t)res * res, ~0ull);
    std::vector<std::vector<int>> cand((size_t)res * res);
    std::vector<int64_t> order(T);
    for (int64_t k = 0; k < T; k++) order[k] = k;
    if (order_seed) {
      uint64_t st = (uint64_t)order_seed * 0x9E3779B97F4A7C15ull + (uint64_t)v;
      for (int64_t k = T - 1; k > 0; k--) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        std::swap(order[k], order[(int64_t)((st >> 33) % (uint64_t)(k + 1))]);
      }
    }
    for (int64_t oi = 0; oi < T; oi++) {
      const int64_t slot = order[oi];
      float4 a = b.tv0[slot], bb = b.tv1[slot], e = b.tv2[slot];
      float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
      TriR t;
      PixRange r;
      if (!raster_project(vf, P0, P1, P2, res, &t, &r)) continue;
      const long long npx = (long long)(r.i1 - r.i0 + 1) * (r.j1 - r.j0 + 1);
      raster_setup(vf, 0.0f, P0, P1, P2, &t);
      if (!(t.flags & 1)) continue;
      t.g0 *= margin_scale; t.g1 *= margin_scale; t.g2 *= margin_scale; t.G *= margin_scale;
      const uint64_t key = front_key(t.zlo, t.zhi, (int)slot);
      const uint32_t zlo_ord = ford_bits(t.zlo);
      const bool linear = pred_mode == 1 || (pred_mode == 0 && npx <= 96);
      TriL L;
      raster_linear_setup(vf, t, r.i0, r.j0, &L);
      for (int j = r.j0; j <= r.j1; j++) {
        const int lo = r.i0, hi = r.i1;
        if (audit)
          for (int i = r.i0; i <= r.i1; i++) {
            const int c2 = linear ? raster_linear_pixel(L, (float)(i - r.i0), (float)(j - r.j0))
                                  : raster_pixel(t, fmaf((float)i, vf.qdx, vf.q0x), fmaf((float)j, vf.qdy, vf.q0y));
            // soundness of the predicate itself: MISS => the exact test rejects, SURE => it accepts
            if (c2 != SC_MAYBE) {
              double o[3], tt, xx[3], ww[3];
              view_ray(vw, res, i, j, spacing, o);
              const bool ex = tri_exact(o, vw.dir, P0, P1, P2, DBL_MAX, 1e-20, &tt, xx, ww);
              if ((c2 == SC_MISS && ex) || (c2 == SC_SURE && !ex)) n_unsound++;
              if (c2 == SC_SURE && ex) {  // and the hit depth lies inside [zlo, zhi]
                const double z = tt * 2.0 * radius;
                if (z < (double)t.zlo || z > (double)t.zhi) n_unsound++;
              }
            } else n_maybe++;
          }
        for (int i = lo; i <= hi; i++) {
          float qx = fmaf((float)i, vf.qdx, vf.q0x), qy = fmaf((float)j, vf.qdy, vf.q0y);
          int cls = linear ? raster_linear_pixel(L, (float)(i - r.i0), (float)(j - r.j0)) : raster_pixel(t, qx, qy);
          n_tests++;
          if (cls == SC_MISS) continue;
          size_t pix = (size_t)j * res + i;
          const uint64_t old = front[pix];
          if (cls == SC_SURE && key < old) front[pix] = key;
          if (cls == SC_SURE && key < old) {
            if (old != ~0ull && front_key_zlo_lower(old) <= t.zhi) cand[pix].push_back(FB_FRONT_SLOT(old));
          } else if (zlo_ord <= (uint32_t)(old >> 32)) {
            cand[pix].push_back((int)slot);
          }
        }
      }
    }
    for (size_t pix = 0; pix < (size_t)res * res; pix++)
      if (front[pix] != ~0ull && cand_cap >= 0) cand[pix].insert(cand[pix].begin(), FB_FRONT_SLOT(front[pix]));
    for (int j = 0; j < res; j++)
      for (int i = 0; i < res; i++) {
        size_t pix = (size_t)j * res + i;
        double o[3];
        view_ray(vw, res, i, j, spacing, o);
        Hit h; h.t = DBL_MAX; h.slot = -1; h.cell = 0x7fffffff;
        if (cand_cap < 0) {  // model of the last-resort path: screened LBVH re-trace of the pixel
          n_over++;
          float qx = fmaf((float)i, vf.qdx, vf.q0x), qy = fmaf((float)j, vf.qdy, vf.q0y);
          h = traverse(f, b.tv0.data(), b.tv1.data(), b.tv2.data(), o, vw.dir, vw.dirf, vw.invf, vw.slack, vw.tslack,
                       1e-20, cnt, vf.enabled ? &vf : nullptr, qx, qy, 0.0f);
        } else {
          // more candidates than the per-pixel lists hold go through the spill list: same minimum of (t, cell id)
          if ((int)cand[pix].size() > cand_cap) n_over++;
          for (int slot : cand[pix]) {
            float4 a = b.tv0[slot], bb = b.tv1[slot], e = b.tv2[slot];
            float P0[3] = {a.x, a.y, a.z}, P1[3] = {bb.x, bb.y, bb.z}, P2[3] = {e.x, e.y, e.z};
            int cell; memcpy(&cell, &a.w, 4);
            double t, x[3], w[3];
            n_exact++;
            if (tri_exact(o, vw.dir, P0, P1, P2, h.t, 1e-20, &t, x, w) && (t < h.t || (t == h.t && cell < h.cell))) {
              h.t = t; h.slot = slot; h.cell = cell;
            }
          }
          if (h.slot < 0) h.cell = -1;
        }
        size_t id = (size_t)v * res * res + pix;
        out_ids[id] = h.cell; out_t[id] = h.slot >= 0 ? h.t : -1.0;
      }
  }
  if (work) { work[0] = n_tests; work[1] = n_exact; work[2] = n_over; if (audit) { work[3] = 0; work[5] = n_unsound; work[6] = n_maybe; } }
  return 0;
}

// ---- raster_kernel's row / column arithmetic: (k * rt_magic(ipr)) >> 16 against k / ipr; returns the mismatches
extern "C" __attribute__((visibility("default")))
long long sim_rt_magic_check(int max_ipr, int max_k) {
  long long bad = 0;
  for (int ipr = 1; ipr <= max_ipr; ipr++) {
    const uint32_t m = rt_magic(ipr);
    for (int k = 0; k < max_k; k++)
      if ((int)(((uint32_t)k * m) >> 16) != k / ipr) bad++;
  }
  return bad;
}

// ---- raster_kernel's flat enumeration, one warp replayed with loops: 32 boxes (w[l] x h[l] pixels, 0 = not visible),
// U pixels per work item.  Emits (lane of the triangle, row, column) of every pixel the kernel would test, in the
// kernel's order, into out[3 * n]; returns n (or -1 if a record index leaves the compacted list).
extern "C" __attribute__((visibility("default")))
long long sim_flat_enumeration(const int* w, const int* h, int U, int* out, long long cap) {
  int rec_lane[32], rec_w[32], S[32];
  uint32_t magic[32];
  int n_rec = 0, total = 0;
  for (int l = 0; l < 32; l++) {
    if (w[l] <= 0 || h[l] <= 0) continue;
    const int ipr = (w[l] + U - 1) / U;
    rec_lane[n_rec] = l; rec_w[n_rec] = w[l]; S[n_rec] = total; magic[n_rec] = rt_magic(ipr);
    total += ipr * h[l];
    n_rec++;
  }
  long long n = 0;
  int started = 0;
  for (int base = 0; base < total; base += 32) {
    uint32_t starts = 0;
    for (int l = 0; l < 32; l++) starts |= rt_start_bit(l < n_rec ? S[l] : 0x7fffffff, base);  // REDUX.OR
    for (int u = 0; u < U; u++)  // the kernel's loop order: pixel u of every lane's item, then u + 1
      for (int lane = 0; lane < 32; lane++) {
        const int item = base + lane;
        if (item >= total) continue;
        const int own = rt_owner(starts, started, rt_lemask(lane));
        if (own < 0 || own >= n_rec) return -1;
        int row, col0;
        rt_item(item - S[own], rec_w[own], magic[own], U, &row, &col0);
        if (u < rec_w[own] - col0) {
          if (n < cap) { out[3 * n] = rec_lane[own]; out[3 * n + 1] = row; out[3 * n + 2] = col0 + u; }
          n++;
        }
      }
    started += rt_popc(starts);
  }
  return n;
}

// ---- audit of the opt-in tolerance shading (flyby_core.cuh shade_approx) on its domain of use: every (triangle,
// pixel) pair the scan-conversion predicate classifies SC_SURE.  For each pair the exact chain (tri_exact + shade)
// and shade_approx are evaluated and compared.
// stats: [0] SURE pairs, [1] pairs shade_approx accepted, [2] SURE pairs the exact test rejects (must be 0)
// err:   [0] largest relative depth error, [1] largest absolute error of a normal channel (in units of the
//        encoding's range: 1 for unit01 / signed, 255 for byte255), [2] largest relative error of a z-multiplied channel
extern "C" __attribute__((visibility("default")))
int sim_shade_approx_audit(const float* verts, int64_t V, const int32_t* tris, int64_t T, const float* normals,
                           const float* views, int n_views, double radius, int res, double plane_scale, double spacing,
                           int use_z, int split_z, int enc, double zscale, long long* stats, double* err) {
  float maxc = 0.f;
  for (int64_t i = 0; i < 3 * V; i++) maxc = fmaxf(maxc, fabsf(verts[i]));
  long long n_sure = 0, n_acc = 0, n_rej = 0;
  double e_depth = 0.0, e_norm = 0.0, e_mul = 0.0;
  const double range = enc == 1 ? 255.0 : 1.0;
  for (int v = 0; v < n_views; v++) {
    View vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw, 1e-10);
    ViewF vf;
    viewf_setup(vw, radius, spacing, maxc, 1e-10, &vf);
    viewf_pixels(vw, spacing, res, &vf);
    for (int64_t c = 0; c < T; c++) {
      const int32_t* tr = tris + 3 * c;
      const float *P0 = verts + 3 * (int64_t)tr[0], *P1 = verts + 3 * (int64_t)tr[1], *P2 = verts + 3 * (int64_t)tr[2];
      const float *N0 = normals + 3 * (int64_t)tr[0], *N1 = normals + 3 * (int64_t)tr[1], *N2 = normals + 3 * (int64_t)tr[2];
      TriR t;
      PixRange r;
      if (!raster_project(vf, P0, P1, P2, res, &t, &r)) continue;
      raster_setup(vf, 0.0f, P0, P1, P2, &t);
      if (!(t.flags & 1)) continue;
      for (int j = r.j0; j <= r.j1; j++)
        for (int i = r.i0; i <= r.i1; i++) {
          if (raster_pixel(t, fmaf((float)i, vf.qdx, vf.q0x), fmaf((float)j, vf.qdy, vf.q0y)) != SC_SURE) continue;
          n_sure++;
          double o[3], tt, x[3], w[3];
          view_ray(vw, res, i, j, spacing, o);
          if (!tri_exact(o, vw.dir, P0, P1, P2, DBL_MAX, 1e-20, &tt, x, w)) { n_rej++; continue; }
          float ref[4] = {0, 0, 0, 0}, got[4] = {0, 0, 0, 0};
          shade(o, x, w, N0, N1, N2, use_z, split_z, enc, zscale, ref);
          if (!shade_approx(o, vw.dir, vw.dlen, P0, P1, P2, N0, N1, N2, use_z, split_z, enc, zscale, got)) continue;
          n_acc++;
          if (use_z && !split_z) {
            // channel = e * z: compare relative to the depth factor (e in [0, range])
            const double zref = tt * vw.dlen * fabs(zscale);
            for (int k = 0; k < 3; k++) e_mul = fmax(e_mul, fabs((double)got[k] - (double)ref[k]) / fmax(zref * range, 1e-300));
          } else {
            for (int k = 0; k < 3; k++) e_norm = fmax(e_norm, fabs((double)got[k] - (double)ref[k]) / range);
            if (use_z) e_depth = fmax(e_depth, fabs((double)got[3] - (double)ref[3]) / fmax(fabs((double)ref[3]), 1e-300));
          }
        }
    }
  }
  stats[0] = n_sure; stats[1] = n_acc; stats[2] = n_rej;
  err[0] = e_depth; err[1] = e_norm; err[2] = e_mul;
  return 0;
}
