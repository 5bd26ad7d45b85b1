/*
 * flyby_oracle.c -- CPU restatement of Fly-by-CNN's ray-cast fly-by feature path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fly-by-cnn_b200/ (the product) may
 * include, link or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * PARITY: PARTLY PINNED.  Everything the reference computes in its own Python is
 * pinned: tests/golden/ref_*.npz hold vectors produced by importing
 * /root/reference/src/py unmodified (tests/golden/make_ref_golden.py, container-only
 * vtk stand-in tests/refshim) and tests/test_ref_golden.py holds this file to them bit
 * for bit -- O1 unit surface (ScaleSurf), O3 icosahedron viewpoints incl. the
 * float-equality duplicates (LinearSubdivisionFilter + normalize_points), O4 spiral,
 * O5 plane basis / segments (python twin predict.py:86-126), the point-id codec and
 * feature lookup, O9 votes / argmax / cell->point, and post_process.py (postproc_oracle.c).
 * PARITY UNPINNED for the arithmetic that lives inside VTK (vtk==9.2.2,
 * reference src/py/challenge-teeth/flyby.yml:93; C++ tool built on VTK-9.0.1,
 * src/app/README.md:22), which is not vendored under /root/reference and is
 * not installable here: O2 vtkPolyDataNormals (incl. its consistency pass), the
 * float32 storage of vtkPlaneSource, O6 vtkCellLocator / vtkTriangle::IntersectWithLine /
 * vtkPlane::IntersectWithLine / vtkTriangle::EvaluatePosition, O7 shading on those
 * weights, and the constants of vtkPlatonicSolidSource.  Those are restated from VTK 9's
 * published algorithms and anchored on the reference's own call sites, cited per
 * function below (paths relative to /root/reference); oracle/pin_vtk.py is their pin
 * for a machine that has VTK.  The reference ships no golden vectors or tests of its own.
 *
 * Arithmetic: double on float32 inputs, exactly as VTK sees float vtkPoints.
 * Build with -ffp-contract=off so no FMA contraction changes a result.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* small helpers                                                       */
/* ------------------------------------------------------------------ */
static double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double norm3(const double a[3]) { return sqrt(dot3(a, a)); }

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* the reference's ray loop is serial (fly_by_features.cxx:652,742; predict.py:83,123): n = 1 times it that way */
ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ */
/* O1  unit surface                                                    */
/*   src/py/utils.py:249-290 (ScaleSurf, canonical)                    */
/*   src/app/fly_by_features.cxx:393-450 (centre-of-mass / magnitude)  */
/* ------------------------------------------------------------------ */
/* centre_mode: 0 = bbox mid (utils.py:263-265), 1 = centre of mass (cxx:397-403)
 * scale_mode : 0 = 1/|bbox_max - centre| (utils.py:281), 1 = 1/max|p - centre| (cxx:429-435)
 * has_translate / has_scale: caller overrides (utils.py:271,279; --translate, --scale_factor)
 * out = float32((double(p) - centre) * scale); centre_scale_out[4] = centre xyz, scale. */
ORC_API void orc_unit_surf(const float* verts, int64_t V, int centre_mode, int scale_mode,
                           int has_translate, const double* translate, int has_scale,
                           double scale_factor, float* out, double* centre_scale_out) {
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  double sum[3] = {0, 0, 0};
  for (int64_t i = 0; i < V; i++)
    for (int k = 0; k < 3; k++) {
      double p = verts[3 * i + k];
      if (p < lo[k]) lo[k] = p;
      if (p > hi[k]) hi[k] = p;
    }
  /* centre of mass: fixed summation order (1024-point chunks, then the chunk
   * sums) so that a parallel implementation can reproduce it bit for bit */
  for (int64_t b = 0; b < V; b += 1024) {
    double cs[3] = {0, 0, 0};
    int64_t e = b + 1024 < V ? b + 1024 : V;
    for (int64_t i = b; i < e; i++)
      for (int k = 0; k < 3; k++) cs[k] += (double)verts[3 * i + k];
    for (int k = 0; k < 3; k++) sum[k] += cs[k];
  }
  double c[3];
  for (int k = 0; k < 3; k++) {
    if (has_translate) c[k] = translate[k];
    else if (centre_mode == 1) c[k] = sum[k] / (double)V;
    else c[k] = (lo[k] + hi[k]) / 2.0;
  }
  double s;
  if (has_scale) s = scale_factor;
  else if (scale_mode == 1) {
    double m = 0;
    for (int64_t i = 0; i < V; i++) {
      double d[3] = {verts[3 * i] - c[0], verts[3 * i + 1] - c[1], verts[3 * i + 2] - c[2]};
      double n = norm3(d);
      if (n > m) m = n;
    }
    s = 1.0 / m;
  } else {
    double d[3] = {hi[0] - c[0], hi[1] - c[1], hi[2] - c[2]};
    s = 1.0 / norm3(d);
  }
  for (int64_t i = 0; i < V; i++)
    for (int k = 0; k < 3; k++) out[3 * i + k] = (float)(((double)verts[3 * i + k] - c[k]) * s);
  if (centre_scale_out) {
    centre_scale_out[0] = c[0]; centre_scale_out[1] = c[1];
    centre_scale_out[2] = c[2]; centre_scale_out[3] = s;
  }
}

/* ------------------------------------------------------------------ */
/* O2  point normals (vtkPolyDataNormals, splitting off)               */
/*   src/py/utils.py:493-501 ; src/app/fly_by_features.cxx:472-478     */
/* VTK: unit polygon normal (vtkTriangle::ComputeNormal: a=p2-p1,      */
/* b=p0-p1, n=a x b, /|n|) stored float32; point normal = float32 sum  */
/* over incident cells in ascending cell id, then /sqrt(|sum|^2).      */
/* ------------------------------------------------------------------ */
ORC_API void orc_point_normals(const float* verts, int64_t V, const int32_t* tris, int64_t T,
                               float* out) {
  memset(out, 0, sizeof(float) * 3 * (size_t)V);
  for (int64_t c = 0; c < T; c++) {
    const float* p0 = verts + 3 * (int64_t)tris[3 * c];
    const float* p1 = verts + 3 * (int64_t)tris[3 * c + 1];
    const float* p2 = verts + 3 * (int64_t)tris[3 * c + 2];
    double a[3], b[3], n[3];
    for (int k = 0; k < 3; k++) {
      a[k] = (double)p2[k] - (double)p1[k];
      b[k] = (double)p0[k] - (double)p1[k];
    }
    cross3(a, b, n);
    double len = norm3(n);
    float fn[3] = {0.f, 0.f, 0.f};
    if (len != 0.0)
      for (int k = 0; k < 3; k++) fn[k] = (float)(n[k] / len);
    for (int j = 0; j < 3; j++) {
      float* o = out + 3 * (int64_t)tris[3 * c + j];
      for (int k = 0; k < 3; k++) o[k] = o[k] + fn[k];
    }
  }
  for (int64_t i = 0; i < V; i++) {
    float* o = out + 3 * i;
    float s0 = o[0] * o[0], s1 = o[1] * o[1], s2 = o[2] * o[2];
    float ss = (s0 + s1) + s2;
    double len = sqrt((double)ss);
    if (len != 0.0)
      for (int k = 0; k < 3; k++) o[k] = (float)((double)o[k] / len);
  }
}

/* ------------------------------------------------------------------ */
/* O2b consistency pass of vtkPolyDataNormals (ConsistencyOn, VTK's      */
/*   default: src/py/utils.py:493-501, src/app/fly_by_features.cxx:472-478) */
/* vtkPolyDataNormals::RequestData / TraverseAndOrder restated: cells are visited in */
/* waves from every still unvisited cell in ascending id; for edge (p1,p2) of the     */
/* current cell (its CURRENT point order) a not yet visited neighbour must contain    */
/* ... p2, p1 ... in that cyclic order, else vtkPolyData::ReverseCell makes it        */
/* (p2,p1,p0).  Deviation shared with the product: edges used by three or more cells  */
/* are not crossed (NonManifoldTraversal off).  tris is edited in place; flipped[T]   */
/* (optional) gets 1 for reversed cells.  Returns the number of reversed cells, -1 if */
/* some component has no consistent orientation (an already visited neighbour         */
/* disagrees).                                                                        */
/* ------------------------------------------------------------------ */
ORC_API int64_t orc_orient_cells(int32_t* tris, int64_t T, int64_t V, uint8_t* flipped) {
  /* point -> cells links, ascending cell id (vtkCellLinks::BuildLinks) */
  int64_t* start = (int64_t*)calloc((size_t)V + 2, sizeof(int64_t));
  for (int64_t i = 0; i < 3 * T; i++) start[tris[i] + 1]++;
  for (int64_t v = 0; v < V; v++) start[v + 1] += start[v];
  int32_t* links = (int32_t*)malloc(sizeof(int32_t) * (size_t)(3 * T > 0 ? 3 * T : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(V + 1));
  memcpy(fill, start, sizeof(int64_t) * (size_t)(V + 1));
  for (int64_t c = 0; c < T; c++)
    for (int j = 0; j < 3; j++) {
      int32_t v = tris[3 * c + j];
      /* a cell that lists a point twice appears once in its links */
      if ((j == 1 && v == tris[3 * c]) || (j == 2 && (v == tris[3 * c] || v == tris[3 * c + 1]))) continue;
      links[fill[v]++] = (int32_t)c;
    }
  uint8_t* visited = (uint8_t*)calloc((size_t)(T > 0 ? T : 1), 1);
  int32_t* wave = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1));
  if (flipped) memset(flipped, 0, (size_t)T);
  int64_t n_flip = 0;
  int bad = 0;
  for (int64_t seed = 0; seed < T; seed++) {
    if (visited[seed]) continue;
    int64_t head = 0, tail = 0;
    wave[tail++] = (int32_t)seed;
    visited[seed] = 1;
    while (head < tail) {
      const int32_t c = wave[head++];
      for (int j = 0; j < 3; j++) {
        const int32_t p1 = tris[3 * (int64_t)c + j], p2 = tris[3 * (int64_t)c + (j + 1) % 3];
        if (p1 == p2) continue;
        /* GetCellEdgeNeighbors: the other cells that use both points */
        int32_t nb = -1;
        int n_nb = 0;
        for (int64_t q = start[p1]; q < fill[p1]; q++) {
          const int32_t n = links[q];
          if (n == c) continue;
          const int32_t* t = tris + 3 * (int64_t)n;
          if (t[0] == p2 || t[1] == p2 || t[2] == p2) { nb = n; n_nb++; }
        }
        if (n_nb != 1) continue;
        const int32_t* t = tris + 3 * (int64_t)nb;
        int l = t[0] == p2 ? 0 : (t[1] == p2 ? 1 : 2);
        const int consistent = t[(l + 1) % 3] == p1;
        if (visited[nb]) {
          if (!consistent) bad = 1;
          continue;
        }
        if (!consistent) {
          int32_t* w = tris + 3 * (int64_t)nb;
          const int32_t t0 = w[0];
          w[0] = w[2];
          w[2] = t0;
          if (flipped) flipped[nb] = 1;
          n_flip++;
        }
        visited[nb] = 1;
        wave[tail++] = nb;
      }
    }
  }
  free(start); free(links); free(fill); free(visited); free(wave);
  return bad ? -1 : n_flip;
}

/* ------------------------------------------------------------------ */
/* O3  icosahedron-subdivision viewpoints                              */
/*   src/py/utils.py:36-50 (CreateIcosahedron)                         */
/*   src/py/LinearSubdivisionFilter.py:42-67 (lattice order, de-dup)   */
/*   src/py/utils.py:22-31 (normalize_points)                          */
/* vtkPlatonicSolidSource tables restated from VTK 9 (un-vendored).    */
/* ------------------------------------------------------------------ */
static const double ICO_PTS[12][3] = {
    {0.0, 0.0, -1.0},           {0.0, 0.894427, -0.447214},      {0.850651, 0.276393, -0.447214},
    {0.525731, -0.723607, -0.447214}, {-0.525731, -0.723607, -0.447214}, {-0.850651, 0.276393, -0.447214},
    {0.0, -0.894427, 0.447214}, {0.850651, -0.276393, 0.447214}, {0.525731, 0.723607, 0.447214},
    {-0.525731, 0.723607, 0.447214}, {-0.850651, -0.276393, 0.447214}, {0.0, 0.0, 1.0}};
static const int ICO_FACES[20][3] = {
    {0, 5, 4},  {0, 4, 3},  {0, 3, 2},  {0, 2, 1},   {0, 1, 5},  {5, 1, 9},  {5, 9, 10},
    {5, 10, 4}, {4, 10, 6}, {4, 6, 3},  {3, 6, 7},   {3, 7, 2},  {2, 7, 8},  {2, 8, 1},
    {1, 8, 9},  {9, 8, 11}, {9, 11, 10}, {10, 11, 6}, {6, 11, 7}, {7, 11, 8}};
/* uniform solid scale applied by vtkPlatonicSolidSource before float storage
 * (recalled constant of vtkPlatonicSolidSource.cxx, not byte-verified; after normalize_points
 * it only decides the last float32 bit of a viewpoint and which float-equality duplicates
 * leak -- tests/refshim/vtk uses the same value, so the reference's own subdivision code
 * run on it reproduces orc_icosahedron(dedupe = 1) bit for bit) */
#define ICO_SOLID_SCALE (1.0 / 0.58778524999243)

/* dedupe: 0 = exact integer lattice (canonical, 10 L^2 + 2 points),
 *         1 = float32 exact-equality emulation of
 *             vtkIncrementalOctreePointLocator (LinearSubdivisionFilter.py:64-66).
 * out must hold 20*(L+1)*(L+2)/2 * 3 floats.  Returns the number of viewpoints. */
ORC_API int orc_icosahedron(int L, double radius, int dedupe, float* out) {
  if (L < 1) return -1;
  float base[12][3];
  for (int i = 0; i < 12; i++)
    for (int k = 0; k < 3; k++) base[i][k] = (float)(ICO_PTS[i][k] * ICO_SOLID_SCALE);
  int cap = 20 * (L + 1) * (L + 2) / 2;
  float* pts = (float*)malloc(sizeof(float) * 3 * (size_t)cap);
  /* lattice key: sorted (vertex, weight) pairs with non-zero weight */
  int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * 3 * (size_t)cap);
  int n = 0;
  for (int f = 0; f < 20; f++) {
    const int* fv = ICO_FACES[f];
    double p1[3], p2[3], p3[3], d12[3], d13[3];
    for (int k = 0; k < 3; k++) {
      p1[k] = base[fv[0]][k]; p2[k] = base[fv[1]][k]; p3[k] = base[fv[2]][k];
      d12[k] = (p2[k] - p1[k]) / (double)L;
      d13[k] = (p3[k] - p1[k]) / (double)L;
    }
    for (int s13 = 0; s13 <= L; s13++)
      for (int s12 = 0; s12 <= L - s13; s12++) {
        float q[3];
        for (int k = 0; k < 3; k++) q[k] = (float)((p1[k] + s12 * d12[k]) + s13 * d13[k]);
        int found = -1;
        if (dedupe == 1) {
          for (int j = 0; j < n && found < 0; j++)
            if (pts[3 * j] == q[0] && pts[3 * j + 1] == q[1] && pts[3 * j + 2] == q[2]) found = j;
        } else {
          int w[3] = {L - s12 - s13, s12, s13};
          int64_t key[3] = {-1, -1, -1};
          int m = 0;
          for (int a = 0; a < 3; a++)
            if (w[a] > 0) key[m++] = (int64_t)fv[a] * 100000 + w[a];
          /* sort the up-to-3 entries */
          for (int a = 0; a < m; a++)
            for (int b = a + 1; b < m; b++)
              if (key[b] < key[a]) { int64_t t = key[a]; key[a] = key[b]; key[b] = t; }
          /* interior points (3 non-zero weights) are unique to the face */
          if (m == 3) { key[0] += (int64_t)(f + 1) * 100000000000LL; }
          for (int j = 0; j < n && found < 0; j++)
            if (keys[3 * j] == key[0] && keys[3 * j + 1] == key[1] && keys[3 * j + 2] == key[2]) found = j;
          if (found < 0) { keys[3 * n] = key[0]; keys[3 * n + 1] = key[1]; keys[3 * n + 2] = key[2]; }
        }
        if (found < 0) {
          pts[3 * n] = q[0]; pts[3 * n + 1] = q[1]; pts[3 * n + 2] = q[2];
          n++;
        }
      }
  }
  /* normalize_points: p / |p| * radius, stored float32 (utils.py:22-31) */
  for (int i = 0; i < n; i++) {
    double p[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    double nn = norm3(p);
    for (int k = 0; k < 3; k++) out[3 * i + k] = (float)(p[k] / nn * radius);
  }
  free(pts); free(keys);
  return n;
}

/* ------------------------------------------------------------------ */
/* O4  spherical spiral  src/py/utils.py:52-89 ; cxx:214-251           */
/* ------------------------------------------------------------------ */
ORC_API void orc_spiral(int N, double turns, double radius, float* out) {
  double c = 2.0 * turns;
  for (int i = 0; i < N; i++) {
    double angle = ((double)i * M_PI) / (double)N;
    out[3 * i + 0] = (float)(radius * sin(angle) * cos(c * angle));
    out[3 * i + 1] = (float)(radius * sin(angle) * sin(c * angle));
    out[3 * i + 2] = (float)(radius * cos(angle));
  }
}

/* ------------------------------------------------------------------ */
/* O5  view-plane basis and rays                                       */
/*   src/app/fly_by_features.cxx:652-702 (basis, vtkPlaneSource)       */
/*   src/app/fly_by_features.cxx:742-748 (segment per plane point)     */
/*   python twin src/py/predict.py:88-126                              */
/* ------------------------------------------------------------------ */
typedef struct {
  double sp[3], n[3], x[3], y[3];
  double origin[3], v1[3], v2[3]; /* vtkPlaneSource Origin, Point1-Origin, Point2-Origin */
  double delta[3];                /* sp - sp*spacing */
  double dir[3];                  /* end - origin = -(n*R)*2 */
} orc_view;

static void view_setup(const float spf[3], double radius, double plane_scale, double spacing,
                       orc_view* v) {
  for (int k = 0; k < 3; k++) v->sp[k] = spf[k];
  double len = norm3(v->sp);
  for (int k = 0; k < 3; k++) v->n[k] = v->sp[k] / len;
  int north = fabs(v->n[0]) <= 1e-8 && fabs(v->n[1]) <= 1e-8 && fabs(v->n[2] - 1.0) <= 1e-8;
  int south = fabs(v->n[0]) <= 1e-8 && fabs(v->n[1]) <= 1e-8 && fabs(v->n[2] + 1.0) <= 1e-8;
  if (north) {
    v->x[0] = 1; v->x[1] = 0; v->x[2] = 0; v->y[0] = 0; v->y[1] = 1; v->y[2] = 0;
  } else if (south) {
    v->x[0] = -1; v->x[1] = 0; v->x[2] = 0; v->y[0] = 0; v->y[1] = -1; v->y[2] = 0;
  } else {
    double up[3] = {0, 0, 1}, c[3];
    cross3(v->n, up, c);
    double l = norm3(c);
    for (int k = 0; k < 3; k++) v->x[k] = c[k] / l;
    cross3(v->n, v->x, c);
    l = norm3(c);
    for (int k = 0; k < 3; k++) v->y[k] = c[k] / l;
  }
  for (int k = 0; k < 3; k++) {
    double hx = v->x[k] * 0.5 * plane_scale, hy = v->y[k] * 0.5 * plane_scale;
    v->origin[k] = (v->sp[k] - hx) - hy;
    double p1 = v->origin[k] + v->x[k] * plane_scale;
    double p2 = v->origin[k] + v->y[k] * plane_scale;
    v->v1[k] = p1 - v->origin[k];
    v->v2[k] = p2 - v->origin[k];
    v->delta[k] = v->sp[k] - v->sp[k] * spacing;
    v->dir[k] = -((v->n[k] * radius) * 2.0);
  }
}

/* ray (i fastest, j slowest; vtkPlaneSource with resolution res-1) */
static void view_ray(const orc_view* v, int res, int i, int j, double spacing, double o[3]) {
  double tc0 = (res > 1) ? (double)i / (double)(res - 1) : 0.0;
  double tc1 = (res > 1) ? (double)j / (double)(res - 1) : 0.0;
  for (int k = 0; k < 3; k++) {
    float pf = (float)((v->origin[k] + tc0 * v->v1[k]) + tc1 * v->v2[k]); /* float vtkPoints */
    o[k] = (double)pf * spacing + v->delta[k];
  }
}

/* rays_out[res*res][6] = origin xyz, end xyz ; basis_out[9] = n, x, y */
ORC_API void orc_view_rays(const float* sp, double radius, double plane_scale, double spacing,
                           int res, double* rays_out, double* basis_out) {
  orc_view v;
  view_setup(sp, radius, plane_scale, spacing, &v);
  if (basis_out)
    for (int k = 0; k < 3; k++) {
      basis_out[k] = v.n[k]; basis_out[3 + k] = v.x[k]; basis_out[6 + k] = v.y[k];
    }
  if (rays_out)
    for (int j = 0; j < res; j++)
      for (int i = 0; i < res; i++) {
        double o[3];
        view_ray(&v, res, i, j, spacing, o);
        double* r = rays_out + 6 * ((size_t)j * res + i);
        for (int k = 0; k < 3; k++) { r[k] = o[k]; r[3 + k] = o[k] + v.dir[k]; }
      }
}

/* ------------------------------------------------------------------ */
/* O6  segment / triangle test, VTK semantics                          */
/*   call site src/app/fly_by_features.cxx:750-758 (tol 1e-10),        */
/*   src/py/predict.py:114-128 (tol 1e-8)                              */
/*   vtkTriangle::IntersectWithLine -> vtkPlane::IntersectWithLine ->  */
/*   vtkTriangle::EvaluatePosition (inclusive [0,1] weights); a point  */
/*   outside the triangle still counts when its squared distance to    */
/*   the triangle is <= tol^2 (this is what keeps a ray through a      */
/*   shared edge from falling between the two cells).                  */
/* Returns 1 on hit and fills t, x[3], w[3].                           */
/* ------------------------------------------------------------------ */
static double dist2pt(const double a[3], const double b[3]) {
  double d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
  return d0 * d0 + d1 * d1 + d2 * d2;
}
/* vtkLine::DistanceToLine: squared distance from x to the segment p1-p2 */
static double dist2seg(const double x[3], const double p1[3], const double p2[3]) {
  double p21[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double num = p21[0] * (x[0] - p1[0]) + p21[1] * (x[1] - p1[1]) + p21[2] * (x[2] - p1[2]);
  double denom = dot3(p21, p21);
  double tolerance = 1.e-05 * num; /* VTK_TOL */
  if (tolerance < 0.0) tolerance = -tolerance;
  double t;
  if (denom < tolerance) return dist2pt(p1, x); /* numerically bad: arbitrary end point */
  if ((t = num / denom) < 0.0) return dist2pt(p1, x);
  if (t > 1.0) return dist2pt(p2, x);
  double c[3] = {p1[0] + t * p21[0], p1[1] + t * p21[1], p1[2] + t * p21[2]};
  return dist2pt(c, x);
}
/* vtkTriangle::EvaluatePosition, point outside: squared distance to the nearest
 * edge / vertex, chosen from the signs of the weights (w[k] belongs to vertex k;
 * pt1 = P1, pt2 = P2, pt3 = P0) */
static double outside_dist2(const double x[3], const double pt1[3], const double pt2[3],
                            const double pt3[3], const double w[3]) {
  double dp, d1, d2, r;
  if (w[1] < 0.0 && w[2] < 0.0) {
    dp = dist2pt(x, pt3); d1 = dist2seg(x, pt1, pt3); d2 = dist2seg(x, pt3, pt2);
  } else if (w[2] < 0.0 && w[0] < 0.0) {
    dp = dist2pt(x, pt1); d1 = dist2seg(x, pt1, pt3); d2 = dist2seg(x, pt1, pt2);
  } else if (w[1] < 0.0 && w[0] < 0.0) {
    dp = dist2pt(x, pt2); d1 = dist2seg(x, pt2, pt3); d2 = dist2seg(x, pt1, pt2);
  } else if (w[0] < 0.0) return dist2seg(x, pt1, pt2);
  else if (w[1] < 0.0) return dist2seg(x, pt2, pt3);
  else if (w[2] < 0.0) return dist2seg(x, pt1, pt3);
  else return DBL_MAX; /* no negative weight: VTK leaves dist2 unset; treated as far */
  r = (dp < d1) ? dp : d1;
  if (d2 < r) r = d2;
  return r;
}

static double g_tol2 = 1.0e-20; /* tol^2, tol = 1e-10 (fly_by_features.cxx:750) */
ORC_API void orc_set_tolerance(double tol) { g_tol2 = tol * tol; }

static int tri_hit(const double o[3], const double d[3], const float* P0, const float* P1,
                   const float* P2, double* t_out, double x[3], double w[3]) {
  double pt1[3] = {P1[0], P1[1], P1[2]}, pt2[3] = {P2[0], P2[1], P2[2]}, pt3[3] = {P0[0], P0[1], P0[2]};
  /* ComputeNormalDirection(pt1, pt2, pt3): a = pt3 - pt2, b = pt1 - pt2, n = a x b */
  double a[3] = {pt3[0] - pt2[0], pt3[1] - pt2[1], pt3[2] - pt2[2]};
  double b[3] = {pt1[0] - pt2[0], pt1[1] - pt2[1], pt1[2] - pt2[2]};
  double n[3];
  cross3(a, b, n);
  if (n[0] == 0.0 && n[1] == 0.0 && n[2] == 0.0) return 0; /* degenerate: treated as miss */
  /* vtkPlane::IntersectWithLine(p1, p2, n, pt1, t, x), VTK_PLANE_TOL = 1e-6 */
  double num = dot3(n, pt1) - (n[0] * o[0] + n[1] * o[1] + n[2] * o[2]);
  double den = n[0] * d[0] + n[1] * d[1] + n[2] * d[2];
  double fabsden = den < 0.0 ? -den : den;
  double fabstol = num < 0.0 ? -num * 1.0e-06 : num * 1.0e-06;
  if (fabsden <= fabstol) return 0;
  double t = num / den;
  for (int k = 0; k < 3; k++) x[k] = o[k] + t * d[k];
  if (!(t >= 0.0 && t <= 1.0)) return 0;
  /* EvaluatePosition(x): project onto the plane (vtkPlane::GeneralizedProjectPoint) */
  double xo[3] = {x[0] - pt1[0], x[1] - pt1[1], x[2] - pt1[2]};
  double tp = dot3(n, xo), n2 = dot3(n, n);
  double cp[3];
  for (int k = 0; k < 3; k++) cp[k] = (n2 != 0.0) ? x[k] - tp * n[k] / n2 : x[k];
  /* drop the dominant axis of n (first max wins) */
  int idx = 0;
  double maxc = 0.0;
  for (int k = 0; k < 3; k++) {
    double f = n[k] < 0 ? -n[k] : n[k];
    if (f > maxc) { maxc = f; idx = k; }
  }
  int ia = (idx == 0) ? 1 : 0, ib = (idx == 2) ? 1 : 2;
  double rhs[2] = {cp[ia] - pt3[ia], cp[ib] - pt3[ib]};
  double c1[2] = {pt1[ia] - pt3[ia], pt1[ib] - pt3[ib]};
  double c2[2] = {pt2[ia] - pt3[ia], pt2[ib] - pt3[ib]};
  double det = c1[0] * c2[1] - c2[0] * c1[1];
  if (det == 0.0) return 0;
  double pc0 = (rhs[0] * c2[1] - c2[0] * rhs[1]) / det;
  double pc1 = (c1[0] * rhs[1] - rhs[0] * c1[1]) / det;
  w[0] = 1 - (pc0 + pc1); w[1] = pc0; w[2] = pc1;
  if (w[0] >= 0.0 && w[0] <= 1.0 && w[1] >= 0.0 && w[1] <= 1.0 && w[2] >= 0.0 && w[2] <= 1.0) {
    *t_out = t;
    return 1;
  }
  /* outside: vtkTriangle::IntersectWithLine still reports a hit when dist2 <= tol^2 */
  if (outside_dist2(x, pt1, pt2, pt3, w) <= g_tol2) {
    *t_out = t;
    return 1;
  }
  return 0;
}

ORC_API int orc_tri_intersect(const double* o, const double* e, const float* p0, const float* p1,
                              const float* p2, double* t, double* x, double* w) {
  double d[3] = {e[0] - o[0], e[1] - o[1], e[2] - o[2]};
  return tri_hit(o, d, p0, p1, p2, t, x, w);
}

/* nearest hit by exhaustive search: min t, ties -> lowest cell id (strict <
 * while scanning ascending ids == vtkCellLocator "first found wins") */
static int cast_brute(const double o[3], const double d[3], const float* verts,
                      const int32_t* tris, int64_t T, double* t, double x[3], double w[3]) {
  int best = -1;
  double bt = DBL_MAX;
  for (int64_t c = 0; c < T; c++) {
    double tt, xx[3], ww[3];
    if (tri_hit(o, d, verts + 3 * (int64_t)tris[3 * c], verts + 3 * (int64_t)tris[3 * c + 1],
                verts + 3 * (int64_t)tris[3 * c + 2], &tt, xx, ww) && tt < bt) {
      bt = tt; best = (int)c;
      memcpy(x, xx, sizeof xx); memcpy(w, ww, sizeof ww);
    }
  }
  *t = bt;
  return best;
}

/* ------------------------------------------------------------------ */
/* vtkCellLocator-like uniform bucket grid (a6/a7)                     */
/*   src/app/fly_by_features.cxx:484-486 (BuildLocator)                */
/*   VTK defaults: 25 cells/bucket, MaxLevel 8, Automatic ->           */
/*   level = ceil(log8(T/25)), 2^level divisions per axis.             */
/* ------------------------------------------------------------------ */
typedef struct {
  int nd;              /* divisions per axis */
  double lo[3], h[3];  /* grid origin and bucket size */
  int64_t* start;      /* nd^3 + 1 */
  int32_t* items;      /* cell ids, ascending within a bucket */
  const float* verts;
  const int32_t* tris;
  int64_t T;
} orc_grid;

ORC_API void orc_grid_free(orc_grid* g) {
  if (!g) return;
  free(g->start); free(g->items); free(g);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* divisions <= 0 -> VTK automatic rule */
ORC_API orc_grid* orc_grid_build(const float* verts, int64_t V, const int32_t* tris, int64_t T,
                                 int divisions) {
  orc_grid* g = (orc_grid*)calloc(1, sizeof(orc_grid));
  g->verts = verts; g->tris = tris; g->T = T;
  if (divisions <= 0) {
    int level = (int)ceil(log((double)T / 25.0) / log(8.0));
    if (level < 0) level = 0;
    if (level > 8) level = 8;
    divisions = 1 << level;
  }
  g->nd = divisions;
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int64_t i = 0; i < V; i++)
    for (int k = 0; k < 3; k++) {
      double p = verts[3 * i + k];
      if (p < lo[k]) lo[k] = p;
      if (p > hi[k]) hi[k] = p;
    }
  for (int k = 0; k < 3; k++) {
    double pad = (hi[k] - lo[k]) * 1e-6 + 1e-9;
    g->lo[k] = lo[k] - pad;
    g->h[k] = ((hi[k] + pad) - g->lo[k]) / g->nd;
  }
  int nd = g->nd;
  int64_t nb = (int64_t)nd * nd * nd;
  g->start = (int64_t*)calloc((size_t)nb + 1, sizeof(int64_t));
  for (int pass = 0; pass < 2; pass++) {
    int64_t* fill = NULL;
    if (pass == 1) {
      int64_t acc = 0;
      for (int64_t b = 0; b < nb; b++) { int64_t c = g->start[b]; g->start[b] = acc; acc += c; }
      g->start[nb] = acc;
      g->items = (int32_t*)malloc(sizeof(int32_t) * (size_t)(acc > 0 ? acc : 1));
      fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)nb);
      memcpy(fill, g->start, sizeof(int64_t) * (size_t)nb);
    }
    for (int64_t c = 0; c < T; c++) {
      int r0[3], r1[3];
      for (int k = 0; k < 3; k++) {
        double a = verts[3 * (int64_t)tris[3 * c] + k], b = verts[3 * (int64_t)tris[3 * c + 1] + k],
               e = verts[3 * (int64_t)tris[3 * c + 2] + k];
        double mn = fmin(a, fmin(b, e)), mx = fmax(a, fmax(b, e));
        /* one-ulp-ish slack so a hit point on a bucket face is never lost */
        r0[k] = clampi((int)floor((mn - g->lo[k]) / g->h[k] - 1e-7), 0, nd - 1);
        r1[k] = clampi((int)floor((mx - g->lo[k]) / g->h[k] + 1e-7), 0, nd - 1);
      }
      for (int z = r0[2]; z <= r1[2]; z++)
        for (int y = r0[1]; y <= r1[1]; y++)
          for (int xx = r0[0]; xx <= r1[0]; xx++) {
            int64_t b = ((int64_t)z * nd + y) * nd + xx;
            if (pass == 0) g->start[b]++;
            else g->items[fill[b]++] = (int32_t)c;
          }
    }
    free(fill);
  }
  return g;
}

/* Walk buckets front to back (3-D DDA), test every cell listed in a bucket,
 * keep lexicographic min (t, cell id); stop once the best t lies before the
 * exit of the current bucket.  Equals cast_brute on every ray (tested). */
static int cast_grid(const orc_grid* g, const double o[3], const double d[3], double* t,
                     double x[3], double w[3]) {
  int nd = g->nd;
  /* clip [0,1] against the grid box */
  double t0 = 0.0, t1 = 1.0;
  for (int k = 0; k < 3; k++) {
    double lo = g->lo[k], hi = g->lo[k] + g->h[k] * nd;
    if (d[k] == 0.0) {
      if (o[k] < lo || o[k] > hi) { *t = DBL_MAX; return -1; }
    } else {
      double ta = (lo - o[k]) / d[k], tb = (hi - o[k]) / d[k];
      if (ta > tb) { double s = ta; ta = tb; tb = s; }
      if (ta > t0) t0 = ta;
      if (tb < t1) t1 = tb;
    }
  }
  if (t0 > t1) { *t = DBL_MAX; return -1; }
  int c[3], step[3];
  double tnext[3], tdelta[3];
  for (int k = 0; k < 3; k++) {
    double p = o[k] + t0 * d[k];
    c[k] = clampi((int)floor((p - g->lo[k]) / g->h[k]), 0, nd - 1);
    if (d[k] > 0) {
      step[k] = 1;
      tnext[k] = (g->lo[k] + (c[k] + 1) * g->h[k] - o[k]) / d[k];
      tdelta[k] = g->h[k] / d[k];
    } else if (d[k] < 0) {
      step[k] = -1;
      tnext[k] = (g->lo[k] + c[k] * g->h[k] - o[k]) / d[k];
      tdelta[k] = -g->h[k] / d[k];
    } else {
      step[k] = 0; tnext[k] = DBL_MAX; tdelta[k] = DBL_MAX;
    }
  }
  int best = -1;
  double bt = DBL_MAX;
  for (;;) {
    int64_t b = ((int64_t)c[2] * nd + c[1]) * nd + c[0];
    for (int64_t q = g->start[b]; q < g->start[b + 1]; q++) {
      int32_t cid = g->items[q];
      double tt, xx[3], ww[3];
      if (tri_hit(o, d, g->verts + 3 * (int64_t)g->tris[3 * cid],
                  g->verts + 3 * (int64_t)g->tris[3 * cid + 1],
                  g->verts + 3 * (int64_t)g->tris[3 * cid + 2], &tt, xx, ww)) {
        if (tt < bt || (tt == bt && cid < best)) {
          bt = tt; best = cid;
          memcpy(x, xx, sizeof xx); memcpy(w, ww, sizeof ww);
        }
      }
    }
    int ax = (tnext[0] <= tnext[1]) ? ((tnext[0] <= tnext[2]) ? 0 : 2) : ((tnext[1] <= tnext[2]) ? 1 : 2);
    double texit = tnext[ax];
    /* a hit strictly inside the walked span is final (small slack for the face) */
    if (best >= 0 && bt < texit - 1e-9) break;
    if (texit > t1 + 1e-9) break;
    c[ax] += step[ax];
    if (c[ax] < 0 || c[ax] >= nd) break;
    tnext[ax] += tdelta[ax];
  }
  *t = bt;
  return best;
}

/* Generic segment casting: rays[n][6] (origin, end).  grid may be NULL -> brute force. */
ORC_API void orc_cast(const orc_grid* g, const float* verts, const int32_t* tris, int64_t T,
                      const double* rays, int64_t n, int32_t* ids, double* t_out, double* x_out,
                      double* w_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n; r++) {
    const double* o = rays + 6 * r;
    double d[3] = {o[3] - o[0], o[4] - o[1], o[5] - o[2]};
    double t, x[3] = {0, 0, 0}, w[3] = {0, 0, 0};
    int id = g ? cast_grid(g, o, d, &t, x, w) : cast_brute(o, d, verts, tris, T, &t, x, w);
    ids[r] = id;
    if (t_out) t_out[r] = id >= 0 ? t : -1.0;
    if (x_out) for (int k = 0; k < 3; k++) x_out[3 * r + k] = x[k];
    if (w_out) for (int k = 0; k < 3; k++) w_out[3 * r + k] = w[k];
  }
}

/* ------------------------------------------------------------------ */
/* O7 + O8  shading and channel assembly                               */
/*   src/app/fly_by_features.cxx:766-807 (bary normal, z, pixel)       */
/*   src/py/fly_by_features.py:124-150 (use_z / split_z / rescale)     */
/* normal_encoding: 0 unit01 (n*0.5+0.5), 1 byte255 ((n*0.5+0.5)*255), */
/*                  2 signed (n; rescale_features)                     */
/* Channels: use_z=0 -> 3 ; use_z=1,split_z=1 -> 4 ; use_z=1,split_z=0 */
/* -> 3 (normals * depth, cxx:802-804).  zscale multiplies depth        */
/* (1/z_far when rescaling, fly_by_features.py:138-146).  Miss -> 0.   */
/* ------------------------------------------------------------------ */
ORC_API int orc_channels(int use_z, int split_z) { return (use_z && split_z) ? 4 : 3; }

static void shade_z(const double w[3], const float* normals, const int32_t* tri, double z, int use_z, int split_z,
                    int enc, float* px) {
  double nn[3] = {0, 0, 0};
  for (int j = 0; j < 3; j++) {
    const float* N = normals + 3 * (int64_t)tri[j];
    for (int k = 0; k < 3; k++) nn[k] = nn[k] + (double)N[k] * w[j];
  }
  double e[3];
  for (int k = 0; k < 3; k++) {
    if (enc == 2) e[k] = nn[k];
    else if (enc == 1) e[k] = (nn[k] * 0.5 + 0.5) * 255.0;
    else if (enc == 3) { /* the uint8 colour an OpenGL RGB read-back returns */
      double q = floor((nn[k] * 0.5 + 0.5) * 255.0 + 0.5);
      e[k] = q < 0.0 ? 0.0 : (q > 255.0 ? 255.0 : q);
    } else e[k] = nn[k] * 0.5 + 0.5;
  }
  if (!use_z) { for (int k = 0; k < 3; k++) px[k] = (float)e[k]; }
  else if (split_z) { for (int k = 0; k < 3; k++) px[k] = (float)e[k]; px[3] = (float)z; }
  else { for (int k = 0; k < 3; k++) px[k] = (float)(e[k] * z); }
}

static void shade(const double o[3], const double x[3], const double w[3], const float* normals,
                  const int32_t* tri, int use_z, int split_z, int enc, double zscale, float* px) {
  double dd[3] = {o[0] - x[0], o[1] - x[1], o[2] - x[2]};
  shade_z(w, normals, tri, norm3(dd) * zscale, use_z, split_z, enc, px);
}

/* Whole path for a set of viewpoints.  verts are the UNIT-SURF float32 points,
 * normals the O2 point normals.  out_features [Vw,res,res,C], out_ids [Vw,res,res]
 * (-1 miss), out_t optional [Vw,res,res] (-1 miss).  use_grid: 0 brute force,
 * >0 bucket grid (divisions = grid_div, <=0 VTK automatic). */
ORC_API void orc_render(const float* verts, int64_t V, const int32_t* tris, int64_t T,
                        const float* normals, const float* views, int n_views, double radius,
                        int res, double plane_scale, double spacing, int use_z, int split_z,
                        int enc, double zscale, int use_grid, int grid_div, float* out_features,
                        int32_t* out_ids, double* out_t) {
  int C = orc_channels(use_z, split_z);
  orc_grid* g = use_grid ? orc_grid_build(verts, V, tris, T, grid_div) : NULL;
  for (int v = 0; v < n_views; v++) {
    orc_view vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = 0; p < (int64_t)res * res; p++) {
      int i = (int)(p % res), j = (int)(p / res);
      double o[3], t, x[3] = {0, 0, 0}, w[3] = {0, 0, 0};
      view_ray(&vw, res, i, j, spacing, o);
      int id = g ? cast_grid(g, o, vw.dir, &t, x, w) : cast_brute(o, vw.dir, verts, tris, T, &t, x, w);
      size_t pix = (size_t)v * res * res + (size_t)p;
      if (out_ids) out_ids[pix] = id;
      if (out_t) out_t[pix] = id >= 0 ? t : -1.0;
      if (out_features) {
        float* px = out_features + pix * C;
        for (int k = 0; k < C; k++) px[k] = 0.f;
        if (id >= 0) shade(o, x, w, normals, tris + 3 * (int64_t)id, use_z, split_z, enc, zscale, px);
      }
    }
  }
  orc_grid_free(g);
}

/* ------------------------------------------------------------------ */
/* Raster-compatible views                          */
/*   src/py/fly_by_features.py:85-107 camera: position = viewpoint,    */
/*   focal point = origin, view-up (0,0,-1) or (+-1,0,0) at the poles; */
/*   vtkCamera default view angle 30; :113-131 RGB + linearised z.     */
/* One ray through every pixel centre, rows bottom to top; the ray     */
/* parameter is the depth along the viewing direction.                 */
/* ------------------------------------------------------------------ */
ORC_API void orc_render_perspective(const float* verts, int64_t V, const int32_t* tris, int64_t T,
                                    const float* normals, const float* views, int n_views, int res,
                                    double view_angle_deg, double z_near, double z_far, int use_z, int split_z,
                                    int enc, double zscale, int use_grid, int grid_div, float* out_features,
                                    int32_t* out_ids, double* out_depth) {
  int C = orc_channels(use_z, split_z);
  orc_grid* g = use_grid ? orc_grid_build(verts, V, tris, T, grid_div) : NULL;
  const double tan_half = tan(view_angle_deg * 0.5 * 3.14159265358979323846 / 180.0);
  for (int v = 0; v < n_views; v++) {
    const float* sp = views + 3 * v;
    double eye[3] = {sp[0], sp[1], sp[2]};
    double len = sqrt((double)sp[0] * (double)sp[0] + (double)sp[1] * (double)sp[1] + (double)sp[2] * (double)sp[2]);
    double nz = (double)sp[2] / len;
    double up[3] = {0.0, 0.0, -1.0};
    if (nz == 1.0) { up[0] = 1.0; up[2] = 0.0; }
    else if (nz == -1.0) { up[0] = -1.0; up[2] = 0.0; }
    double f[3] = {-(double)sp[0] / len, -(double)sp[1] / len, -(double)sp[2] / len};
    double s[3], u[3];
    cross3(f, up, s);
    double sl = norm3(s);
    for (int k = 0; k < 3; k++) s[k] = s[k] / sl;
    cross3(s, f, u);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t p = 0; p < (int64_t)res * res; p++) {
      int i = (int)(p % res), j = (int)(p / res);
      double sx = ((((double)i + 0.5) / (double)res) * 2.0 - 1.0) * tan_half;
      double sy = ((((double)j + 0.5) / (double)res) * 2.0 - 1.0) * tan_half;
      double span = z_far - z_near;
      double o[3], d[3], t, x[3] = {0, 0, 0}, w[3] = {0, 0, 0};
      for (int k = 0; k < 3; k++) {
        double dir = (f[k] + sx * s[k]) + sy * u[k];
        o[k] = eye[k] + dir * z_near;
        d[k] = dir * span;
      }
      int id = g ? cast_grid(g, o, d, &t, x, w) : cast_brute(o, d, verts, tris, T, &t, x, w);
      size_t pix = (size_t)v * res * res + (size_t)p;
      double z = id >= 0 ? z_near + t * span : -1.0;
      if (out_ids) out_ids[pix] = id;
      if (out_depth) out_depth[pix] = z;
      if (out_features) {
        float* px = out_features + pix * C;
        for (int k = 0; k < C; k++) px[k] = 0.f;
        if (id >= 0) shade_z(w, normals, tris + 3 * (int64_t)id, z * zscale, use_z, split_z, enc, px);
      }
    }
  }
  orc_grid_free(g);
}

/* ------------------------------------------------------------------ */
/* a10  point-feature lookup through the id map                        */
/*   src/py/utils.py:604-621 (cell colour = id of the cell's vertex 0) */
/*   src/py/utils.py:639-684 (ExtractPointFeatures; miss -> zero)      */
/* ------------------------------------------------------------------ */
ORC_API void orc_point_features(const int32_t* cell_ids, int64_t n_pix, const int32_t* tris,
                                const float* feats, int F, float zero, float* out) {
  for (int64_t p = 0; p < n_pix; p++) {
    int32_t c = cell_ids[p];
    for (int k = 0; k < F; k++)
      out[p * F + k] = (c >= 0) ? feats[(int64_t)tris[3 * (int64_t)c] * F + k] : zero;
  }
}

/* Barycentric lookup at the hits: out[pixel, f] = w0 f0 + w1 f1 + w2 f2 with the weights of
 * vtkTriangle::EvaluatePosition for the cell the render reported -- how the C++ tool interpolates its
 * normals (src/app/fly_by_features.cxx:766-790) and what the vertex-colour shading of
 * src/py/challenge-teeth/teeth_nets.py:121-153 computes.  Products and sums in double, left to right. */
ORC_API void orc_interpolate_features(const float* verts, const int32_t* tris, const float* views, int n_views,
                                      double radius, int res, double plane_scale, double spacing,
                                      const int32_t* cell_ids, const float* feats, int F, float zero, float* out) {
  for (int v = 0; v < n_views; v++) {
    orc_view vw;
    view_setup(views + 3 * v, radius, plane_scale, spacing, &vw);
    for (int64_t p = 0; p < (int64_t)res * res; p++) {
      size_t pix = (size_t)v * res * res + (size_t)p;
      int32_t c = cell_ids[pix];
      double o[3], t, x[3], w[3];
      int hit = 0;
      if (c >= 0) {
        const int32_t* tr = tris + 3 * (int64_t)c;
        view_ray(&vw, res, (int)(p % res), (int)(p / res), spacing, o);
        hit = tri_hit(o, vw.dir, verts + 3 * (int64_t)tr[0], verts + 3 * (int64_t)tr[1], verts + 3 * (int64_t)tr[2],
                      &t, x, w);
      }
      for (int k = 0; k < F; k++) {
        if (!hit) { out[pix * F + k] = zero; continue; }
        const int32_t* tr = tris + 3 * (int64_t)c;
        double a = w[0] * (double)feats[(int64_t)tr[0] * F + k];
        double b = w[1] * (double)feats[(int64_t)tr[1] * F + k];
        double e = w[2] * (double)feats[(int64_t)tr[2] * F + k];
        out[pix * F + k] = (float)((a + b) + e);
      }
    }
  }
}

/* ------------------------------------------------------------------ */
/* O9  back-projection of per-pixel labels onto cells                  */
/*   src/py/dental_model_seg.py:192-199 (per cell, through pix_to_face)*/
/*   src/py/predict_v3.py:59-80 ; src/py/predict.py:154-169            */
/* Miss pixels contribute nothing (deliberate fix of the -1 row bug).  */
/* ------------------------------------------------------------------ */
ORC_API void orc_backproject(const int32_t* cell_ids, const int32_t* labels, int64_t n_pix, int K,
                             int32_t* votes) {
  for (int64_t p = 0; p < n_pix; p++) {
    int32_t c = cell_ids[p], l = labels[p];
    if (c >= 0 && l >= 0 && l < K) votes[(int64_t)c * K + l] += 1;
  }
}

/* numpy argmax semantics (lowest label wins ties); no votes -> dflt
 * (predict_v3.py:75-80, predict.py:164-169) */
ORC_API void orc_votes_argmax(const int32_t* votes, int64_t T, int K, int32_t dflt, int32_t* out) {
  for (int64_t c = 0; c < T; c++) {
    int32_t best = 0, arg = -1;
    for (int k = 0; k < K; k++)
      if (votes[c * K + k] > best) { best = votes[c * K + k]; arg = k; }
    out[c] = arg >= 0 ? arg : dflt;
  }
}

/* cell label -> the cell's vertex 0, later cells overwrite earlier ones
 * (src/py/dental_model_seg.py:208-211) */
ORC_API void orc_cells_to_points(const int32_t* cell_labels, const int32_t* tris, int64_t T,
                                 int64_t V, int32_t dflt, int32_t* out) {
  for (int64_t i = 0; i < V; i++) out[i] = dflt;
  for (int64_t c = 0; c < T; c++) out[tris[3 * c]] = cell_labels[c];
}
