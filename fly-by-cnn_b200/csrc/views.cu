// views.cu -- viewpoint generation on the GPU.
//
// Icosahedron-subdivision viewpoints: CreateIcosahedron (reference
// src/py/utils.py:36-50) = vtkPlatonicSolidSource + LinearSubdivisionFilter
// (src/py/LinearSubdivisionFilter.py:42-67: per face, s13 outer / s12 inner,
// first-insertion order, duplicate lattice points dropped) + normalize_points
// (src/py/utils.py:22-31).  Spherical spiral: CreateSpiral
// (src/py/utils.py:52-89; src/app/fly_by_features.cxx:214-251).
//
// The lattice is tiny (20 (L+1)(L+2)/2 candidates); one thread per candidate,
// one scan for the first-insertion ranks.
#include "ctx.h"

namespace fb {

// vtkPlatonicSolidSource icosahedron (VTK 9, restated; un-vendored dependency)
__constant__ double c_ico_pts[12][3] = {
    {0.0, 0.0, -1.0},           {0.0, 0.894427, -0.447214},      {0.850651, 0.276393, -0.447214},
    {0.525731, -0.723607, -0.447214}, {-0.525731, -0.723607, -0.447214}, {-0.850651, 0.276393, -0.447214},
    {0.0, -0.894427, 0.447214}, {0.850651, -0.276393, 0.447214}, {0.525731, 0.723607, 0.447214},
    {-0.525731, 0.723607, 0.447214}, {-0.850651, -0.276393, 0.447214}, {0.0, 0.0, 1.0}};
__constant__ int c_ico_faces[20][3] = {
    {0, 5, 4},  {0, 4, 3},  {0, 3, 2},  {0, 2, 1},   {0, 1, 5},  {5, 1, 9},  {5, 9, 10},
    {5, 10, 4}, {4, 10, 6}, {4, 6, 3},  {3, 6, 7},   {3, 7, 2},  {2, 7, 8},  {2, 8, 1},
    {1, 8, 9},  {9, 8, 11}, {9, 11, 10}, {10, 11, 6}, {6, 11, 7}, {7, 11, 8}};
#define FB_ICO_SOLID_SCALE (1.0 / 0.58778524999243)

__device__ __forceinline__ bool face_has(int f, int v) {
  return c_ico_faces[f][0] == v || c_ico_faces[f][1] == v || c_ico_faces[f][2] == v;
}

// candidate index -> (face, s13, s12) in the reference's insertion order
__device__ __forceinline__ void lattice_decode(int cand, int L, int* f, int* s13, int* s12) {
  int per = (L + 1) * (L + 2) / 2;
  *f = cand / per;
  int r = cand % per;
  int row = 0;
  while (r >= L + 1 - row) {
    r -= L + 1 - row;
    row++;
  }
  *s13 = row;
  *s12 = r;
}

__global__ void ico_lattice_kernel(int L, int n_cand, int dedupe, float* __restrict__ cand_pts,
                                   uint32_t* __restrict__ first_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  int f, s13, s12;
  lattice_decode(i, L, &f, &s13, &s12);
  const int* fv = c_ico_faces[f];
  float q[3];
  for (int k = 0; k < 3; k++) {
    // float vtkPoints of the source, then double interpolation, then float storage
    double p1 = (double)(float)FB_MUL(c_ico_pts[fv[0]][k], FB_ICO_SOLID_SCALE);
    double p2 = (double)(float)FB_MUL(c_ico_pts[fv[1]][k], FB_ICO_SOLID_SCALE);
    double p3 = (double)(float)FB_MUL(c_ico_pts[fv[2]][k], FB_ICO_SOLID_SCALE);
    double d12 = FB_DIV(FB_SUB(p2, p1), (double)L);
    double d13 = FB_DIV(FB_SUB(p3, p1), (double)L);
    q[k] = (float)FB_ADD(FB_ADD(p1, FB_MUL((double)s12, d12)), FB_MUL((double)s13, d13));
    cand_pts[3 * i + k] = q[k];
  }
  if (dedupe == 0) {
    // exact lattice identity: a point with 1 (vertex) or 2 (edge) non-zero
    // weights is first-inserted by the lowest face that contains it
    int w[3] = {L - s12 - s13, s12, s13};
    int nz = (w[0] > 0) + (w[1] > 0) + (w[2] > 0);
    bool first = true;
    if (nz < 3) {
      int u = -1, v = -1;
      for (int a = 0; a < 3; a++)
        if (w[a] > 0) { if (u < 0) u = fv[a]; else v = fv[a]; }
      for (int g = 0; g < f && first; g++)
        if (face_has(g, u) && (v < 0 || face_has(g, v))) first = false;
    }
    first_flag[i] = first ? 1u : 0u;
  }
}

// float32-equality emulation of vtkIncrementalOctreePointLocator (dedupe = 1)
__global__ void ico_dedupe_float_kernel(int n_cand, const float* __restrict__ cand_pts,
                                        uint32_t* __restrict__ first_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand) return;
  float x = cand_pts[3 * i], y = cand_pts[3 * i + 1], z = cand_pts[3 * i + 2];
  bool first = true;
  for (int j = 0; j < i && first; j++)
    if (cand_pts[3 * j] == x && cand_pts[3 * j + 1] == y && cand_pts[3 * j + 2] == z) first = false;
  first_flag[i] = first ? 1u : 0u;
}

__global__ void ico_emit_kernel(int n_cand, const float* __restrict__ cand_pts,
                                const uint32_t* __restrict__ flag, const uint32_t* __restrict__ rank,
                                double radius, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cand || !flag[i]) return;
  double p[3] = {(double)cand_pts[3 * i], (double)cand_pts[3 * i + 1], (double)cand_pts[3 * i + 2]};
  double nn = norm3(p);
  for (int k = 0; k < 3; k++) out[3 * rank[i] + k] = (float)FB_MUL(FB_DIV(p[k], nn), radius);
}

int views_icosahedron(flyby_ctx* c, int L, double radius, int dedupe) {
  if (L < 1 || L > 1000) return fail(c, FLYBY_ERR_INVALID, "subdivision must be in [1,1000]");
  if (dedupe == 1 && L > 64)
    return fail(c, FLYBY_ERR_UNSUPPORTED, "float32 de-dup emulation is limited to subdivision <= 64");
  cudaStream_t st = c->stream;
  const int n_cand = 20 * (L + 1) * (L + 2) / 2;
  FB_CUDA(c, c->stage_in0.reserve(sizeof(float) * 3 * (size_t)n_cand));   // candidate points
  FB_CUDA(c, c->stage_in1.reserve(sizeof(uint32_t) * 2 * (size_t)n_cand + 64));  // flags + ranks
  FB_CUDA(c, c->view_pts.reserve(sizeof(float) * 3 * (size_t)n_cand));
  float* cand = c->stage_in0.as<float>();
  uint32_t* flag = c->stage_in1.as<uint32_t>();
  uint32_t* rank = flag + n_cand;
  const unsigned g = (unsigned)cdiv(n_cand, 128);
  ico_lattice_kernel<<<g, 128, 0, st>>>(L, n_cand, dedupe, cand, flag);
  FB_LAUNCH_CHECK(c);
  if (dedupe == 1) {
    ico_dedupe_float_kernel<<<g, 128, 0, st>>>(n_cand, cand, flag);
    FB_LAUNCH_CHECK(c);
  }
  FB_CUDA(c, cudaMemcpyAsync(rank, flag, sizeof(uint32_t) * (size_t)n_cand, cudaMemcpyDeviceToDevice, st));
  int rc = exclusive_scan_u32(c, rank, n_cand, c->mesh.scan_tmp);
  if (rc) return rc;
  ico_emit_kernel<<<g, 128, 0, st>>>(n_cand, cand, flag, rank, radius, c->view_pts.as<float>());
  FB_LAUNCH_CHECK(c);
  uint32_t last_rank = 0, last_flag = 0;
  FB_CUDA(c, cudaMemcpyAsync(&last_rank, rank + n_cand - 1, 4, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaMemcpyAsync(&last_flag, flag + n_cand - 1, 4, cudaMemcpyDeviceToHost, st));
  FB_CUDA(c, cudaStreamSynchronize(st));
  c->n_views = (int)(last_rank + last_flag);
  c->radius = radius;
  return FLYBY_OK;
}

// CreateSpiral (utils.py:52-89) calls libm's sin / cos from Python; device sin / cos may differ from
// libm in the last bit, and one float32 ulp of a viewpoint moves every plane point of that view.  The N
// points (64 in the reference's callers) are therefore computed on the host with libm, in the
// reference's expression order, and uploaded: bit-equal to the reference (tests/golden/ref_views.npz).
int views_spiral(flyby_ctx* c, int n, double turns, double radius) {
  if (n < 1) return fail(c, FLYBY_ERR_INVALID, "spiral needs at least one sample");
  FB_CUDA(c, c->view_pts.reserve(sizeof(float) * 3 * (size_t)n));
  const double pi = 3.141592653589793;  // math.pi
  const double cturn = 2.0 * turns;
  float* h = (float*)malloc(sizeof(float) * 3 * (size_t)n);
  if (!h) return fail(c, FLYBY_ERR_INVALID, "spiral: out of host memory");
  for (int i = 0; i < n; i++) {
    volatile double angle = ((double)i * pi) / (double)n;
    volatile double sa = sin(angle), ca = cos(angle);
    volatile double ta = cturn * angle;
    volatile double x = radius * sa, y = radius * sa;
    x = x * cos(ta);
    y = y * sin(ta);
    volatile double z = radius * ca;
    h[3 * i + 0] = (float)x;
    h[3 * i + 1] = (float)y;
    h[3 * i + 2] = (float)z;
  }
  cudaError_t e = cudaMemcpyAsync(c->view_pts.p, h, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  free(h);
  FB_CUDA(c, e);
  c->n_views = n;
  c->radius = radius;
  return FLYBY_OK;
}

}  // namespace fb
