/* postproc_oracle.c -- CPU restatement of the reference's mesh-label post-processing.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/flyby_oracle.c): the product never links or loads this.
 *
 * Follows, loop for loop, src/py/post_process.py of the reference (the step every predictor runs right
 * after the vote back-projection: predict_v3.py:93-133, dental_model_seg.py:228-239, predict.py:190-201):
 *   ConnectedRegion       :243-255   (+ GetNeighborIds, src/py/utils.py:233-247)
 *   NeighborLabel         :257-281
 *   RemoveIslands         :285-295
 *   ConnectivityLabeling  :297-304
 *   ErodeLabel            :306-346   (+ GetNeighbors, src/py/utils.py:218-231)
 *   DilateLabel           :348-375
 *   ReLabel               :377-380
 * The reference walks vtkPolyData links (GetPointCells / GetCellPoints); here the links are a CSR of the
 * incident triangles of every point, cells in ascending id as vtkCellLinks builds them.  Labels are the
 * integer values the reference stores in a vtk array and reads back with GetTuple(pid)[0].
 *
 * Parity status: "unpinned" like the rest of the oracle -- the reference has no tests or fixtures for
 * these functions and cannot run here (vtk is not installed); the restatement is literal, including the
 * reference's quirk that RemoveIslands relabels nothing unless ignore_neg1 is true (:292).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

typedef struct {
  int64_t V, T;
  const int32_t* tris;
  int64_t* start; /* [V+1] */
  int32_t* cells; /* [3T] incident cells of every point, ascending */
} links_t;

static void links_build(links_t* L, const int32_t* tris, int64_t T, int64_t V) {
  L->V = V; L->T = T; L->tris = tris;
  L->start = (int64_t*)calloc((size_t)V + 1, sizeof(int64_t));
  L->cells = (int32_t*)malloc(sizeof(int32_t) * 3 * (size_t)(T > 0 ? T : 1));
  for (int64_t c = 0; c < T; c++)
    for (int k = 0; k < 3; k++) L->start[tris[3 * c + k] + 1]++;
  for (int64_t v = 0; v < V; v++) L->start[v + 1] += L->start[v];
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(V > 0 ? V : 1));
  memcpy(cur, L->start, sizeof(int64_t) * (size_t)V);
  for (int64_t c = 0; c < T; c++) /* ascending c => every list ascending */
    for (int k = 0; k < 3; k++) {
      int32_t p = tris[3 * c + k];
      /* a cell that names a point twice is linked once per mention by vtk as well */
      L->cells[cur[p]++] = (int32_t)c;
    }
  free(cur);
}
static void links_free(links_t* L) { free(L->start); free(L->cells); }

static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}
/* np.unique of a small list: sort + drop repeats, returns the new length */
static int unique_i32(int32_t* a, int n) {
  if (n == 0) return 0;
  qsort(a, (size_t)n, sizeof(int32_t), cmp_i32);
  int m = 1;
  for (int i = 1; i < n; i++) if (a[i] != a[m - 1]) a[m++] = a[i];
  return m;
}

typedef struct { int32_t* p; int64_t n, cap; } vec_t;
static void vec_push(vec_t* v, int32_t x) {
  if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 64; v->p = (int32_t*)realloc(v->p, sizeof(int32_t) * (size_t)v->cap); }
  v->p[v->n++] = x;
}

/* utils.py:233-247 -- same-label unvisited neighbours of pid through its cells; marks them visited */
static void neighbor_ids(const links_t* L, int32_t pid, const int32_t* labels, int32_t label, uint8_t* visited, vec_t* out) {
  int64_t n0 = out->n;
  for (int64_t q = L->start[pid]; q < L->start[pid + 1]; q++) {
    const int32_t* cell = L->tris + 3 * (int64_t)L->cells[q];
    for (int k = 0; k < 3; k++) {
      int32_t pi = cell[k];
      if (labels[pi] == label && pi != pid && !visited[pi]) { visited[pi] = 1; vec_push(out, pi); }
    }
  }
  (void)n0; /* np.unique only sorts: every id was appended once because of the visited mark */
}

/* post_process.py:243-255 -- returns the sorted unique point ids of the region grown from pid */
static void connected_region(const links_t* L, int32_t pid, const int32_t* labels, int32_t label, uint8_t* visited, vec_t* region) {
  vec_t stack = {0, 0, 0};
  region->n = 0;
  neighbor_ids(L, pid, labels, label, visited, &stack);
  vec_push(region, pid);
  for (int64_t i = 0; i < stack.n; i++) vec_push(region, stack.p[i]);
  while (stack.n) {
    int32_t npid = stack.p[--stack.n];
    int64_t before = stack.n;
    neighbor_ids(L, npid, labels, label, visited, &stack);
    for (int64_t i = before; i < stack.n; i++) vec_push(region, stack.p[i]);
  }
  region->n = unique_i32(region->p, (int)region->n);
  free(stack.p);
}

/* post_process.py:257-281 -- most frequent label among the unique points around the region whose label
 * differs; max(list, key=list.count) returns the first element (ascending point id) that reaches the
 * maximum count; -1 when there is none */
static int32_t neighbor_label(const links_t* L, const int32_t* labels, int32_t label, const vec_t* region) {
  vec_t nb = {0, 0, 0};
  for (int64_t r = 0; r < region->n; r++) {
    int32_t pid = region->p[r];
    for (int64_t q = L->start[pid]; q < L->start[pid + 1]; q++) {
      const int32_t* cell = L->tris + 3 * (int64_t)L->cells[q];
      for (int k = 0; k < 3; k++) if (labels[cell[k]] != label) vec_push(&nb, cell[k]);
    }
  }
  int n = unique_i32(nb.p, (int)nb.n);
  int32_t best = -1;
  int64_t best_count = 0;
  for (int i = 0; i < n; i++) {
    int32_t li = labels[nb.p[i]];
    int64_t cnt = 0;
    for (int j = 0; j < n; j++) cnt += labels[nb.p[j]] == li;
    if (cnt > best_count) { best_count = cnt; best = li; }
  }
  free(nb.p);
  return n ? best : -1;
}

/* post_process.py:285-295 */
ORC_API void orc_pp_remove_islands(const int32_t* tris, int64_t T, int64_t V, int32_t* labels, int32_t label,
                                   int64_t min_count, int ignore_neg1) {
  links_t L;
  links_build(&L, tris, T, V);
  uint8_t* visited = (uint8_t*)calloc((size_t)(V > 0 ? V : 1), 1);
  vec_t region = {0, 0, 0};
  for (int64_t pid = 0; pid < V; pid++)
    if (labels[pid] == label && !visited[pid]) {
      connected_region(&L, (int32_t)pid, labels, label, visited, &region);
      if (region.n < min_count) {
        int32_t nl = neighbor_label(&L, labels, label, &region);
        if (ignore_neg1 && nl != -1)
          for (int64_t i = 0; i < region.n; i++) labels[region.p[i]] = nl;
      }
    }
  free(region.p); free(visited); links_free(&L);
}

/* post_process.py:297-304; returns the number of regions labelled */
ORC_API int64_t orc_pp_connectivity(const int32_t* tris, int64_t T, int64_t V, int32_t* labels, int32_t label,
                                    int32_t start_label) {
  links_t L;
  links_build(&L, tris, T, V);
  uint8_t* visited = (uint8_t*)calloc((size_t)(V > 0 ? V : 1), 1);
  vec_t region = {0, 0, 0};
  int64_t n = 0;
  for (int64_t pid = 0; pid < V; pid++)
    if (labels[pid] == label && !visited[pid]) {
      connected_region(&L, (int32_t)pid, labels, label, visited, &region);
      for (int64_t i = 0; i < region.n; i++) labels[region.p[i]] = start_label;
      start_label++;
      n++;
    }
  free(region.p); free(visited); links_free(&L);
  return n;
}

/* utils.py:218-231 -- sorted unique neighbours of pid */
static int neighbors(const links_t* L, int32_t pid, vec_t* out) {
  out->n = 0;
  for (int64_t q = L->start[pid]; q < L->start[pid + 1]; q++) {
    const int32_t* cell = L->tris + 3 * (int64_t)L->cells[q];
    for (int k = 0; k < 3; k++) if (cell[k] != pid) vec_push(out, cell[k]);
  }
  out->n = unique_i32(out->p, (int)out->n);
  return (int)out->n;
}

/* post_process.py:306-346; iterations < 0 = math.inf; returns the iterations that changed something */
ORC_API int64_t orc_pp_erode(const int32_t* tris, int64_t T, int64_t V, int32_t* labels, int32_t label, int has_ignore,
                             int32_t ignore_label, int64_t iterations, int has_target, int32_t target) {
  links_t L;
  links_build(&L, tris, T, V);
  vec_t todo = {0, 0, 0}, remain = {0, 0, 0}, nb = {0, 0, 0}, ch_pid = {0, 0, 0}, ch_lab = {0, 0, 0};
  for (int64_t pid = 0; pid < V; pid++) if (labels[pid] == label) vec_push(&todo, (int32_t)pid);
  int64_t done = 0;
  while (todo.n && iterations != 0) {
    vec_t t = remain; remain = todo; todo = t; todo.n = 0;
    ch_pid.n = ch_lab.n = 0;
    while (remain.n) {
      int32_t pid = remain.p[--remain.n];
      neighbors(&L, pid, &nb);
      int found = 0;
      for (int64_t i = 0; i < nb.n; i++) {
        int32_t nl = labels[nb.p[i]];
        if (nl != label && (!has_ignore || nl != ignore_label) && (!has_target || nl == target)) {
          vec_push(&ch_pid, pid); vec_push(&ch_lab, nl); found = 1;
          break;
        }
      }
      if (!found) vec_push(&todo, pid);
    }
    if (!ch_pid.n) break;
    for (int64_t i = 0; i < ch_pid.n; i++) labels[ch_pid.p[i]] = ch_lab.p[i];
    if (iterations > 0) iterations--;
    done++;
  }
  free(todo.p); free(remain.p); free(nb.p); free(ch_pid.p); free(ch_lab.p); links_free(&L);
  return done;
}

/* post_process.py:348-375 */
ORC_API void orc_pp_dilate(const int32_t* tris, int64_t T, int64_t V, int32_t* labels, int32_t label, int64_t iterations,
                           int over_target, int32_t target) {
  links_t L;
  links_build(&L, tris, T, V);
  vec_t nb = {0, 0, 0}, grow = {0, 0, 0};
  while (iterations > 0) {
    grow.n = 0;
    for (int64_t pid = 0; pid < V; pid++)
      if (labels[pid] == label) {
        neighbors(&L, (int32_t)pid, &nb);
        for (int64_t i = 0; i < nb.n; i++) {
          int32_t nl = labels[nb.p[i]];
          if (over_target ? nl == target : nl != label) vec_push(&grow, nb.p[i]);
        }
      }
    for (int64_t i = 0; i < grow.n; i++) labels[grow.p[i]] = label;
    iterations--;
  }
  free(nb.p); free(grow.p); links_free(&L);
}

/* post_process.py:377-380 */
ORC_API void orc_pp_relabel(int32_t* labels, int64_t V, int32_t label, int32_t relabel) {
  for (int64_t i = 0; i < V; i++) if (labels[i] == label) labels[i] = relabel;
}
