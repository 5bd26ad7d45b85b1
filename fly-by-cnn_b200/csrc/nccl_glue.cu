// nccl_glue.cu -- the one collective of the path: the sum of label votes over
// the ranks that rendered different views / meshes.  NCCL is
// bound at run time from the libnccl.so.2 that PyTorch ships, so the library
// has no link-time dependency on it; the all-reduce is enqueued on the context
// stream directly behind the scatter kernel (no host synchronisation between).
#include <dlfcn.h>

#include "ctx.h"

namespace fb {

// minimal NCCL surface (matches nccl.h 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2 };
enum { ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load(flyby_ctx* c) {
  if (g_nccl.handle) return FLYBY_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    const char* env = getenv("FLYBY_NCCL_LIB");
    if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!h) return fail(c, FLYBY_ERR_NCCL, "libnccl.so.2 not found (set FLYBY_NCCL_LIB): %s", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.Reduce = (decltype(g_nccl.Reduce))dlsym(h, "ncclReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.Reduce)
    return fail(c, FLYBY_ERR_NCCL, "libnccl is missing required symbols");
  g_nccl.handle = h;
  return FLYBY_OK;
}

#define FB_NCCL(ctx, expr)                                                                   \
  do {                                                                                       \
    ncclResult_t _r = (expr);                                                                \
    if (_r != 0)                                                                             \
      return fail((ctx), FLYBY_ERR_NCCL, "%s: %s", #expr,                                    \
                  g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");         \
  } while (0)

int nccl_unique_id(flyby_ctx* c, uint8_t* out) {
  int rc = nccl_load(c);
  if (rc) return rc;
  ncclUniqueId id;
  FB_NCCL(c, g_nccl.GetUniqueId(&id));
  memcpy(out, id.internal, 128);
  return FLYBY_OK;
}

int nccl_init(flyby_ctx* c, const uint8_t* idbytes, int rank, int world) {
  int rc = nccl_load(c);
  if (rc) return rc;
  if (c->nccl_comm) {
    g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
  ncclUniqueId id;
  memcpy(id.internal, idbytes, 128);
  ncclComm_t comm;
  FB_CUDA(c, cudaSetDevice(c->device));
  FB_NCCL(c, g_nccl.CommInitRank(&comm, world, id, rank));
  c->nccl_comm = comm;
  c->nccl_rank = rank;
  c->nccl_world = world;
  return FLYBY_OK;
}

void nccl_destroy(flyby_ctx* c) {
  if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
}

int nccl_votes_allreduce(flyby_ctx* c, int32_t* votes, int64_t n, int root) {
  if (!c->nccl_comm) return fail(c, FLYBY_ERR_NCCL, "flyby_nccl_init has not been called");
  if (root < 0)
    FB_NCCL(c, g_nccl.AllReduce(votes, votes, (size_t)n, ncclInt32, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
  else
    FB_NCCL(c, g_nccl.Reduce(votes, votes, (size_t)n, ncclInt32, ncclSum, root, (ncclComm_t)c->nccl_comm, c->stream));
  return FLYBY_OK;
}

}  // namespace fb
