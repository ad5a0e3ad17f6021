// comm.cu — cross-GPU reduction of the per-iteration accumulators.
//
// The reference has no multi-GPU path.  Here the source cloud is sharded across ranks (one process per
// GPU, target index replicated); the only exchange per ICP iteration is a sum of kAccum (= 40) doubles,
// issued on the same stream as the iteration kernel so the solve kernel that follows sees the global sums.
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy torch already loaded, if any) so the library
// has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>

#include "internal.cuh"

namespace pclb200 {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

static NcclApi& nccl()
{
  static NcclApi api;
  if (!api.h) {
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h)
      api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    PCLB_REQUIRE(api.h != nullptr, PCLB200_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.h, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(api.h, "ncclAllReduce"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.h, "ncclGetErrorString"));
    PCLB_REQUIRE(api.GetUniqueId && api.CommInitRank && api.AllReduce, PCLB200_ERR_NCCL, "libnccl lacks expected symbols");
  }
  return api;
}

#define PCLB_NCCL(expr)                                                                                  \
  do {                                                                                                   \
    int _r = (expr);                                                                                     \
    if (_r != ncclSuccess)                                                                               \
      throw Error(PCLB200_ERR_NCCL, std::string(#expr) + ": " +                                          \
                                        (nccl().GetErrorString ? nccl().GetErrorString(_r) : "nccl error")); \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

void comm_unique_id(void* out128)
{
  ncclUniqueId id;
  PCLB_NCCL(nccl().GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
}

void comm_init(Ctx& c, int rank, int nranks, const void* unique_id)
{
  PCLB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, PCLB200_ERR_INVALID, "bad rank / nranks");
  if (c.comm) {
    if (c.comm->comm)
      nccl().CommDestroy(c.comm->comm);
    delete c.comm;
    c.comm = nullptr;
  }
  if (nranks == 1)
    return;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  Comm* cm = new Comm();
  cm->rank = rank;
  cm->nranks = nranks;
  PCLB_CUDA(cudaSetDevice(c.device));
  int r = nccl().CommInitRank(&cm->comm, nranks, id, rank);
  if (r != ncclSuccess) {
    delete cm;
    throw Error(PCLB200_ERR_NCCL, "ncclCommInitRank failed");
  }
  c.comm = cm;
}

void comm_destroy(Ctx& c)
{
  if (c.comm) {
    if (c.comm->comm)
      nccl().CommDestroy(c.comm->comm);
    delete c.comm;
    c.comm = nullptr;
  }
}

bool comm_active(const Ctx& c) { return c.comm && c.comm->nranks > 1; }

void comm_allreduce_sum(Ctx& c, double* d_buf, int count)
{
  if (!comm_active(c))
    return;
  PCLB_NCCL(nccl().AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, c.comm->comm, c.stream));
  ++c.launches;
}

}  // namespace pclb200
