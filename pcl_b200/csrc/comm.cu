// comm.cu — cross-GPU reduction of the per-iteration accumulators.
//
// The reference has no multi-GPU path.  Here the source cloud is sharded across ranks (one process per
// GPU, target index replicated); the only exchange per ICP iteration is a sum of kAccum (= 40) doubles,
// issued on the same stream as the iteration kernel so the solve kernel that follows sees the global sums.
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy torch already loaded, if any) so the library
// has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>

#include <cstdlib>
#include <vector>

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
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.h, "ncclAllGather"));
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

enum { ncclUint8 = 1 };

struct Comm {
  ncclComm_t comm = nullptr;  // nullptr: peer-memory exchange bootstrapped by the caller (comm_import), no NCCL at all
  int rank = 0, nranks = 1;
  int mode = PCLB200_REDUCE_FUSED;
  // fused peer-memory reduce (NVLink): one IPC-shared block per rank
  unsigned char* local_block = nullptr;
  void* peer_block[kMaxRanks] = {};
  PeerView view;
  bool peer_ok = false;
  unsigned long long seq = 0;
};

void comm_destroy(Ctx& c);

static size_t peer_block_bytes() { return kMaxRanks * sizeof(unsigned long long) + 2ull * kMaxRanks * kAccum * sizeof(double); }

// Maps every rank's exchange block into this process (CUDA IPC over NVLink / NVSwitch).  Any failure leaves
// peer_ok = false and the per-iteration reduce falls back to ncclAllReduce.
static void setup_peer_reduce(Ctx& c, Comm& cm)
{
  if (cm.nranks > kMaxRanks || !nccl().AllGather)
    return;
  cudaStream_t st = c.stream;
  if (cudaMalloc(reinterpret_cast<void**>(&cm.local_block), peer_block_bytes()) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  cudaMemsetAsync(cm.local_block, 0, peer_block_bytes(), st);
  cudaIpcMemHandle_t mine;
  bool ok = cudaIpcGetMemHandle(&mine, cm.local_block) == cudaSuccess;
  // exchange handles (+ an "ok" byte) through the communicator we already have
  struct Rec { cudaIpcMemHandle_t h; unsigned char ok; unsigned char pad[63]; };
  static_assert(sizeof(Rec) == 128, "record size");
  Rec r;
  memset(&r, 0, sizeof(r));
  r.h = mine;
  r.ok = ok ? 1 : 0;
  DevBuf<unsigned char> d_send, d_recv;
  d_send.alloc(sizeof(Rec), st);
  d_recv.alloc(sizeof(Rec) * cm.nranks, st);
  PCLB_CUDA(cudaMemcpyAsync(d_send.p, &r, sizeof(Rec), cudaMemcpyHostToDevice, st));
  PCLB_NCCL(nccl().AllGather(d_send.p, d_recv.p, sizeof(Rec), ncclUint8, cm.comm, st));
  std::vector<Rec> all(cm.nranks);
  PCLB_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, sizeof(Rec) * cm.nranks, cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  bool all_ok = true;
  for (int p = 0; p < cm.nranks; ++p)
    all_ok = all_ok && all[p].ok;
  if (all_ok) {
    for (int p = 0; p < cm.nranks && all_ok; ++p) {
      if (p == cm.rank) {
        cm.peer_block[p] = cm.local_block;
        continue;
      }
      if (cudaIpcOpenMemHandle(&cm.peer_block[p], all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        cm.peer_block[p] = nullptr;
        all_ok = false;
      }
    }
  }
  // every rank must take the same decision: agree through one more tiny all-reduce
  DevBuf<double> d_flag;
  d_flag.alloc(1, st);
  double hv = all_ok ? 0.0 : 1.0;
  PCLB_CUDA(cudaMemcpyAsync(d_flag.p, &hv, sizeof(double), cudaMemcpyHostToDevice, st));
  PCLB_NCCL(nccl().AllReduce(d_flag.p, d_flag.p, 1, ncclFloat64, ncclSum, cm.comm, st));
  PCLB_CUDA(cudaMemcpyAsync(&hv, d_flag.p, sizeof(double), cudaMemcpyDeviceToHost, st));
  PCLB_CUDA(cudaStreamSynchronize(st));
  if (hv != 0.0)
    return;
  cm.view.rank = cm.rank;
  cm.view.nranks = cm.nranks;
  for (int p = 0; p < cm.nranks; ++p) {
    unsigned char* b = static_cast<unsigned char*>(cm.peer_block[p]);
    cm.view.flags[p] = reinterpret_cast<unsigned long long*>(b);
    cm.view.slots[p] = reinterpret_cast<double*>(b + kMaxRanks * sizeof(unsigned long long));
  }
  cm.peer_ok = true;
}

static void teardown_peer_reduce(Comm& cm)
{
  for (int p = 0; p < cm.nranks && p < kMaxRanks; ++p)
    if (p != cm.rank && cm.peer_block[p])
      cudaIpcCloseMemHandle(cm.peer_block[p]);
  if (cm.local_block)
    cudaFree(cm.local_block);
  cm.local_block = nullptr;
  cm.peer_ok = false;
}

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
    teardown_peer_reduce(*c.comm);
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
  try {
    setup_peer_reduce(c, *cm);
  }
  catch (const Error&) {
    cm->peer_ok = false;  // keep the NCCL path
  }
}

void comm_destroy(Ctx& c)
{
  if (c.comm) {
    teardown_peer_reduce(*c.comm);
    if (c.comm->comm)
      nccl().CommDestroy(c.comm->comm);
    delete c.comm;
    c.comm = nullptr;
  }
}

// ---- caller-bootstrapped peer exchange (no NCCL): export the local block's IPC handle, import everybody's ----------------
// For deployments whose launcher already has a way to all-gather 64 bytes per rank (MPI, gloo, a file), and for
// ranks that share one GPU (NCCL refuses duplicate devices; CUDA IPC does not).
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");

void comm_export(Ctx& c, void* out64)
{
  comm_destroy(c);
  Comm* cm = new Comm();
  c.comm = cm;
  PCLB_CUDA(cudaSetDevice(c.device));
  PCLB_CUDA(cudaMalloc(reinterpret_cast<void**>(&cm->local_block), peer_block_bytes()));
  PCLB_CUDA(cudaMemsetAsync(cm->local_block, 0, peer_block_bytes(), c.stream));
  PCLB_CUDA(cudaStreamSynchronize(c.stream));
  cudaIpcMemHandle_t h;
  PCLB_CUDA(cudaIpcGetMemHandle(&h, cm->local_block));
  memcpy(out64, &h, sizeof(h));
}

void comm_import(Ctx& c, int rank, int nranks, const void* handles)
{
  PCLB_REQUIRE(c.comm && c.comm->local_block && !c.comm->comm, PCLB200_ERR_INVALID, "comm_import: call comm_export first");
  PCLB_REQUIRE(nranks >= 2 && nranks <= kMaxRanks && rank >= 0 && rank < nranks, PCLB200_ERR_INVALID, "bad rank / nranks");
  Comm& cm = *c.comm;
  cm.rank = rank;
  cm.nranks = nranks;
  const cudaIpcMemHandle_t* all = static_cast<const cudaIpcMemHandle_t*>(handles);
  for (int p = 0; p < nranks; ++p) {
    if (p == rank) {
      cm.peer_block[p] = cm.local_block;
      continue;
    }
    cudaError_t e = cudaIpcOpenMemHandle(&cm.peer_block[p], all[p], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      cm.peer_block[p] = nullptr;
      throw Error(PCLB200_ERR_CUDA, std::string("comm_import: cudaIpcOpenMemHandle failed for rank ") + std::to_string(p) +
                                        ": " + cudaGetErrorString(e));
    }
  }
  cm.view.rank = rank;
  cm.view.nranks = nranks;
  for (int p = 0; p < nranks; ++p) {
    unsigned char* b = static_cast<unsigned char*>(cm.peer_block[p]);
    cm.view.flags[p] = reinterpret_cast<unsigned long long*>(b);
    cm.view.slots[p] = reinterpret_cast<double*>(b + kMaxRanks * sizeof(unsigned long long));
  }
  cm.peer_ok = true;
}

void comm_set_mode(Ctx& c, int mode)
{
  PCLB_REQUIRE(mode == PCLB200_REDUCE_FUSED || mode == PCLB200_REDUCE_NCCL, PCLB200_ERR_INVALID, "unknown reduce mode");
  PCLB_REQUIRE(c.comm != nullptr, PCLB200_ERR_INVALID, "no communicator");
  PCLB_REQUIRE(mode == PCLB200_REDUCE_FUSED || c.comm->comm != nullptr, PCLB200_ERR_INVALID,
               "the NCCL reduce needs a communicator made by pclb200_comm_init");
  c.comm->mode = mode;
}

bool comm_active(const Ctx& c) { return c.comm && c.comm->nranks > 1; }

bool comm_peer_fused(const Ctx& c) { return comm_active(c) && c.comm->peer_ok && c.comm->mode == PCLB200_REDUCE_FUSED; }

// view + next sequence number for a fused in-kernel exchange; returns false when the NCCL path must be used
bool comm_peer_view(Ctx& c, PeerView* view, unsigned long long* seq)
{
  if (!comm_active(c) || !c.comm->peer_ok || c.comm->mode != PCLB200_REDUCE_FUSED)
    return false;
  *view = c.comm->view;
  *seq = ++c.comm->seq;
  return true;
}

void comm_allreduce_sum(Ctx& c, double* d_buf, int count)
{
  if (!comm_active(c))
    return;
  PCLB_REQUIRE(c.comm->comm != nullptr, PCLB200_ERR_NCCL, "no NCCL communicator for the library all-reduce");
  PCLB_NCCL(nccl().AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, c.comm->comm, c.stream));
  ++c.launches;
}

}  // namespace pclb200
