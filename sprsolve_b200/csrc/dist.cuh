// dist.cuh -- one-process-per-GPU plumbing: NCCL (resolved at run time with dlopen, so the library
// loads on machines without NCCL or a GPU), row-block partition bookkeeping.
#pragma once
#include <nccl.h>

#include <vector>

#include "common.cuh"
#include "peer.cuh"

namespace spb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};
const NcclApi& nccl();  // throws SPB_NCCL_ERROR if libnccl.so.2 cannot be loaded

#define SPB_NCCL(call)                                                                       \
  do {                                                                                       \
    ncclResult_t _r = (call);                                                                \
    if (_r != ncclSuccess) {                                                                 \
      char _b[512];                                                                          \
      snprintf(_b, sizeof(_b), "NCCL error at %s:%d: %s", __FILE__, __LINE__,                \
               ::spb::nccl().GetErrorString(_r));                                            \
      ::spb::set_last_error(_b);                                                             \
      throw ::spb::SpbError{SPB_NCCL_ERROR};                                                 \
    }                                                                                        \
  } while (0)

// A block of device memory every rank of the node has mapped through CUDA IPC (NVLink peer
// access): the transport of the scalar all-reduce and of the halo "put" (peer.cuh).
struct PeerWindow {
  void* local = nullptr;
  size_t bytes = 0;
  int world = 1, rank = 0;
  std::vector<void*> mapped;  // [world], mapped[rank] == local
  PeerPtrs ptrs() const {
    PeerPtrs pp;
    for (int q = 0; q < kMaxPeers; ++q) pp.p[q] = q < world ? mapped[q] : nullptr;
    pp.world = world;
    pp.rank = rank;
    return pp;
  }
};
// Collective over all ranks of ctx's communicator.  Returns nullptr (on every rank) when some
// rank could not map some peer (no NVLink/IPC between them): the caller falls back to NCCL.
PeerWindow* window_create(Ctx* ctx, size_t bytes);
void window_destroy(PeerWindow* w);

struct Dist {
  int world = 1, rank = 0;
  ncclComm_t comm = nullptr;       // scalar all-reduces (NCCL transport), set-up all-gathers
  ncclComm_t comm_halo = nullptr;  // halo send/recv (NCCL transport), on the comm stream
  DevBuf scratch;                  // small device scratch for all-gathers
  DevBuf dd_buf;                   // NCCL transport: unrounded (hi, lo) pairs of a reduction point, [1 + world][8]
  // peer-memory transport (default; SPB_COMM=nccl selects the NCCL transport instead)
  bool peer = false;
  PeerWindow* scal = nullptr;      // ScalWin of every rank
};

inline bool peer_mode(const Ctx* ctx) { return ctx->dist && ctx->dist->peer; }
// Raises SPB_NCCL_ERROR / SPB_CUDA_ERROR if a device-side spin timed out since the last check (scalar
// all-reduce window, halo put / halo flags, Gauss-Seidel cross-block hand-off) and clears the flag.
// Synchronises the stream.
void device_check(Ctx* ctx);

// In-place sum of `count` (<= 4) doubles across ranks on the compute stream (no-op when single
// GPU).  Peer transport: one 32-thread kernel; NCCL transport: ncclAllReduce.
void allreduce_sum(Ctx* ctx, double* dev, size_t count);
void allgather_i64(Ctx* ctx, const int64_t* host_in, size_t count, std::vector<int64_t>& host_out);

void stencil_partition(int kind, int64_t nx, int64_t ny, int64_t nz, int world, int rank,
                       int64_t* row_begin, int64_t* row_end);

}  // namespace spb
