// dist.cuh -- one-process-per-GPU plumbing: NCCL (resolved at run time with dlopen, so the library
// loads on machines without NCCL or a GPU), row-block partition bookkeeping.
#pragma once
#include <nccl.h>

#include <vector>

#include "common.cuh"

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

struct Dist {
  int world = 1, rank = 0;
  ncclComm_t comm = nullptr;       // scalar all-reduces, on the compute stream
  ncclComm_t comm_halo = nullptr;  // halo send/recv, on the comm stream
  DevBuf scratch;                  // small device scratch for all-gathers
};

// In-place sum of `count` doubles across ranks on the compute stream (no-op when single GPU).
void allreduce_sum(Ctx* ctx, double* dev, size_t count);
void allgather_i64(Ctx* ctx, const int64_t* host_in, size_t count, std::vector<int64_t>& host_out);

void stencil_partition(int kind, int64_t nx, int64_t ny, int64_t nz, int world, int rank,
                       int64_t* row_begin, int64_t* row_end);

}  // namespace spb
