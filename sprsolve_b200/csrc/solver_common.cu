// solver_common.cu -- lagged status polling shared by the device-resident solvers.
#include "solver.cuh"

namespace spb {

Poller::Poller(Ctx* ctx) : c(ctx) {
  StateHead* base = nullptr;
  SPB_CUDA(cudaMallocHost(&base, sizeof(StateHead) * 2));
  pinned[0] = base;
  pinned[1] = base + 1;
  SPB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  SPB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
}

Poller::~Poller() {
  for (int s = 0; s < 2; ++s) {
    if (pending[s]) cudaEventSynchronize(ev[s]);
    cudaEventDestroy(ev[s]);
  }
  cudaFreeHost(pinned[0]);
}

void Poller::post(const void* d_state) {
  const int s = cur;
  if (pending[s]) SPB_CUDA(cudaEventSynchronize(ev[s]));
  SPB_CUDA(cudaMemcpyAsync(pinned[s], d_state, sizeof(StateHead), cudaMemcpyDeviceToHost, c->stream));
  SPB_CUDA(cudaEventRecord(ev[s], c->stream));
  pending[s] = true;
  cur ^= 1;
}

bool Poller::wait_oldest(StateHead* out) {
  if (!(pending[0] && pending[1])) return false;
  const int s = cur;  // the slot that would be overwritten next is the older one
  SPB_CUDA(cudaEventSynchronize(ev[s]));
  *out = *pinned[s];
  pending[s] = false;
  return true;
}

bool Poller::drain(StateHead* out) {
  bool any = false;
  const int order[2] = {cur, cur ^ 1};
  for (int k = 0; k < 2; ++k) {
    const int s = order[k];
    if (!pending[s]) continue;
    SPB_CUDA(cudaEventSynchronize(ev[s]));
    *out = *pinned[s];
    pending[s] = false;
    any = true;
  }
  return any;
}

}  // namespace spb
