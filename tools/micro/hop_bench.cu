// Measures the producer -> consumer hand-off latency between two CTAs on different SMs:
// CTA 0 publishes value i, CTA 1 waits for it and answers; one round = 2 hops.
// Variants: st.relaxed.gpu + ld.relaxed.gpu poll | atomicExch publish | atomic read poll | volatile.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <int MODE>
__global__ void pingpong(unsigned long long* a, unsigned long long* b, int rounds, long long* clocks, int npoll) {
  // extra polling threads (npoll per CTA) spin on unrelated addresses to load the memory system
  if (threadIdx.x > 0) {
    if ((int)threadIdx.x <= npoll) {
      const unsigned long long* p = a + 1024 + 64 * (blockIdx.x * blockDim.x + threadIdx.x);
      while (ld_relaxed(a + 512) == 0) {
        if (ld_relaxed(p) == 12345) break;
      }
    }
    return;
  }
  if (blockIdx.x > 1) {  // bystander CTAs: their thread 0 just waits for the end flag
    while (ld_relaxed(a + 512) == 0) {
    }
    return;
  }
  const long long t0 = clock64();
  for (int i = 1; i <= rounds; ++i) {
    if (blockIdx.x == 0) {
      if (MODE == 1) atomicExch(a, (unsigned long long)i); else st_relaxed(a, i);
      if (MODE == 2) { while (atomicAdd(b, 0ULL) < (unsigned long long)i) {} }
      else { while (ld_relaxed(b) < (unsigned long long)i) {} }
    } else {
      if (MODE == 2) { while (atomicAdd(a, 0ULL) < (unsigned long long)i) {} }
      else { while (ld_relaxed(a) < (unsigned long long)i) {} }
      if (MODE == 1) atomicExch(b, (unsigned long long)i); else st_relaxed(b, i);
    }
  }
  clocks[blockIdx.x] = clock64() - t0;
  if (blockIdx.x == 0) st_relaxed(a + 512, 1);
}
int main() {
  unsigned long long* buf;
  long long* clk;
  cudaMalloc(&buf, 64 << 20);
  cudaMalloc(&clk, 64);
  const int rounds = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int ctas : {2, 148})
      for (int npoll : {0, 127, 511}) {
        cudaMemset(buf, 0, 64 << 20);
        const int threads = npoll + 1;
        if (mode == 0) pingpong<0><<<ctas, threads>>>(buf, buf + 128, rounds, clk, npoll);
        if (mode == 1) pingpong<1><<<ctas, threads>>>(buf, buf + 128, rounds, clk, npoll);
        if (mode == 2) pingpong<2><<<ctas, threads>>>(buf, buf + 128, rounds, clk, npoll);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
        printf("mode %d (0 st/ld relaxed, 1 atomicExch publish, 2 atomic-read poll) ctas %3d pollers/cta %3d: %6.0f clocks per hop (%s)\n", mode, ctas,
               npoll, (double)h[0] / rounds / 2.0, cudaGetErrorString(e));
      }
  return 0;
}
