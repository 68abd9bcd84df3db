// Cost of mbarrier waits on B200: (a) wait on an already completed phase, (b) wake-up latency after another warp's arrive,
// each for mbarrier.try_wait (may suspend) and mbarrier.test_wait (pure poll).  Build: nvcc -arch=sm_100a -o mbar_probe mbar_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ bool test_wait(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
template <int MODE> __device__ __forceinline__ void wait(uint64_t* b, uint32_t par) {
  if (MODE == 0) { while (!try_wait(b, par)) {} } else { while (!test_wait(b, par)) {} }
}
// out[0..]: cycles
template <int MODE>
__global__ void probe(long long* out, int spinners) {
  __shared__ uint64_t bar[4];
  __shared__ volatile long long t_arrive[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i]))); }
  __syncthreads();
  // (a) completed-phase wait: thread 0 arrives itself (count 1) then waits 16 times on alternating phases
  if (threadIdx.x == 0) {
    long long acc = 0;
    for (int i = 0; i < 16; ++i) {
      arrive(&bar[0]);
      for (int k = 0; k < 50; ++k) asm volatile("" ::: "memory");
      long long t0 = clock64();
      wait<MODE>(&bar[0], i & 1);
      acc += clock64() - t0;
    }
    out[0] = acc / 16;
  }
  __syncthreads();
  // (b) ping-pong: warp 0 lane 0 waits on bar[1], warp 1 lane 0 arrives after a delay and stamps; other warps (spinners) spin on bar[3]
  if (warp == 0 && lane == 0) {
    long long acc = 0;
    for (int i = 0; i < 32; ++i) {
      wait<MODE>(&bar[1], i & 1);
      long long t1 = clock64();
      acc += t1 - t_arrive[i];
      arrive(&bar[2]);
    }
    out[1] = acc / 32;
    arrive(&bar[3]);
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < 32; ++i) {
      if (i > 0) wait<MODE>(&bar[2], (i - 1) & 1);
      long long t = clock64();
      while (clock64() - t < 2000) {}
      t_arrive[i] = clock64();
      arrive(&bar[1]);
    }
  } else if (warp >= 2 && warp < 2 + spinners) {
    wait<MODE>(&bar[3], 0);   // all lanes spin until the end
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  long long h[2];
  for (int sp : {0, 6, 22}) {
    probe<0><<<1, 32 * 24>>>(d, sp); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("try_wait : completed-phase wait %lld cycles, wake-up after arrive %lld cycles (%d spinning warps)\n", h[0], h[1], sp);
    probe<1><<<1, 32 * 24>>>(d, sp); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("test_wait: completed-phase wait %lld cycles, wake-up after arrive %lld cycles (%d spinning warps)\n", h[0], h[1], sp);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
