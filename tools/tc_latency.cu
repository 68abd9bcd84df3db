// Probe: throughput of dependent vs independent tcgen05.mma accumulation chains (kind::tf32, K = 8 per MMA),
// issued back to back from one elected thread with immediate descriptors (no issue-side overhead).
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cfdbench_b200/csrc tools/tc_latency.cu -o tools/tc_latency
#include <cstdio>
#include "fno_common.cuh"
#include "tc_common.cuh"
using namespace fno;

template <int M, int N, int NACC, int NMMA, int SKEWA = 0, int SKEWB = 0>
__global__ void __launch_bounds__(128) lat(long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a = reinterpret_cast<float*>(smem);   // [M x 8]
  float* b = a + 128 * 8;                      // [N x 8]
  __shared__ uint32_t tb;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int warp = tc::warp_index_uniform();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&tb);
  for (int e = tid; e < 128 * 8 + 256 * 8; e += 128) a[e] = 0.f;
  tc::fence_proxy_async_smem(); tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  if (warp == 0) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_tf32(M, N);
      const uint64_t da = tc::make_smem_desc(tc::smem_addr(a), (M / 8) * 128 + SKEWA, 128);
      const uint64_t db = tc::make_smem_desc(tc::smem_addr(b), (N / 8) * 128 + SKEWB, 128);
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          if (i < NACC) tc::mma_tf32_imm<false>(tb + (i % NACC) * N, da, db, idesc);
          else tc::mma_tf32_imm<true>(tb + (i % NACC) * N, da, db, idesc);
        }
        tc::mma_commit(&bar);
        const long long t1 = clock64();
        while (!mbar_try_wait(&bar, rep & 1)) {}
        const long long t2 = clock64();
        out[rep * 2] = t1 - t0;
        out[rep * 2 + 1] = t2 - t0;
      }
    }
    __syncwarp();
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tb);
}

template <int M, int N, int NACC, int NMMA, int SKEWA = 0, int SKEWB = 0>
void run(long long* d) {
  long long h[6];
  lat<M, N, NACC, NMMA, SKEWA, SKEWB><<<1, 128, (128 * 8 + 256 * 8) * 4 + 256>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); return; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("skewA %2d skewB %2d  M=%3d N=%3d accumulators=%d, %d MMAs: issue %5lld cyc, complete %5lld cyc = %5.1f cyc/MMA (%.0f MAC/clk)\n", SKEWA, SKEWB, M, N, NACC,
         NMMA, h[4], h[5], (double)h[5] / NMMA, (double)M * N * 8 * NMMA / h[5]);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<128, 32, 1, 32>(d);
  run<128, 32, 1, 32, 16, 0>(d);
  run<128, 32, 1, 32, 0, 16>(d);
  run<64, 48, 1, 32, 0, 16>(d);
  run<64, 48, 1, 32, 16, 16>(d);
  run<128, 64, 1, 32, 16, 0>(d);
  run<128, 32, 2, 32>(d);
  run<128, 32, 4, 32>(d);
  run<128, 64, 1, 32>(d);
  run<128, 64, 2, 32>(d);
  run<128, 128, 1, 32>(d);
  run<128, 128, 2, 32>(d);
  run<128, 256, 1, 32>(d);
  run<64, 48, 1, 32>(d);
  run<64, 48, 2, 32>(d);
  run<64, 48, 4, 32>(d);
  run<64, 96, 1, 32>(d);
  run<64, 192, 1, 32>(d);
  return 0;
}
