// Probe: latency of dependent vs independent tcgen05.mma (kind::tf32, M=128) accumulation chains.
#include <cstdio>
#include "fno_common.cuh"
#include "tc_common.cuh"
using namespace fno;

template <int N>
__global__ void __launch_bounds__(128) lat(long long* out, int n_mma, int n_acc) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a = reinterpret_cast<float*>(smem);          // [128 x 8]
  float* b = a + 128 * 8;                              // [N x 8]
  __shared__ uint32_t tb;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (tid < 32) tc::tmem_alloc<512>(&tb);
  for (int e = tid; e < 128 * 8 + N * 8; e += 128) a[e] = 0.f;
  tc::fence_proxy_async_smem(); tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    const uint64_t da = tc::make_smem_desc(tc::smem_addr(a), 2048, 128);
    const uint64_t db = tc::make_smem_desc(tc::smem_addr(b), (N / 8) * 128, 128);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) tc::mma_tf32(tb + (i % n_acc) * N, da, db, idesc, i >= n_acc);
      tc::mma_commit(&bar);
      long long t1 = clock64();
      while (!mbar_try_wait(&bar, rep & 1)) {}
      long long t2 = clock64();
      out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
    }
  }
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc<512>(tb);
}

template <int N>
void run(int n_mma, int n_acc) {
  long long* d; cudaMalloc(&d, 64); long long h[6];
  size_t smem = (128 * 8 + N * 8) * 4;
  lat<N><<<1, 128, smem>>>(d, n_mma, n_acc);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("N=%3d n_mma=%3d n_acc=%d: issue %lld cyc, complete %lld cyc (%.1f cyc/MMA)  [%s]\n", N, n_mma, n_acc, h[4], h[5],
         (double)h[5] / n_mma, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int acc : {1, 2, 4, 8}) run<32>(32, acc);
  run<32>(64, 1); run<32>(64, 16);
  for (int acc : {1, 2, 4}) run<128>(12, acc);
  run<128>(48, 1); run<128>(48, 4);
  run<64>(32, 1); run<64>(32, 4); run<256>(16, 1); run<256>(16, 2);
  return 0;
}
