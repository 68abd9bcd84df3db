// Stand-alone probe of the tcgen05 path used by the tensor-core kernels: one CTA computes
// D[128 x 32] = A[128 x K] * B[32 x K]^T with kind::tf32 from un-swizzled K-major smem operands,
// once with tf32-exact inputs (must be bit-exact) and once as 3xTF32 on random fp32 (error ~1e-6).
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cfdbench_b200/csrc tools/tc_probe.cu -o tools/tc_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "fno_common.cuh"
#include "tc_common.cuh"
using namespace fno;

constexpr int M = 128, N = 32, K = 80;

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B,
                                             float* __restrict__ D, int split3) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a_hi = reinterpret_cast<float*>(smem);
  float* a_lo = a_hi + M * K;
  float* b_hi = a_lo + M * K;
  float* b_lo = b_hi + N * K;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<32>(&tmem_base_s);
  for (int e = tid; e < M * K; e += 128) {
    int m = e / K, k = e % K; float hi, lo; tc::split_tf32(A[e], hi, lo);
    *reinterpret_cast<float*>(reinterpret_cast<char*>(a_hi) + tc::kmajor_offset(m, k, M)) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(a_lo) + tc::kmajor_offset(m, k, M)) = lo;
  }
  for (int e = tid; e < N * K; e += 128) {
    int n = e / K, k = e % K; float hi, lo; tc::split_tf32(B[e], hi, lo);
    *reinterpret_cast<float*>(reinterpret_cast<char*>(b_hi) + tc::kmajor_offset(n, k, N)) = hi;
    *reinterpret_cast<float*>(reinterpret_cast<char*>(b_lo) + tc::kmajor_offset(n, k, N)) = lo;
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(M, N);
    constexpr uint32_t lboA = (M / 8) * 128, lboB = (N / 8) * 128;
    bool acc = false;
    for (int pass = 0; pass < (split3 ? 3 : 1); ++pass) {
      const float* pa = (pass == 1) ? a_lo : a_hi;
      const float* pb = (pass == 2) ? b_lo : b_hi;
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = tc::make_smem_desc(tc::smem_addr(pa) + ks * 2 * lboA, lboA, 128);
        const uint64_t db = tc::make_smem_desc(tc::smem_addr(pb) + ks * 2 * lboB, lboB, 128);
        tc::mma_tf32(tbase, da, db, idesc, acc);
        acc = true;
      }
    }
    tc::mma_commit(&bar);
  }
  // bounded wait so a wrong encoding cannot hang the box
  uint32_t spins = 0;
  while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { if (tid == 0) printf("TIMEOUT waiting for MMA\n"); __trap(); } }
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tbase + (static_cast<uint32_t>(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) D[(warp * 32 + (tid & 31)) * N + n] = v[n];
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<32>(tbase);
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  const size_t smem = (2 * M * K + 2 * N * K) * sizeof(float);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int fails = 0;
  for (int mode = 0; mode < 2; ++mode) {
    srand(1 + mode);
    for (auto& x : A) x = mode ? (float)rand() / RAND_MAX * 2 - 1 : (float)((rand() % 33) - 16) / 8.f;
    for (auto& x : B) x = mode ? (float)rand() / RAND_MAX * 2 - 1 : (float)((rand() % 33) - 16) / 8.f;
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
      maxerr = fmax(maxerr, fabs(r - D[m * N + n])); maxref = fmax(maxref, fabs(r));
    }
    printf("mode %d (%s): max abs err %.3e (max |ref| %.3f)\n", mode, mode ? "3xTF32 random" : "tf32-exact", maxerr, maxref);
    if (mode == 0 ? maxerr != 0.0 : maxerr > 6e-6) { fails++; printf("  first row: %f %f %f %f\n", D[0], D[1], D[2], D[3]); }
  }
  printf(fails ? "PROBE FAILED\n" : "PROBE OK\n");
  return fails;
}
