// Probe 3: tcgen05.mma with the A operand in TENSOR MEMORY (filled by tcgen05.st), B from un-swizzled K-major smem.
// Checks exactness on tf32-exact data and times a 32-MMA chain against the both-operands-in-smem form
// (tools/tc_latency.cu: 45.6 cycles per M=128,N=32,K=8 MMA = shared-memory operand fetch at ~128 B/clk).
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cfdbench_b200/csrc tools/tc_probe3.cu -o tools/tc_probe3
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "fno_common.cuh"
#include "tc_common.cuh"
using namespace fno;

constexpr int M = 128, N = 32, K = 64;

template <bool kAcc>
__device__ __forceinline__ void mma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  if constexpr (kAcc)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "r"(a_tmem), "l"(b_desc), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                             long long* __restrict__ cyc) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* b_s = reinterpret_cast<float*>(smem);
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<128>(&tmem_base_s);
  for (int e = tid; e < N * K; e += 128) b_s[tc::kmajor_offset(e / K, e % K, N) / 4] = B[e];
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t a_tmem = tbase + 32;  // columns 32..95: A[m][k], row m in lane m
  // each thread writes its own row (lane quadrant of its warp)
  for (int k0 = 0; k0 < K; k0 += 16) {
    float v[16];
    for (int j = 0; j < 16; ++j) v[j] = A[(warp * 32 + lane) * K + k0 + j];
    tmem_st16(a_tmem + k0 + (static_cast<uint32_t>(warp * 32) << 16), v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  if (warp == 0) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_tf32(M, N);
      constexpr uint32_t lboB = (N / 8) * 128;
      const uint64_t db0 = tc::make_smem_desc(tc::smem_addr(b_s), lboB, 128);
      const long long t0 = clock64();
#pragma unroll
      for (int rep = 0; rep < 4; ++rep) {  // 32 MMAs: the K = 64 product, four times (last one kept)
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t db = db0 + ((ks * 2 * lboB) >> 4);
          if (ks == 0) mma_tf32_ta<false>(tbase, a_tmem + ks * 8, db, idesc);
          else mma_tf32_ta<true>(tbase, a_tmem + ks * 8, db, idesc);
        }
      }
      tc::mma_commit(&bar);
      const long long t1 = clock64();
      uint32_t spins = 0;
      while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { printf("TIMEOUT\n"); __trap(); } }
      const long long t2 = clock64();
      cyc[0] = t1 - t0;
      cyc[1] = t2 - t0;
    }
    __syncwarp();
  }
  uint32_t spins = 0;
  while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 24)) __trap(); }
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tbase + (static_cast<uint32_t>(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * N + n] = v[n];
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tbase);
}

int main() {
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(7);
  for (auto& x : A) x = (float)((rand() % 33) - 16) / 8.f;
  for (auto& x : B) x = (float)((rand() % 33) - 16) / 8.f;
  float *dA, *dB, *dD; long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, D.size() * 4);
  probe<<<1, 128, N * K * 4>>>(dA, dB, dD, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  long long c[2];
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
    maxerr = fmax(maxerr, fabs(r - D[m * N + n]));
  }
  printf("A in TMEM (tcgen05.st), B in smem: max abs err %.3e %s; 32 MMAs (M=128,N=32,K=8): issue %lld cyc, complete %lld cyc = %.1f cyc/MMA\n",
         maxerr, maxerr == 0 ? "OK" : "FAILED", c[0], c[1], c[1] / 32.0);
  if (maxerr != 0) printf("  D[0][0..3] = %f %f %f %f\n", D[0], D[1], D[2], D[3]);
  return maxerr != 0;
}
