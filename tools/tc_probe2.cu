// Probe 2: (A) kind::f16 (bf16 x bf16 -> f32) UMMA from un-swizzled K-major operands, the A operand staged with
// 16-byte cp.async chunks straight from a row-major bf16 matrix and a skewed LBO; (B) where the 64 rows of an M=64
// accumulator land in TMEM.
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cfdbench_b200/csrc tools/tc_probe2.cu -o tools/tc_probe2
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "fno_common.cuh"
#include "tc_common.cuh"
using namespace fno;

__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}

constexpr int MA = 128, NA = 32, KA = 64;
constexpr uint32_t kLboA = (MA / 8) * 128 + 16, kLboB = (NA / 8) * 128;

__global__ void __launch_bounds__(128) probe_bf16(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                  float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* a_s = smem;                       // 8 K-chunks x 2064 B
  unsigned char* b_s = smem + 8 * kLboA + 112;     // keep 128-byte alignment: 8*2064 = 16512 = 129*128
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<32>(&tmem_base_s);
  // A: row m = 64 bf16 = 8 chunks of 16 B; chunk (m, kc) -> kc*LBO + (m>>3)*128 + (m&7)*16
  for (int t = tid; t < MA * 8; t += 128) {
    const int m = t >> 3, kc = t & 7;
    const uint32_t dst = tc::smem_addr(a_s) + kc * kLboA + (m >> 3) * 128 + (m & 7) * 16;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(A + m * KA + kc * 8) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int e = tid; e < NA * KA; e += 128) {
    const int n = e / KA, k = e % KA;
    *reinterpret_cast<__nv_bfloat16*>(b_s + (k >> 3) * kLboB + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = B[e];
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t idesc = idesc_bf16(MA, NA);
    for (int ks = 0; ks < KA / 16; ++ks) {
      const uint64_t da = tc::make_smem_desc(tc::smem_addr(a_s) + ks * 2 * kLboA, kLboA, 128);
      const uint64_t db = tc::make_smem_desc(tc::smem_addr(b_s) + ks * 2 * kLboB, kLboB, 128);
      mma_f16(tbase, da, db, idesc, ks > 0);
    }
    tc::mma_commit(&bar);
  }
  uint32_t spins = 0;
  while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { if (tid == 0) printf("TIMEOUT (bf16)\n"); __trap(); } }
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tbase + (static_cast<uint32_t>(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) D[(warp * 32 + (tid & 31)) * NA + n] = v[n];
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<32>(tbase);
}

constexpr int MB = 64, NB = 32, KB = 16;
__global__ void __launch_bounds__(128) probe_m64(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* a_s = reinterpret_cast<float*>(smem);
  float* b_s = a_s + MB * KB;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<32>(&tmem_base_s);
  for (int e = tid; e < MB * KB; e += 128) a_s[tc::kmajor_offset(e / KB, e % KB, MB) / 4] = A[e];
  for (int e = tid; e < NB * KB; e += 128) b_s[tc::kmajor_offset(e / KB, e % KB, NB) / 4] = B[e];
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(MB, NB);
    constexpr uint32_t lboA = (MB / 8) * 128, lboB = (NB / 8) * 128;
    for (int ks = 0; ks < KB / 8; ++ks)
      tc::mma_tf32(tbase, tc::make_smem_desc(tc::smem_addr(a_s) + ks * 2 * lboA, lboA, 128),
                   tc::make_smem_desc(tc::smem_addr(b_s) + ks * 2 * lboB, lboB, 128), idesc, ks > 0);
    tc::mma_commit(&bar);
  }
  uint32_t spins = 0;
  while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { if (tid == 0) printf("TIMEOUT (m64)\n"); __trap(); } }
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tbase + (static_cast<uint32_t>(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) D[(warp * 32 + (tid & 31)) * NB + n] = v[n];   // D[tmem lane][column]
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<32>(tbase);
}

int main() {
  int fails = 0;
  {  // ---------------- A: bf16
    std::vector<__nv_bfloat16> A(MA * KA), B(NA * KA);
    std::vector<float> Af(MA * KA), Bf(NA * KA), D(MA * NA);
    srand(3);
    for (int i = 0; i < MA * KA; ++i) { Af[i] = (float)((rand() % 33) - 16) / 8.f; A[i] = __float2bfloat16(Af[i]); }
    for (int i = 0; i < NA * KA; ++i) { Bf[i] = (float)((rand() % 33) - 16) / 8.f; B[i] = __float2bfloat16(Bf[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const size_t smem = 8 * kLboA + 112 + 8 * kLboB + 256;
    cudaFuncSetAttribute(probe_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_bf16<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("bf16: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < MA; ++m) for (int n = 0; n < NA; ++n) {
      double r = 0; for (int k = 0; k < KA; ++k) r += (double)Af[m * KA + k] * Bf[n * KA + k];
      maxerr = fmax(maxerr, fabs(r - D[m * NA + n]));
    }
    printf("A: kind::f16 bf16, cp.async-staged skewed-LBO A operand: max abs err %.3e %s\n", maxerr, maxerr == 0 ? "OK" : "FAILED");
    if (maxerr != 0) { fails++; printf("   D[0][0..3] = %f %f %f %f\n", D[0], D[1], D[2], D[3]); }
  }
  {  // ---------------- B: M = 64 accumulator layout
    std::vector<float> A(MB * KB), B(NB * KB), D(128 * NB), R(MB * NB);
    srand(4);
    for (auto& x : A) x = (float)((rand() % 33) - 16) / 8.f;
    for (auto& x : B) x = (float)((rand() % 33) - 16) / 8.f;
    for (int m = 0; m < MB; ++m) for (int n = 0; n < NB; ++n) { double r = 0; for (int k = 0; k < KB; ++k) r += (double)A[m * KB + k] * B[n * KB + k]; R[m * NB + n] = (float)r; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (MB * KB + NB * KB) * 4;
    probe_m64<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("m64: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    printf("B: M=64 accumulator: row -> TMEM lane:");
    int found = 0;
    for (int m = 0; m < MB; ++m) {
      int lane = -1;
      for (int l = 0; l < 128 && lane < 0; ++l) { bool eq = true; for (int n = 0; n < NB; ++n) eq = eq && D[l * NB + n] == R[m * NB + n]; if (eq) lane = l; }
      if (m % 16 == 0) printf("\n   ");
      printf("%d->%d ", m, lane);
      found += lane >= 0;
    }
    printf("\n   %d of 64 rows located\n", found);
    if (found != 64) fails++;
  }
  printf(fails ? "PROBE2 FAILED\n" : "PROBE2 OK\n");
  return fails;
}
