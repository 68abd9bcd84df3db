// Probe 5: the operand forms the fused Fourier-block kernel (fno_block_fused.cu) relies on.
//   T1  kind::tf32, B operand MN-major SWIZZLE_128B (a [K][32] fp32 tile = one 128-byte row per k), A K-major.
//   T2  kind::tf32, A operand MN-major SWIZZLE_128B with 4 groups of 32 rows (LBO between groups), B K-major,
//       the second half of K issued with the a_negate bit of the instruction descriptor.
//   T3  kind::f16 (bf16 x bf16), A operand MN-major SWIZZLE_128B written by TMA (cuTensorMapEncodeTiled, two
//       {64 px, 32 ch} boxes), accumulated ON TOP of a kind::tf32 chain in the same TMEM accumulator;
//       T3b: B operand as fp16 with a bf16 A (mixed 16-bit types).
//   T4  A operand (tf32) in TENSOR MEMORY with the MN-major swizzled B of T1.
//   T5  timings of the MMA shapes the kernel issues.
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -I cfdbench_b200/csrc tools/tc_probe5.cu -o tools/tc_probe5
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "fno_common.cuh"
#include "tc_common.cuh"
using namespace fno;

__host__ __device__ constexpr uint32_t swz128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }
// 32-bit MN-major operands: "128B swizzle with a 32B base" (layout type 1, the only MN-major layout kind::tf32
// accepts): atom = 4 k-rows of 128 B, the 32-byte chunk index is XORed with the row index mod 4.
__host__ __device__ constexpr uint32_t swz128_32(uint32_t off) { return off ^ (((off >> 7) & 3u) << 5); }
__device__ __forceinline__ uint64_t desc_sw128_32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return tc::make_smem_desc(saddr, lbo, sbo) | (static_cast<uint64_t>(1) << 61);
}

// smem descriptor with SWIZZLE_128B (layout type 2 in bits [61,64))
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return tc::make_smem_desc(saddr, lbo, sbo) | (static_cast<uint64_t>(2) << 61);
}
constexpr uint32_t kAMajorMN = 1u << 15, kBMajorMN = 1u << 16, kANeg = 1u << 13;
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int afmt, int bfmt) {  // 0 = f16, 1 = bf16
  return (1u << 4) | (static_cast<uint32_t>(afmt) << 7) | (static_cast<uint32_t>(bfmt) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ta(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, bool acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

__device__ void wait_bar(uint64_t* bar, uint32_t parity, const char* what) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) { if (threadIdx.x == 0) printf("TIMEOUT %s\n", what); __trap(); }
  }
}

// ---------------------------------------------------------------------------------------------- T1 / T4
// D[128][32] = A[128][48] * B[48][32];  TMEM_A = 1 puts A in tensor memory.
template <int TMEM_A>
__global__ void __launch_bounds__(128) probe_b_mn(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* b_s = smem;                                  // 48 rows x 128 B (swizzled), 6 KB
  float* a_s = reinterpret_cast<float*>(smem + 6144);          // K-major, 128 x 48
  __shared__ uint32_t tb_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<128>(&tb_s);
  for (int e = tid; e < 48 * 32; e += 128) {
    const int k = e / 32, n = e % 32;
    *reinterpret_cast<float*>(b_s + swz128_32(k * 128 + n * 4)) = B[e];
  }
  for (int e = tid; e < 128 * 48; e += 128) a_s[tc::kmajor_offset(e / 48, e % 48, 128) / 4] = A[e];
  tc::fence_proxy_async_smem(); tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  const uint32_t tb = tb_s, a_tmem = tb + 32;
  if (TMEM_A) {
    for (int k0 = 0; k0 < 48; k0 += 16) {
      float v[16];
      for (int j = 0; j < 16; ++j) v[j] = A[(warp * 32 + lane) * 48 + k0 + j];
      tc::tmem_st16(a_tmem + k0 + (static_cast<uint32_t>(warp * 32) << 16), v);
    }
    tc::tmem_wait_st();
    tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  }
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, 32) | kBMajorMN;
    for (int ks = 0; ks < 6; ++ks) {
      const uint64_t db = desc_sw128_32(tc::smem_addr(b_s) + ks * 1024, 0, 512);
      if (TMEM_A) mma_tf32_ta(tb, a_tmem + ks * 8, db, idesc, ks > 0);
      else tc::mma_tf32(tb, tc::make_smem_desc(tc::smem_addr(a_s) + ks * 2 * 2048, 2048, 128), db, idesc, ks > 0);
    }
    tc::mma_commit(&bar);
  }
  wait_bar(&bar, 0, "T1/T4");
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 32 + n] = v[n];
  tc::fence_before_thread_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tb);
}

// ---------------------------------------------------------------------------------------------- T2
// D[128][64] = sum_{k<24} A[k][m] B[n][k] - sum_{k>=24} A[k][m] B[n][k];  A MN-major SW128 in 4 groups of 32 rows.
__global__ void __launch_bounds__(128) probe_a_mn(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_s = smem;                                   // 4 groups x 6144 B
  float* b_s = reinterpret_cast<float*>(smem + 4 * 6144);      // K-major, 64 x 48
  __shared__ uint32_t tb_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<64>(&tb_s);
  for (int e = tid; e < 48 * 128; e += 128) {
    const int k = e / 128, m = e % 128;
    *reinterpret_cast<float*>(a_s + (m / 32) * 6144 + swz128_32(k * 128 + (m % 32) * 4)) = A[e];   // A[k][m]
  }
  for (int e = tid; e < 64 * 48; e += 128) b_s[tc::kmajor_offset(e / 48, e % 48, 64) / 4] = B[e];   // B[n][k]
  tc::fence_proxy_async_smem(); tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  const uint32_t tb = tb_s;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, 64) | kAMajorMN;
    for (int ks = 0; ks < 6; ++ks) {
      const uint64_t da = desc_sw128_32(tc::smem_addr(a_s) + ks * 1024, 6144, 512);
      const uint64_t db = tc::make_smem_desc(tc::smem_addr(b_s) + ks * 2 * 1024, 1024, 128);
      tc::mma_tf32(tb, da, db, ks >= 3 ? (idesc | kANeg) : idesc, ks > 0);
    }
    tc::mma_commit(&bar);
  }
  wait_bar(&bar, 0, "T2");
  tc::fence_after_thread_sync();
  for (int c = 0; c < 2; ++c) {
    float v[32];
    tc::tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
    for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 64 + c * 32 + n] = v[n];
  }
  tc::fence_before_thread_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc<64>(tb);
}

// ---------------------------------------------------------------------------------------------- T3
// D[128 px][32 o] = E[128][8] Z[8][32] (tf32)  +  X[ch][px0 + m] W[o][ch] (bf16 A from TMA, B bf16 or fp16)
template <int B_F16>
__global__ void __launch_bounds__(128) probe_tma(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ E,
                                                 const float* __restrict__ Z, const void* __restrict__ W, float* __restrict__ D,
                                                 int px0) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* x_s = smem;                                    // 2 groups x 4096 B (TMA, swizzled)
  unsigned char* w_s = smem + 8192;                             // bf16/fp16 K-major [32 n][32 k], 2 KB
  float* e_s = reinterpret_cast<float*>(smem + 8192 + 2048);    // K-major 128 x 8
  unsigned char* z_s = smem + 8192 + 2048 + 4096;               // 8 rows x 128 B swizzled (1024-aligned: 14336)
  __shared__ uint32_t tb_s;
  __shared__ __align__(8) uint64_t bar, tbar;
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&tbar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<32>(&tb_s);
  for (int e = tid; e < 32 * 32; e += 128) {
    const int n = e / 32, k = e % 32;
    reinterpret_cast<uint16_t*>(w_s)[(((k >> 3) * 4 + (n >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2) / 2] =
        reinterpret_cast<const uint16_t*>(W)[e];
  }
  for (int e = tid; e < 128 * 8; e += 128) e_s[tc::kmajor_offset(e / 8, e % 8, 128) / 4] = E[e];
  for (int e = tid; e < 8 * 32; e += 128) *reinterpret_cast<float*>(z_s + swz128_32((e / 32) * 128 + (e % 32) * 4)) = Z[e];
  tc::fence_proxy_async_smem(); tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  const uint32_t tb = tb_s;
  if (tid == 0) {
    mbar_expect_tx(&tbar, 8192);
    tma_load_2d(x_s, &tmap, px0, 0, &tbar);
    tma_load_2d(x_s + 4096, &tmap, px0 + 64, 0, &tbar);
    wait_bar(&tbar, 0, "T3 tma");
    tc::mma_tf32(tb, tc::make_smem_desc(tc::smem_addr(e_s), 2048, 128), desc_sw128_32(tc::smem_addr(z_s), 0, 512),
                 tc::make_idesc_tf32(128, 32) | kBMajorMN, false);
    const uint32_t idesc = idesc_f16(128, 32, 1, B_F16 ? 0 : 1) | kAMajorMN;
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t da = desc_sw128(tc::smem_addr(x_s) + ks * 2048, 4096, 1024);
      const uint64_t db = tc::make_smem_desc(tc::smem_addr(w_s) + ks * 2 * 512, 512, 128);
      mma_f16(tb, da, db, idesc, true);
    }
    tc::mma_commit(&bar);
  }
  wait_bar(&bar, 0, "T3");
  tc::fence_after_thread_sync();
  float v[32];
  tc::tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16), v);
  for (int n = 0; n < 32; ++n) D[(warp * 32 + lane) * 32 + n] = v[n];
  tc::fence_before_thread_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc<32>(tb);
}

// ---------------------------------------------------------------------------------------------- T5 timings
// MODE 0: tf32 SS, A K-major, B MN-SW128 (N=32)     MODE 1: tf32 TMEM-A, B MN-SW128 (N=32)
// MODE 2: bf16 SS, A MN-SW128, B K-major (N=32,K16)  MODE 3: tf32 SS, A MN-SW128 4 groups, B K-major (N=64)
// MODE 4: tf32 TMEM-A, B MN-SW128 N=128 (4 groups)   MODE 5: bf16 SS N=32 with 2 alternating accumulators
template <int MODE, int NMMA>
__global__ void __launch_bounds__(128) timing(long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tb_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tc::warp_index_uniform();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tc::tmem_alloc<512>(&tb_s);
  for (int e = tid; e < 49152 / 4; e += 128) reinterpret_cast<float*>(smem)[e] = 0.f;
  tc::fence_proxy_async_smem(); tc::fence_before_thread_sync(); __syncthreads(); tc::fence_after_thread_sync();
  const uint32_t tb = tb_s;
  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t s0 = tc::smem_addr(smem);
      for (int rep = 0; rep < 3; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NMMA; ++i) {
          const bool acc = i > 0;
          if (MODE == 0) tc::mma_tf32(tb, tc::make_smem_desc(s0 + 8192, 2048, 128), desc_sw128_32(s0, 0, 512), tc::make_idesc_tf32(128, 32) | kBMajorMN, acc);
          if (MODE == 1) mma_tf32_ta(tb, tb + 256, desc_sw128_32(s0, 0, 512), tc::make_idesc_tf32(128, 32) | kBMajorMN, acc);
          if (MODE == 2) mma_f16(tb, desc_sw128(s0, 4096, 1024), tc::make_smem_desc(s0 + 16384, 512, 128), idesc_f16(128, 32, 1, 1) | kAMajorMN, acc);
          if (MODE == 3) tc::mma_tf32(tb, desc_sw128_32(s0, 6144, 512), tc::make_smem_desc(s0 + 32768, 1024, 128), tc::make_idesc_tf32(128, 64) | kAMajorMN, acc);
          if (MODE == 4) mma_tf32_ta(tb, tb + 256, desc_sw128_32(s0, 6144, 512), tc::make_idesc_tf32(128, 128) | kBMajorMN, acc);
          if (MODE == 5) mma_f16(tb + (i & 1) * 32, desc_sw128(s0, 4096, 1024), tc::make_smem_desc(s0 + 16384, 512, 128), idesc_f16(128, 32, 1, 1) | kAMajorMN, i > 1);
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
  tc::fence_before_thread_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tb);
}
template <int MODE>
void run_timing(long long* d, const char* what) {
  long long h[6];
  cudaFuncSetAttribute(timing<MODE, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  timing<MODE, 32><<<1, 128, 49152>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("T5 %s: %s\n", what, cudaGetErrorString(e)); return; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("T5 %-58s issue %5lld cyc, complete %5lld cyc = %5.1f cyc/MMA\n", what, h[4], h[5], (double)h[5] / 32);
}

static float rnd8() { return (float)((rand() % 33) - 16) / 8.f; }

template <typename F>
static int check(const char* name, const std::vector<float>& D, int M, int N, F ref) {
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) maxerr = fmax(maxerr, fabs(ref(m, n) - (double)D[m * N + n]));
  printf("%s: max abs err %.3e %s\n", name, maxerr, maxerr == 0 ? "OK" : "FAILED");
  if (maxerr != 0) printf("   D[0][0..3] = %f %f %f %f   ref %f %f %f %f\n", D[0], D[1], D[2], D[3], ref(0, 0), ref(0, 1), ref(0, 2), ref(0, 3));
  return maxerr != 0;
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main() {
  int fails = 0;
  srand(5);
  // ---------------- T1 / T4
  for (int tmem_a = 0; tmem_a < 2; ++tmem_a) {
    std::vector<float> A(128 * 48), B(48 * 32), D(128 * 32);
    for (auto& x : A) x = rnd8();
    for (auto& x : B) x = rnd8();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 6144 + 128 * 48 * 4;
    if (tmem_a) probe_b_mn<1><<<1, 128, smem>>>(dA, dB, dD); else probe_b_mn<0><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T1/T4: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    fails += check(tmem_a ? "T4 tf32 A in TMEM, B MN-major SW128" : "T1 tf32 A K-major smem, B MN-major SW128", D, 128, 32,
                   [&](int m, int n) { double r = 0; for (int k = 0; k < 48; ++k) r += (double)A[m * 48 + k] * B[k * 32 + n]; return r; });
  }
  // ---------------- T2
  {
    std::vector<float> A(48 * 128), B(64 * 48), D(128 * 64);
    for (auto& x : A) x = rnd8();
    for (auto& x : B) x = rnd8();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 4 * 6144 + 64 * 48 * 4;
    probe_a_mn<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T2: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    fails += check("T2 tf32 A MN-major SW128 (4 groups, LBO), a_negate on K-steps 3..5", D, 128, 64, [&](int m, int n) {
      double r = 0;
      for (int k = 0; k < 48; ++k) r += (k < 24 ? 1.0 : -1.0) * A[k * 128 + m] * B[n * 48 + k];
      return r;
    });
  }
  // ---------------- T3
  {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { printf("T3: cuTensorMapEncodeTiled not found\n"); return 2; }
    const int NPX = 4096, NCH = 32, px0 = 256;
    std::vector<__nv_bfloat16> X(NCH * NPX);
    std::vector<float> Xf(NCH * NPX), Wf(32 * 32), E(128 * 8), Z(8 * 32), D(128 * 32);
    for (int i = 0; i < NCH * NPX; ++i) { Xf[i] = rnd8(); X[i] = __float2bfloat16(Xf[i]); }
    for (auto& x : Wf) x = rnd8();
    for (auto& x : E) x = rnd8();
    for (auto& x : Z) x = rnd8();
    __nv_bfloat16* dX; void* dW; float *dE, *dZ, *dD;
    CK(cudaMalloc(&dX, X.size() * 2)); CK(cudaMalloc(&dW, 32 * 32 * 2)); CK(cudaMalloc(&dE, E.size() * 4));
    CK(cudaMalloc(&dZ, Z.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dE, E.data(), E.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dZ, Z.data(), Z.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {NPX, NCH}, gstride[1] = {NPX * 2};
    const cuuint32_t box[2] = {64, 32}, estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dX, gdim, gstride, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("T3: cuTensorMapEncodeTiled failed (%d)\n", (int)r); return 2; }
    for (int bf16b = 1; bf16b >= 1; --bf16b) {  // bf16 x fp16 (T3b) is an illegal instruction on sm_100a: measured
      std::vector<uint16_t> Wh(32 * 32);
      for (int i = 0; i < 32 * 32; ++i) {
        if (bf16b) { __nv_bfloat16 t = __float2bfloat16(Wf[i]); Wh[i] = *reinterpret_cast<uint16_t*>(&t); }
        else { __half t = __float2half(Wf[i]); Wh[i] = *reinterpret_cast<uint16_t*>(&t); }
      }
      CK(cudaMemcpy(dW, Wh.data(), Wh.size() * 2, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0xff, D.size() * 4));
      const int smem = 8192 + 2048 + 4096 + 1024;
      if (bf16b) probe_tma<0><<<1, 128, smem>>>(tmap, dE, dZ, dW, dD, px0); else probe_tma<1><<<1, 128, smem>>>(tmap, dE, dZ, dW, dD, px0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("T3: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      fails += check(bf16b ? "T3  tf32 chain + kind::f16 (bf16 A via TMA MN-major SW128, bf16 B) in one accumulator"
                           : "T3b same with an fp16 B operand (mixed bf16 x fp16)",
                     D, 128, 32, [&](int m, int n) {
                       double r = 0;
                       for (int k = 0; k < 8; ++k) r += (double)E[m * 8 + k] * Z[k * 32 + n];
                       for (int c = 0; c < 32; ++c) r += (double)Xf[c * NPX + px0 + m] * Wf[n * 32 + c];
                       return r;
                     });
    }
  }
  // ---------------- T5
  long long* d;
  CK(cudaMalloc(&d, 64));
  run_timing<0>(d, "tf32 SS   M128 N32  K8  (A K-major, B MN-SW128)");
  run_timing<1>(d, "tf32 TMEM-A M128 N32 K8 (B MN-SW128)");
  run_timing<2>(d, "bf16 SS   M128 N32  K16 (A MN-SW128, B K-major)");
  run_timing<5>(d, "bf16 SS   M128 N32  K16, two accumulators");
  run_timing<3>(d, "tf32 SS   M128 N64  K8  (A MN-SW128 x4 groups, B K-major)");
  run_timing<4>(d, "tf32 TMEM-A M128 N128 K8 (B MN-SW128 x4 groups)");
  printf(fails ? "PROBE5 FAILED (%d)\n" : "PROBE5 OK\n", fails);
  return fails;
}
