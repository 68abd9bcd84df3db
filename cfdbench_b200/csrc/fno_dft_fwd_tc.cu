// K1 (tensor-core stage 1) -- truncated forward 2-D DFT of activation planes: x[b][c][64][64] -> Xm[b][k][c].
//
// Replaces torch.fft.rfft2 + the two corner slices of the reference (src/models/fno/fno2d.py:62,73-78): only
// kx in {0..11, 52..63} x ky in {0..11} is ever used, so the 64x33 spectrum is never materialised.
//
// Stage 1 (real DFT along h, bins kx' = 0..12) is a GEMM with the constant matrix F[(kx',re|im)][h]:
//     D[128 = 2 planes x 64 w][32 = (kx', re|im) padded] = A[(p, w)][h] * F^T     (tcgen05.mma kind::tf32, 3xTF32)
// the A operand is the activation tile itself: threads load x[p][4q..4q+3][w] (coalesced along w), split into tf32
// hi/lo (bf16 activations are tf32-exact: no lo pass) and store 16-byte K-major operand rows.  The accumulator is read
// back from TMEM (thread = column w) and written as the 13 complex rows A[kx'][w] of each plane in shared memory.
// Stage 2 (complex DFT along w, bins -11..11) stays on the CUDA cores: 4 threads per row, each running a pruned
// 64-point DIF codelet for the bins = j (mod 4) (warp-uniform j, no exchange); X[64-kx', ky] = conj(F[kx'][-ky]).
// Persistent CTA of two independent 256-thread pipelines, tile = 2 planes, operands prefetched one tile ahead.
#include "fft_codelets.cuh"
#include "fno_common.cuh"
#include "tc_common.cuh"
#include <math.h>
#include <stddef.h>
#include <string.h>

namespace fno {

constexpr int kD1Threads = 512;
constexpr int kD1Group = 256;
constexpr int kD1M = 128;                        // 2 planes x 64 columns
constexpr int kD1K = kH;                         // 64 rows h
constexpr int kD1N = 32;                         // 13 kx' x (re, im) = 26, padded to 32
constexpr uint32_t kD1LboA = (kD1M / 8) * 128;   // 2048
constexpr uint32_t kD1LboB = (kD1N / 8) * 128;   // 512
constexpr int kD1Rows = 26;                      // complex rows per tile (2 planes x 13)
constexpr int kD1Pitch = 65;                     // float2 elements; +1 keeps stage-2 row gathers conflict-free
constexpr int kD1TilesPerSample = kC / 2;        // 16
constexpr int kFTabFloats = 2 * kD1N * kD1K;     // hi image then lo image (16 KB)

struct D1Smem {
  alignas(128) float f_hi[kD1N * kD1K];          // B operand (constant)            8 KB
  alignas(128) float f_lo[kD1N * kD1K];
  alignas(128) float a_hi[2][kD1M * kD1K];       // A operand per pipeline          2 x 32 KB
  alignas(128) float a_lo[2][kD1M * kD1K];
  alignas(16) float2 as[2][2][kD1Rows * kD1Pitch];  // [pipeline][buffer] stage-1 output rows   4 x 13.2 KB
  alignas(8) uint64_t mma_bar[2][2];
  alignas(8) uint64_t ftab_bar;
  uint32_t tmem_base;
};

// bins (in codelet output order) produced by cfft64_r<J>: the members of {0..11, 53..63} that are = J mod 4
template <int J>
__host__ __device__ constexpr int d1_bin_count() {
  return J == 0 ? 5 : 6;
}
template <int J>
__host__ __device__ constexpr int d1_bin(int e) {
  return e < 3 ? 4 * e + J : (J == 0 ? 4 * e + 44 : 4 * e + J + 40);
}

template <int J>
__device__ __forceinline__ void d1_row_transform_and_emit(const float2* __restrict__ row, float2* __restrict__ xm_b,
                                                          int kxp, int c, float s0, float s1) {
  float xre[64], xim[64], ore[6], oim[6];
#pragma unroll
  for (int n = 0; n < 64; ++n) {
    const float2 v = row[n];
    xre[n] = v.x;
    xim[n] = v.y;
  }
  if constexpr (J == 0) fno_codelets::cfft64_r0<float>(xre, xim, ore, oim);
  if constexpr (J == 1) fno_codelets::cfft64_r1<float>(xre, xim, ore, oim);
  if constexpr (J == 2) fno_codelets::cfft64_r2<float>(xre, xim, ore, oim);
  if constexpr (J == 3) fno_codelets::cfft64_r3<float>(xre, xim, ore, oim);
#pragma unroll
  for (int e = 0; e < d1_bin_count<J>(); ++e) {
    const int q = d1_bin<J>(e);
    if (q <= 11) {  // X[kx', q] = F[kx'][q], rows 0..11 (weights1 block)
      if (kxp <= 11) {
        const float s = (q == 0) ? s0 : s1;
        xm_b[(kxp * kM2 + q) * kC + c] = make_float2(ore[e] * s, oim[e] * s);
      }
    }
    const int qq = (64 - q) & 63;
    if (qq <= 11) {  // X[64-kx', qq] = conj(F[kx'][-qq]), rows 52..63 (weights2 block)
      if (kxp >= 1) {
        const float s = (qq == 0) ? s0 : s1;
        xm_b[((kKX - kxp) * kM2 + qq) * kC + c] = make_float2(ore[e] * s, -oim[e] * s);
      }
    }
  }
}

template <typename TAct>
struct D1Regs {
  TAct v[8][4];  // task = rep*256 + gtid -> (m = task & 127 = (plane, w), h quad = task >> 7)
};

__device__ __forceinline__ float d1_to_float(float v) { return v; }
__device__ __forceinline__ float d1_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename TAct>
__device__ __forceinline__ void d1_prefetch(D1Regs<TAct>& r, const TAct* __restrict__ x, int tile, int gtid) {
  const TAct* base = x + static_cast<size_t>(tile) * 2 * kHW;  // tile = (b*32 + c0)/2: two consecutive planes
#pragma unroll
  for (int rep = 0; rep < 8; ++rep) {
    const int task = rep * kD1Group + gtid;
    const int m = task & (kD1M - 1), hq = task >> 7;
    const TAct* src = base + (m >> 6) * kHW + (4 * hq) * kW + (m & 63);
#pragma unroll
    for (int c = 0; c < 4; ++c) r.v[rep][c] = __ldg(src + c * kW);
  }
}

template <typename TAct>
__device__ __forceinline__ void d1_split_store(const D1Regs<TAct>& r, float* a_hi, float* a_lo, int gtid) {
#pragma unroll
  for (int rep = 0; rep < 8; ++rep) {
    const int task = rep * kD1Group + gtid;
    const int m = task & (kD1M - 1), hq = task >> 7;
    float hi[4], lo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float v = d1_to_float(r.v[rep][c]);
      if constexpr (sizeof(TAct) == 4) tc::split_tf32(v, hi[c], lo[c]);
      else hi[c] = v;
    }
    const uint32_t off = tc::kmajor_offset(m, 4 * hq, kD1M) / 4;
    *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    if constexpr (sizeof(TAct) == 4) *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <int GRP>
__device__ __forceinline__ void d1_group_barrier() {
  asm volatile("bar.sync %0, %1;" ::"n"(GRP + 1), "n"(kD1Group) : "memory");
}

template <typename TAct, int GRP>
__device__ __forceinline__ void d1_pipeline(D1Smem& sm, const TAct* __restrict__ x, float2* __restrict__ xm, float s0,
                                            float s1, int n_tiles) {
  constexpr bool kBf16 = sizeof(TAct) == 2;
  const int tid = threadIdx.x, lane = tid & 31;
  const int gtid = tid & (kD1Group - 1), gwarp = (tid >> 5) & 7;
  const uint32_t tmem_base = sm.tmem_base + GRP * (2 * kD1N);
  constexpr uint32_t idesc = tc::make_idesc_tf32(kD1M, kD1N);

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_cta = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;
  const int n_mine = (n_cta + 1 - GRP) / 2;
  auto tile_of = [&](int it) { return first + (2 * it + GRP) * stride; };

  // consume tile `it`: TMEM accumulator -> 13 complex rows per plane in smem -> stage 2 -> modes in global memory
  auto consume = [&](int it) {
    const int buf = it & 1;
    mbar_wait(&sm.mma_bar[GRP][buf], (it >> 1) & 1);
    tc::fence_after_thread_sync();
    float2* as = sm.as[GRP][buf];
    {
      const int quad = gwarp & 3, half = gwarp >> 2;  // TMEM lane quadrant, column half (kx' 0..7 / 8..12)
      const int m = quad * 32 + lane;                 // (plane, w)
      uint32_t rr[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kD1N + half * 16;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
            "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float2* dst = as + ((m >> 6) * 13 + half * 8) * kD1Pitch + (m & 63);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (half == 0 || j < 5) dst[j * kD1Pitch] = make_float2(__uint_as_float(rr[2 * j]), __uint_as_float(rr[2 * j + 1]));
    }
    tc::fence_before_thread_sync();
    d1_group_barrier<GRP>();
    if (gwarp < 4 && lane < kD1Rows) {  // stage 2: row rho = lane = plane*13 + kx', bins = gwarp (mod 4)
      const int tile = tile_of(it);
      const int b = tile / kD1TilesPerSample, c = 2 * (tile % kD1TilesPerSample) + lane / 13;
      const int kxp = lane % 13;
      const float2* row = as + lane * kD1Pitch;
      float2* xm_b = xm + static_cast<size_t>(b) * kModes * kC;
      switch (gwarp) {
        case 0: d1_row_transform_and_emit<0>(row, xm_b, kxp, c, s0, s1); break;
        case 1: d1_row_transform_and_emit<1>(row, xm_b, kxp, c, s0, s1); break;
        case 2: d1_row_transform_and_emit<2>(row, xm_b, kxp, c, s0, s1); break;
        default: d1_row_transform_and_emit<3>(row, xm_b, kxp, c, s0, s1); break;
      }
    }
  };

  D1Regs<TAct> regs;
  if (n_mine > 0) d1_prefetch<TAct>(regs, x, tile_of(0), gtid);

  for (int it = 0; it < n_mine; ++it) {
    const int buf = it & 1;
    // the single-buffered A operand was last read by the MMAs of tile it-1: wait for them (normally long done)
    if (it >= 1) mbar_wait(&sm.mma_bar[GRP][(it - 1) & 1], ((it - 1) >> 1) & 1);
    d1_split_store<TAct>(regs, sm.a_hi[GRP], sm.a_lo[GRP], gtid);
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    d1_group_barrier<GRP>();
    tc::fence_after_thread_sync();
    if (it + 1 < n_mine) d1_prefetch<TAct>(regs, x, tile_of(it + 1), gtid);
    if (gwarp == 0) {
      if (tc::elect_one()) {
        const uint32_t d_tmem = tmem_base + buf * kD1N;
        const uint32_t a_s[3] = {tc::smem_addr(sm.a_hi[GRP]), tc::smem_addr(sm.a_lo[GRP]), tc::smem_addr(sm.a_hi[GRP])};
        const uint32_t b_s[3] = {tc::smem_addr(sm.f_hi), tc::smem_addr(sm.f_hi), tc::smem_addr(sm.f_lo)};
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          if (kBf16 && pass == 1) continue;  // bf16 activations have no lo part
          const uint64_t da0 = tc::make_smem_desc(a_s[pass], kD1LboA, 128);
          const uint64_t db0 = tc::make_smem_desc(b_s[pass], kD1LboB, 128);
#pragma unroll
          for (int ks = 0; ks < kD1K / 8; ++ks) {
            const uint64_t da = da0 + ((ks * 2 * kD1LboA) >> 4), db = db0 + ((ks * 2 * kD1LboB) >> 4);
            if (pass == 0 && ks == 0) tc::mma_tf32_imm<false>(d_tmem, da, db, idesc);
            else tc::mma_tf32_imm<true>(d_tmem, da, db, idesc);
          }
        }
        tc::mma_commit(&sm.mma_bar[GRP][buf]);
      }
      __syncwarp();
    }
    if (it >= 1) consume(it - 1);
  }
  if (n_mine >= 1) consume(n_mine - 1);
}

template <typename TAct>
__global__ void __launch_bounds__(kD1Threads, 1)
    dft_fwd_tc_kernel(const TAct* __restrict__ x, float2* __restrict__ xm, const float* __restrict__ ftab, float s0,
                      float s1, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // no pointer arithmetic: keeps LDS/STS addressing
  D1Smem& sm = *reinterpret_cast<D1Smem*>(smem_raw);
  if ((smem_u32(smem_raw) & 127u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5;
  const int grp = warp >> 3;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&sm.mma_bar[i >> 1][i & 1], 1);
    mbar_init(&sm.ftab_bar, 1);
    fence_mbar_init();
    constexpr uint32_t kBytes = kFTabFloats * sizeof(float);
    static_assert(offsetof(D1Smem, f_lo) == offsetof(D1Smem, f_hi) + kBytes / 2, "f_hi / f_lo must be contiguous");
    mbar_expect_tx(&sm.ftab_bar, kBytes);
    bulk_g2s(sm.f_hi, ftab, kBytes, &sm.ftab_bar);
  }
  if (warp == 0) tc::tmem_alloc<4 * kD1N>(&sm.tmem_base);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  mbar_wait(&sm.ftab_bar, 0);
  if (grp == 0) d1_pipeline<TAct, 0>(sm, x, xm, s0, s1, n_tiles);
  else d1_pipeline<TAct, 1>(sm, x, xm, s0, s1, n_tiles);
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<4 * kD1N>(sm.tmem_base);
}

// Constant B operand: F[n = 2 kx' + ri][h] = cos(2 pi kx' h/64) (ri = 0), -sin(2 pi kx' h/64) (ri = 1), kx' = 0..12,
// rows 26..31 zero; float64 -> tf32 hi/lo (round to nearest), K-major image, built once per device.
static float d1_round_tf32_host(double v) {
  float f = static_cast<float>(v);
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&f, &u, 4);
  return f;
}

static float* g_ftab[64] = {nullptr};

static cudaError_t ensure_ftab(const float** out, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (g_ftab[dev] == nullptr) {
    static float host[kFTabFloats];
    for (int i = 0; i < kFTabFloats; ++i) host[i] = 0.f;
    for (int kx = 0; kx <= 12; ++kx)
      for (int h = 0; h < kH; ++h) {
        const double ang = 2.0 * 3.14159265358979323846 * ((kx * h) % 64) / 64.0;
        const double val[2] = {cos(ang), -sin(ang)};
        for (int ri = 0; ri < 2; ++ri) {
          const float hi = d1_round_tf32_host(val[ri]);
          const float lo = d1_round_tf32_host(val[ri] - static_cast<double>(hi));
          const uint32_t off = tc::kmajor_offset(2 * kx + ri, h, kD1N) / 4;
          host[off] = hi;
          host[kD1N * kD1K + off] = lo;
        }
      }
    float* d = nullptr;
    e = cudaMalloc(&d, sizeof(host));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d, host, sizeof(host), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    g_ftab[dev] = d;
  }
  *out = g_ftab[dev];
  return cudaSuccess;
}

template <typename TAct>
cudaError_t launch_dft_fwd_tc(const void* x, void* xm, int batch, float s0, float s1, cudaStream_t stream) {
  auto kern = dft_fwd_tc_kernel<TAct>;
  constexpr size_t smem = sizeof(D1Smem);
  static bool configured = false;
  static int n_sm = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const float* ftab = nullptr;
  cudaError_t e = ensure_ftab(&ftab, stream);
  if (e != cudaSuccess) return e;
  const int n_tiles = batch * kD1TilesPerSample;
  const int grid = n_tiles < 2 * n_sm ? (n_tiles + 1) / 2 : n_sm;
  kern<<<grid, kD1Threads, smem, stream>>>(static_cast<const TAct*>(x), static_cast<float2*>(xm), ftab, s0, s1, n_tiles);
  return cudaGetLastError();
}

template cudaError_t launch_dft_fwd_tc<float>(const void*, void*, int, float, float, cudaStream_t);
template cudaError_t launch_dft_fwd_tc<__nv_bfloat16>(const void*, void*, int, float, float, cudaStream_t);

}  // namespace fno
