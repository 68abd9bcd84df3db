// Fused Fourier-block output stage for bf16 activation storage: ONE kernel replaces inv_kx_kernel + block_tc_kernel
//   out[b][o][h][w] = GELU( irfft2(pad(Y))[b][o][h][w] + sum_i W0[o][i] x[b][i][h][w] + bias[o] )      (bf16 in, bf16 out)
// i.e. irfft2 + Conv2d(32,32,1) + add + GELU of the reference FnoBlock (src/models/fno/fno2d.py:65-72,81,104-111).
// The half-inverted spectrum Z never leaves the SM (it used to be written and re-read through HBM, 50 MB per layer at
// B = 256) and no operand is staged through registers on the load side:
//
//   work unit  = (sample, half image of 32 rows) = 16 tiles of 128 pixels; units are dealt round-robin to the CTAs.
//   GEMM1      inverse DFT along kx of the unit's 32 rows, on the tensor cores, 3xTF32:
//                  D1[(ky,o) 3 x 128 lanes][(h',re|im) 64 cols] = Y^T[(ky,o)][(kx,re|im) 48] * F[(kx,re|im)][(h',re|im)]
//              A = the mixed modes of the sample exactly as mode_mix_tc_kernel's epilogue wrote them: tf32 hi/lo images,
//                  MN-major, 128B/32B-base swizzle (the only MN-major form kind::tf32 accepts: tools/tc_probe5.cu), bulk-copied
//                  24 KB at a time;   B = constant twiddles (K-major), second half image = first with odd kx negated
//                  (a_negate bit of the instruction descriptor).
//   converters 12 warps, thread = (ky, o): pull D1 out of tensor memory (tcgen05.ld), split into tf32 hi/lo and write
//              the per-tile B operand  Zt[(j,ky,re|im) 48][o 32]  (MN-major: a warp writes whole 128-byte rows).
//   tile MMA   D2[128 px][32 o] = X[128 px][32 i] W0^T          kind::f16: x tile arrives by TMA (tensor map, 128B swizzle)
//                                                               straight from the NCHW bf16 activation (bf16 is exact),
//                                                               W0 as three bf16 pieces (24 significant bits)
//                               + (E (+) E)[128 px][48] Zt       kind::tf32, 3xTF32; E = C2R stage (cos,-sin)(2 pi ky w/64)
//                                                               c_ky/HW folded in, resident in TENSOR MEMORY as the A operand;
//                                                               its (ky=0, Im) column is 1 and the matching Zt row = bias.
//   epilogue   8 warps: TMEM -> registers -> exact-erf GELU -> bf16 -> global.
// Warp roles (704 threads): 0-11 converters, 12-19 epilogue, 20 MMA issue (one elected lane), 21 producers (lane 0: x tiles
// by TMA, lane 1: mode images by bulk copy).  All hand-offs are mbarriers; rings: 8 x tiles, 6 Zt operands, 4 accumulators.
#include "fno_common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <math.h>
#include <string.h>

namespace fno {

constexpr int kFzThreads = 704;
constexpr int kFzConvWarps = 12, kFzEpiWarps = 8;
constexpr int kFzMmaWarp = 20, kFzProdWarp = 21;
constexpr int kFzTilesPerUnit = 16;         // 32 rows / 2
constexpr int kFzNX = 8, kFzNB = 6, kFzND = 4, kFzNY = 2;
constexpr uint32_t kFzXBytes = 8192;        // 2 boxes x (32 ch x 128 B)
constexpr uint32_t kFzBtBytes = 12288;      // hi + lo, 48 rows x 128 B each
constexpr uint32_t kFzYStage = 24576;       // (M-tile, kx parity): hi + lo, 4 ky groups x 24 rows x 128 B
constexpr uint32_t kFzFBytes = 24576;       // twiddle operand: hi + lo images of [64][48]
constexpr uint32_t kFzWBytes = 3 * 2048;    // three bf16 pieces of W0
// mode image of one sample (written by mode_mix_tc_kernel): [hi|lo][M-tile 3][ky group 4][48 rows][32 o] fp32
constexpr size_t kYmImgPart = 73728, kYmImgMtile = 24576, kYmImgGroup = 6144;
constexpr size_t kYmImgBytes = 2 * kYmImgPart;
// tensor memory columns
constexpr uint32_t kFzColE = 0;      // E hi (48) | E lo (48)
constexpr uint32_t kFzColD1 = 96;    // 3 x 64
constexpr uint32_t kFzColD2 = 288;   // 4 x 32

struct FzSmem {
  alignas(1024) unsigned char x[kFzNX][kFzXBytes];
  alignas(1024) unsigned char bt[kFzNB][kFzBtBytes];
  alignas(1024) unsigned char y[kFzNY][kFzYStage];
  alignas(1024) unsigned char f[kFzFBytes];
  alignas(1024) unsigned char w[kFzWBytes];
  alignas(16) float bias[kC];
  alignas(8) uint64_t x_full[kFzNX], x_empty[kFzNX];
  uint64_t bt_full[kFzNB], bt_empty[kFzNB];
  uint64_t d2_full[kFzND], d2_empty[kFzND];
  uint64_t y_full[kFzNY], y_empty[kFzNY];
  uint64_t d1_full, d1_free, f_bar;
  uint32_t tmem_base;
};

constexpr uint32_t kAMajorMN = 1u << 15, kBMajorMN = 1u << 16, kANegate = 1u << 13;
__host__ __device__ constexpr uint32_t fz_idesc_bf16(int m, int n) {  // D f32, A/B bf16
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ uint64_t fz_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {      // 16-bit MN-major
  return tc::make_smem_desc(saddr, lbo, sbo) | (static_cast<uint64_t>(2) << 61);
}
__device__ __forceinline__ uint64_t fz_desc_sw128_32(uint32_t saddr, uint32_t lbo, uint32_t sbo) {   // 32-bit MN-major
  return tc::make_smem_desc(saddr, lbo, sbo) | (static_cast<uint64_t>(1) << 61);
}
__device__ __forceinline__ void fz_mma_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fz_mma_tf32_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fz_mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void fz_tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fz_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(kFzThreads, 1)
    block_fused_kernel(const __grid_constant__ CUtensorMap x_map, const unsigned char* __restrict__ ym_img,
                       const float* __restrict__ w0t, const float* __restrict__ bias, const float* __restrict__ etab,
                       const float* __restrict__ ftab, __nv_bfloat16* __restrict__ out, int n_units) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  FzSmem& sm = *reinterpret_cast<FzSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = (first < n_units) ? (n_units - first + stride - 1) / stride : 0;
  auto unit_of = [&](int k) { return first + k * stride; };   // unit u = 2 * sample + half

  // ---------------------------------------------------------------- prologue (weights / constant tables only)
  if (tid == 0) {
    for (int i = 0; i < kFzNX; ++i) { mbar_init(&sm.x_full[i], 1); mbar_init(&sm.x_empty[i], 1); }
    for (int i = 0; i < kFzNB; ++i) { mbar_init(&sm.bt_full[i], kFzConvWarps); mbar_init(&sm.bt_empty[i], 1); }
    for (int i = 0; i < kFzND; ++i) { mbar_init(&sm.d2_full[i], 1); mbar_init(&sm.d2_empty[i], kFzEpiWarps); }
    for (int i = 0; i < kFzNY; ++i) { mbar_init(&sm.y_full[i], 1); mbar_init(&sm.y_empty[i], 1); }
    mbar_init(&sm.d1_full, 1);
    mbar_init(&sm.d1_free, kFzConvWarps);
    mbar_init(&sm.f_bar, 1);
    fence_mbar_init();
    mbar_expect_tx(&sm.f_bar, kFzFBytes);
    bulk_g2s(sm.f, ftab, kFzFBytes, &sm.f_bar);   // twiddle operand, already in its K-major tf32 hi|lo layout
  }
  if (warp == kFzMmaWarp) tc::tmem_alloc<512>(&sm.tmem_base);
  // W0 -> three bf16 pieces, B operand [n = o][k = i], K-major, no swizzle (8 x 16-byte core matrices)
  for (int e = tid; e < kC * kC; e += kFzThreads) {
    const int i = e / kC, o = e % kC;   // w0t[i][o] = W0[o][i]
    const float wv = w0t[e];
    const __nv_bfloat16 p0 = __float2bfloat16_rn(wv);
    const float r1 = wv - __bfloat162float(p0);
    const __nv_bfloat16 p1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
    const uint32_t off = ((i >> 3) * 4 + (o >> 3)) * 128 + (o & 7) * 16 + (i & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sm.w + off) = p0;
    *reinterpret_cast<__nv_bfloat16*>(sm.w + 2048 + off) = p1;
    *reinterpret_cast<__nv_bfloat16*>(sm.w + 4096 + off) = p2;
  }
  if (tid < kC) sm.bias[tid] = bias ? bias[tid] : 0.f;
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = sm.tmem_base;
  if (warp >= kFzConvWarps && warp < kFzConvWarps + 4) {   // constant E operand -> tensor memory (row m in lane m)
    const int m = (warp & 3) * 32 + lane;
    const float* row = etab + m * 96;
#pragma unroll
    for (int c0 = 0; c0 < 96; c0 += 16) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __ldg(row + c0 + j);
      tc::tmem_st16(tmem + kFzColE + c0 + (static_cast<uint32_t>((warp & 3) * 32) << 16), v);
    }
    tc::tmem_wait_st();
  }
  mbar_wait(&sm.f_bar, 0);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  pdl_wait();   // ym_img and x come from the previous kernels of the chain
  pdl_launch_dependents();

  // ================================================================ converters
  if (warp < kFzConvWarps) {
    const int mt = warp >> 2, q = warp & 3, ky = 4 * mt + q, o = lane;
    const float bias_o = sm.bias[o];
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    // byte offsets of this thread's 4 operand rows (j, ky, ri) at column o; row k = 24 j + 2 ky + ri
    uint32_t roff[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = 24 * (r >> 1) + 2 * ky + (r & 1);
      roff[r] = k * 128 + ((((o >> 3) ^ (k & 3)) & 3) << 5) + (o & 7) * 4;
    }
    for (int k = 0; k < n_mine; ++k) {
      mbar_wait(&sm.d1_full, k & 1);
      tc::fence_after_thread_sync();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
        tc::tmem_ld32(tmem + kFzColD1 + mt * 64 + hh * 32 + lane_base, v);
        if (hh == 1) {   // D1 fully read: the next unit's GEMM1 may overwrite it
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.d1_free);
        }
#pragma unroll
        for (int tt = 0; tt < 8; ++tt) {
          const int T = k * kFzTilesPerUnit + hh * 8 + tt;
          const int sb = T % kFzNB;
          mbar_wait(&sm.bt_empty[sb], ((T / kFzNB) & 1) ^ 1);
          unsigned char* slot = sm.bt[sb];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float z = v[tt * 4 + r];
            if ((r & 1) && ky == 0) z = bias_o;   // Im of the ky = 0 column is dropped by C2R; the row carries the bias
            float hi, lo;
            tc::split_tf32(z, hi, lo);
            *reinterpret_cast<float*>(slot + roff[r]) = hi;
            *reinterpret_cast<float*>(slot + 6144 + roff[r]) = lo;
          }
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.bt_full[sb]);
        }
      }
    }
  }
  // ================================================================ epilogue
  else if (warp < kFzConvWarps + kFzEpiWarps) {
    const int q = warp & 3, half = (warp - kFzConvWarps) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int n_tiles = n_mine * kFzTilesPerUnit;
    for (int T = 0; T < n_tiles; ++T) {
      const int buf = T % kFzND;
      const int u = unit_of(T / kFzTilesPerUnit), t = T % kFzTilesPerUnit;
      mbar_wait(&sm.d2_full[buf], (T / kFzND) & 1);
      tc::fence_after_thread_sync();
      float v[16];
      fz_ld16(tmem + kFzColD2 + buf * 32 + half * 16 + lane_base, v);
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.d2_empty[buf]);
      const int b = u >> 1, px = (u & 1) * 2048 + t * 128 + q * 32 + lane;
      __nv_bfloat16* dst = out + (static_cast<size_t>(b) * kC + half * 16) * kHW + px;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const float2 g = gelu_erf2(make_float2(v[c], v[c + 1]));
        dst[static_cast<size_t>(c) * kHW] = __float2bfloat16_rn(g.x);
        dst[static_cast<size_t>(c + 1) * kHW] = __float2bfloat16_rn(g.y);
      }
    }
  }
  // ================================================================ MMA issue
  else if (warp == kFzMmaWarp) {
    if (tc::elect_one()) {
      const uint32_t f_hi = tc::smem_addr(sm.f), f_lo = f_hi + kFzFBytes / 2;
      const uint32_t w_s = tc::smem_addr(sm.w);
      constexpr uint32_t idesc_g1 = tc::make_idesc_tf32(128, 64) | kAMajorMN;
      constexpr uint32_t idesc_e = tc::make_idesc_tf32(128, 32) | kBMajorMN;
      constexpr uint32_t idesc_c = fz_idesc_bf16(128, 32) | kAMajorMN;

      auto issue_gemm1 = [&](int k) {
        const uint32_t neg = (unit_of(k) & 1) ? kANegate : 0u;   // second half image: odd kx change sign
        if (k >= 1) {
          mbar_wait(&sm.d1_free, (k - 1) & 1);
          tc::fence_after_thread_sync();
        }
#pragma unroll 1
        for (int st = 0; st < 6; ++st) {
          const int c = k * 6 + st, slot = c % kFzNY;
          const int mt = st >> 1, par = st & 1;
          mbar_wait(&sm.y_full[slot], (c / kFzNY) & 1);
          tc::fence_after_thread_sync();
          const uint32_t a_hi = tc::smem_addr(sm.y[slot]), a_lo = a_hi + kFzYStage / 2;
          const uint32_t d = tmem + kFzColD1 + mt * 64;
          const uint32_t idesc = idesc_g1 | (par ? neg : 0u);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_s = (pass == 1) ? a_lo : a_hi;
            const uint32_t b_s = (pass == 2) ? f_lo : f_hi;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
              const uint64_t da = fz_desc_sw128_32(a_s + ks * 1024, 3072, 512);
              const uint64_t db = tc::make_smem_desc(b_s + (3 * par + ks) * 2048, 1024, 128);
              fz_mma_tf32_ss(d, da, db, idesc, (par | pass | ks) ? 1u : 0u);
            }
          }
          tc::mma_commit(&sm.y_empty[slot]);
        }
        tc::mma_commit(&sm.d1_full);
      };

      if (n_mine > 0) issue_gemm1(0);
      for (int k = 0; k < n_mine; ++k) {
#pragma unroll 1
        for (int t = 0; t < kFzTilesPerUnit; ++t) {
          if (t == 8 && k + 1 < n_mine) issue_gemm1(k + 1);
          const int T = k * kFzTilesPerUnit + t;
          const int sx = T % kFzNX, sb = T % kFzNB, buf = T % kFzND;
          const uint32_t d = tmem + kFzColD2 + buf * 32;
          mbar_wait(&sm.x_full[sx], (T / kFzNX) & 1);
          mbar_wait(&sm.d2_empty[buf], ((T / kFzND) & 1) ^ 1);
          tc::fence_after_thread_sync();
          const uint32_t x_s = tc::smem_addr(sm.x[sx]);
#pragma unroll
          for (int pc = 0; pc < 3; ++pc)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              fz_mma_f16_ss(d, fz_desc_sw128(x_s + ks * 2048, 4096, 1024),
                            tc::make_smem_desc(w_s + pc * 2048 + ks * 1024, 512, 128), idesc_c, (pc | ks) ? 1u : 0u);
          tc::mma_commit(&sm.x_empty[sx]);
          mbar_wait(&sm.bt_full[sb], (T / kFzNB) & 1);
          tc::fence_after_thread_sync();
          const uint32_t z_hi = tc::smem_addr(sm.bt[sb]), z_lo = z_hi + 6144;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t a_t = tmem + kFzColE + ((pass == 1) ? 48u : 0u);
            const uint32_t b_s = (pass == 2) ? z_lo : z_hi;
#pragma unroll
            for (int ks = 0; ks < 6; ++ks)
              fz_mma_tf32_ts(d, a_t + ks * 8, fz_desc_sw128_32(b_s + ks * 1024, 0, 512), idesc_e, 1u);
          }
          tc::mma_commit(&sm.bt_empty[sb]);
          tc::mma_commit(&sm.d2_full[buf]);
        }
      }
    }
    __syncwarp();
  }
  // ================================================================ producers
  else if (warp == kFzProdWarp) {
    if (lane == 0) {          // x tiles: two {64 px, 32 ch} boxes per tile
      const int n_tiles = n_mine * kFzTilesPerUnit;
      for (int T = 0; T < n_tiles; ++T) {
        const int sx = T % kFzNX;
        const int u = unit_of(T / kFzTilesPerUnit), t = T % kFzTilesPerUnit;
        const int b = u >> 1, px0 = (u & 1) * 2048 + t * 128;
        mbar_wait(&sm.x_empty[sx], ((T / kFzNX) & 1) ^ 1);
        mbar_expect_tx(&sm.x_full[sx], kFzXBytes);
        fz_tma_load_2d(sm.x[sx], &x_map, px0, b * kC, &sm.x_full[sx]);
        fz_tma_load_2d(sm.x[sx] + 4096, &x_map, px0 + 64, b * kC, &sm.x_full[sx]);
      }
    } else if (lane == 1) {   // mode images: per (M-tile, kx parity) 8 runs of 24 rows x 128 B
      for (int k = 0; k < n_mine; ++k) {
        const unsigned char* img = ym_img + static_cast<size_t>(unit_of(k) >> 1) * kYmImgBytes;
        for (int st = 0; st < 6; ++st) {
          const int c = k * 6 + st, slot = c % kFzNY;
          const int mt = st >> 1, par = st & 1;
          mbar_wait(&sm.y_empty[slot], ((c / kFzNY) & 1) ^ 1);
          mbar_expect_tx(&sm.y_full[slot], kFzYStage);
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int g = 0; g < 4; ++g)
              bulk_g2s(sm.y[slot] + part * (kFzYStage / 2) + g * 3072,
                       img + part * kYmImgPart + mt * kYmImgMtile + g * kYmImgGroup + par * 3072, 3072, &sm.y_full[slot]);
        }
      }
    }
    __syncwarp();
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == kFzMmaWarp) tc::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Constant tables, built once per device in float64 and split into tf32 hi/lo (round to nearest).
//   etab[m][96]: row m = 64 j + w of the A operand (E (+) E): columns 0..47 hi, 48..95 lo; column k = 24 j' + 2 ky + ri:
//                c_ky/4096 * cos(2 pi ky w/64) (ri = 0), -c_ky/4096 * sin(..) (ri = 1), 0 for j != j';
//                the (ky = 0, ri = 1) column is 1 (bias row of the B operand).   c_0 = 1, c_ky = 2 (Hermitian fold).
//   ftab: B operand of GEMM1, [n = 2 h' + ri (64)][k = 24 p + 2 q + ri' (48)], kxi = 2 q + p, kx = kxi (< 12) or kxi + 40:
//         (ri,ri') = (0,0): cos t, (0,1): -sin t, (1,0): sin t, (1,1): cos t,  t = 2 pi kx h'/64;  K-major hi image | lo image.
// ------------------------------------------------------------------------------------------------
static float fz_round_tf32_host(double v) {
  float f = static_cast<float>(v);
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&f, &u, 4);
  return f;
}

struct FzTables {
  float* etab = nullptr;
  float* ftab = nullptr;
  int n_sm = 0;
  bool configured = false;
};
static FzTables g_fz[64];

static cudaError_t fz_ensure(int dev, cudaStream_t stream) {
  FzTables& t = g_fz[dev];
  if (t.configured) return cudaSuccess;
  const double two_pi = 6.283185307179586476925286766559;
  static float h_e[128 * 96];
  static float h_f[2 * 64 * 48];
  for (int m = 0; m < 128; ++m) {
    const int j = m >> 6, w = m & 63;
    for (int k = 0; k < 48; ++k) {
      const int jj = k / 24, ky = (k % 24) >> 1, ri = k & 1;
      double val = 0.0;
      if (jj == j) {
        const double c = (ky == 0 ? 1.0 : 2.0) / 4096.0, ang = two_pi * ((ky * w) % 64) / 64.0;
        val = ri == 0 ? c * cos(ang) : (ky == 0 ? 1.0 : -c * sin(ang));
      }
      const float hi = fz_round_tf32_host(val);
      h_e[m * 96 + k] = hi;
      h_e[m * 96 + 48 + k] = fz_round_tf32_host(val - static_cast<double>(hi));
    }
  }
  for (int n = 0; n < 64; ++n) {
    const int hp = n >> 1, ri = n & 1;
    for (int k = 0; k < 48; ++k) {
      const int p = k / 24, q = (k % 24) >> 1, rip = k & 1;
      const int kxi = 2 * q + p, kx = kxi < 12 ? kxi : kxi + 40;
      const double ang = two_pi * ((kx * hp) % 64) / 64.0;
      const double val = (ri == rip) ? cos(ang) : (ri == 0 ? -sin(ang) : sin(ang));
      const float hi = fz_round_tf32_host(val);
      const uint32_t off = tc::kmajor_offset(n, k, 64) / 4;
      h_f[off] = hi;
      h_f[64 * 48 + off] = fz_round_tf32_host(val - static_cast<double>(hi));
    }
  }
  cudaError_t e = cudaMalloc(&t.etab, sizeof(h_e));
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&t.ftab, sizeof(h_f));
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(t.etab, h_e, sizeof(h_e), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(t.ftab, h_f, sizeof(h_f), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);   // the host arrays are static
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(block_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FzSmem));
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&t.n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  t.configured = true;
  return cudaSuccess;
}

// tensor map of a bf16 activation [batch * 32 rows][4096 px], box {64 px, 32 rows}, 128B swizzle
typedef CUresult (*FzEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cudaError_t fz_make_map(const void* act, int batch, CUtensorMap* out) {
  static FzEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (!p) return cudaErrorNotSupported;
    fn = reinterpret_cast<FzEncodeFn>(p);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kHW), static_cast<cuuint64_t>(batch) * kC};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kHW) * 2};
  const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(kC)}, estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(act), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

size_t ym_image_bytes(int batch) { return static_cast<size_t>(batch) * kYmImgBytes; }

cudaError_t launch_block_fused(const void* ym_img, const void* x, const float* w0t, const float* bias, void* out, int batch,
                               cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  e = fz_ensure(dev, stream);
  if (e != cudaSuccess) return e;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(ym_img) & 15)) return cudaErrorMisalignedAddress;
  CUtensorMap map;
  e = fz_make_map(x, batch, &map);
  if (e != cudaSuccess) return e;
  const int n_units = 2 * batch;
  const int grid = n_units < g_fz[dev].n_sm ? n_units : g_fz[dev].n_sm;
  return launch_chained(block_fused_kernel, dim3(grid), dim3(kFzThreads), sizeof(FzSmem), stream, map,
                        static_cast<const unsigned char*>(ym_img), w0t, bias, static_cast<const float*>(g_fz[dev].etab),
                        static_cast<const float*>(g_fz[dev].ftab), static_cast<__nv_bfloat16*>(out), n_units);
}

}  // namespace fno
