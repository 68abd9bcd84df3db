// K2 -- per-mode complex channel mix on the tensor cores:  Y[k][b][o] = sum_i X[k][b][i] * Wk[k][i][o].
//
// Replaces the two torch.einsum("bixy,ioxy->boxy") corner products plus the zero-filled
// (B,32,64,33) cfloat buffer of the reference (src/models/fno/fno2d.py:54-57, 65-78).
//
// Modes are stored mode-major (xm[k][b][c], written that way by the forward DFT kernels), so the 128 rows of a tile are
// one contiguous 32 KB block.  For one mode k the mix over a tile of 128 samples is a real GEMM on the interleaved
// complex64 rows exactly as they sit in memory:
//     D[128 samples][64 = (o, re|im)] = A[128][64 = (i, re|im)] * B_k^T         (tcgen05.mma kind::tf32, K = 64)
//     B_k[(o,re)][(i,re)] = Wre,  B_k[(o,re)][(i,im)] = -Wim,  B_k[(o,im)][(i,re)] = Wim,  B_k[(o,im)][(i,im)] = Wre
// run as 3xTF32.  The B operand of every mode is prepared once per weight update (pack_mix_operand_*_kernel:
// real-expanded, split into tf32 hi/lo, laid out as the K-major UMMA image) and arrives with ONE 32 KB bulk copy per mode.
//
// Warp-specialised (round 2; round 1's version staged A through registers in two 256-thread pipelines and was
// latency-bound at ~2 tiles per pipeline: 19 us against the ~10 us its 66 MB need):
//   * work item = one MODE (all its sample tiles): B_k is fetched once per mode, 288 items on 148 persistent CTAs;
//   * warp 16 lane 0 -- producer: the A tile is the raw fp32 block itself, dropped by TMA (two {32 floats, 128 rows}
//     boxes, 128-byte swizzle) into a 4-slot ring as a K-major operand.  At B = 256 the ring holds ALL the tiles of a CTA,
//     so every load of the kernel is in flight from the first microsecond;
//   * the tensor core truncates what it reads to tf32, so the raw tile IS the hi operand; warps 0-7 compute the lo part
//     (x - trunc(x), exact, then rounded) as soon as a tile lands and put it into TENSOR MEMORY (thread = sample row =
//     TMEM lane; one 64-column block per ring slot), where it is the A operand of the third pass.  (A first version kept
//     one lo buffer in shared memory, refilled per tile: the refill's LDS/STS then ran concurrently with the previous
//     tile's 64 KB of global stores and took ~4,000 cycles instead of ~300, serialising the tiles -- tools/trace_mix.py.)
//   * warp 17 -- MMA issue: A_raw x B_hi, A_raw x B_lo (16 MMAs, need only the TMA data) then A_lo x B_hi (8, A in TMEM);
//     4 accumulators of 64 columns, so the epilogue of a tile never holds up the next tile's MMAs;
//   * warps 8-15 -- epilogue: mode-major ym rows (fp32 path / backward) or the per-sample operand image of
//     block_fused_kernel's GEMM1 (tf32 hi/lo split here, 256-bit stores).
// With the conj-transposed pack the same kernel is the adjoint mix of the backward pass
// (Xbar[b,i,k] = sum_o G[b,o,k] conj(W[i,o,k])).
#include "fno_common.cuh"
#include "tc_common.cuh"
#include "tc_tma.cuh"

namespace fno {

constexpr int kMxM = 128;        // samples per tile
constexpr int kMxK = 2 * kC;     // 64 real (i, re|im)
constexpr int kMxN = 2 * kC;     // 64 real (o, re|im)
constexpr uint32_t kMxLboB = (kMxN / 8) * 128;       // 1024
constexpr int kMxBFloats = kMxN * kMxK;              // 4096 per image (hi or lo)
constexpr int kMxOperandFloats = 2 * kMxBFloats;     // per mode: hi image then lo image (32 KB)
constexpr int kMxConvWarps = 8, kMxEpiWarps = 8;
constexpr int kMxProdWarp = 16, kMxMmaWarp = 17;
constexpr int kMxThreads = 18 * 32;
constexpr int kMxRing = 4;                           // A slots (and accumulators)
constexpr uint32_t kMxABytes = kMxM * kMxK * 4;      // 32,768 B per tile: two K halves of 128 rows x 128 B
constexpr uint32_t kMxBBytes = kMxOperandFloats * 4; // 32,768 B per mode

// Optional timeline trace (-DFNO_FZ_TRACE build, tools/trace_mix.py): CTA 0 stamps clock64() at the hand-off points,
// trace[(role * 16 + tile) * 8 + event]; per-CTA clock / globaltimer stamps follow at 4 * 16 * 8.
#ifdef FNO_FZ_TRACE
__device__ long long* g_mx_trace = nullptr;
#define MX_T(role, T, ev)                                                                          \
  do {                                                                                             \
    if (mx_tr != nullptr && blockIdx.x == 0 && (T) < 16) mx_tr[((role) * 16 + (T)) * 8 + (ev)] = clock64(); \
  } while (0)
#define MX_CTA(ev)                                                                                 \
  do {                                                                                             \
    if (mx_tr != nullptr && threadIdx.x == 0) {                                                    \
      mx_tr[4 * 16 * 8 + blockIdx.x * 4 + (ev)] = clock64();                                       \
      long long gt_;                                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                      \
      mx_tr[4 * 16 * 8 + 148 * 4 + blockIdx.x * 4 + (ev)] = gt_;                                   \
    }                                                                                              \
  } while (0)
#else
#define MX_T(role, T, ev) do { } while (0)
#define MX_CTA(ev) do { } while (0)
#endif

struct MxSmem {
  alignas(1024) unsigned char a[kMxRing][kMxABytes];   // raw fp32 tiles (TMA, 128B swizzle): the hi operand
  alignas(1024) float b[2][kMxOperandFloats];          // [mode parity] hi | lo images
  alignas(8) uint64_t a_full[kMxRing], a_free[kMxRing], d_full[kMxRing], d_free[kMxRing], lo_ready[kMxRing];
  uint64_t b_full[2], b_free[2];
  uint32_t tmem_base;
};
constexpr uint32_t kMxColD = 0, kMxColLo = kMxRing * kMxN;   // tensor memory: 4 accumulators, then 4 lo operands
constexpr int kMxTmemCols = 2 * kMxRing * kMxN;              // 512

// 256-bit global store (sm_100: st.global.v8): the image epilogue writes 32-byte chunks of 32 different samples per warp
// instruction, so the instruction count -- not the bytes -- is what the LSU queue sees (lg_throttle 5.0 per issue with
// 16-byte stores, profiles/ncu_r02*.md).
__device__ __forceinline__ void mx_store32(void* dst, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

__global__ void __launch_bounds__(kMxThreads, 1)
    mode_mix_tc_kernel(const __grid_constant__ CUtensorMap x_map, const float* __restrict__ wop, float4* __restrict__ ym,
                       unsigned char* __restrict__ ym_img, int batch, int n_btiles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  MxSmem& sm = *reinterpret_cast<MxSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();
#ifdef FNO_FZ_TRACE
  long long* const mx_tr = g_mx_trace;
#endif
  MX_CTA(0);

  // modes of this CTA: k = first + j * stride; tiles are numbered t = j * n_btiles + bt in processing order
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_modes = (first < kModes) ? (kModes - first + stride - 1) / stride : 0;
  const int n_tiles = n_modes * n_btiles;

  if (tid == 0) {
    for (int i = 0; i < kMxRing; ++i) {
      mbar_init(&sm.a_full[i], 1);
      mbar_init(&sm.a_free[i], 1);
      mbar_init(&sm.d_full[i], 1);
      mbar_init(&sm.d_free[i], kMxEpiWarps);
      mbar_init(&sm.lo_ready[i], kMxConvWarps);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.b_full[i], 1); mbar_init(&sm.b_free[i], 1); }
    fence_mbar_init();
    // the weights of the first two modes do not depend on the previous kernel of the chain: fetch them now
    for (int j = 0; j < 2 && j < n_modes; ++j) {
      mbar_expect_tx(&sm.b_full[j], kMxBBytes);
      bulk_g2s(sm.b[j], wop + static_cast<size_t>(first + j * stride) * kMxOperandFloats, kMxBBytes, &sm.b_full[j]);
    }
  }
  if (warp == kMxMmaWarp) tc::tmem_alloc<kMxTmemCols>(&sm.tmem_base);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = sm.tmem_base;
  MX_CTA(1);
  pdl_wait();  // xm comes from the previous kernel of the chain
  pdl_launch_dependents();

  // ================================================================ lo-part converters
  if (warp < kMxConvWarps) {
    // thread = sample row m = 32 quad + lane (its TMEM lane) and one K half (32 floats = one swizzled 128-byte segment)
    const int quad = warp & 3, kh = warp >> 2, m = quad * 32 + lane;
    const uint32_t t_dst0 = tmem + kMxColLo + kh * 32 + (static_cast<uint32_t>(quad * 32) << 16);
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kMxRing;
      if (tid == 0) MX_T(0, t, 0);
      mbar_wait(&sm.a_full[s], (t / kMxRing) & 1);
      if (tid == 0) MX_T(0, t, 1);
      // the lo block of this slot was last read by the MMAs of tile t - kMxRing, whose completion released the slot to the
      // producer (a_free) before this tile could land: no separate barrier
      if (tid == 0) MX_T(0, t, 2);
      const unsigned char* row = sm.a[s] + kh * (kMxABytes / 2) + m * 128;
      float lo[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {   // logical 16-byte chunk c of the row sits at position c ^ (m & 7)
        const float4 x = *reinterpret_cast<const float4*>(row + ((c ^ (m & 7)) << 4));
        // x - trunc_tf32(x) is exact; +0x1000 rounds what the tensor core then truncates
        lo[4 * c + 0] = __uint_as_float(__float_as_uint(x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u)) + 0x1000u);
        lo[4 * c + 1] = __uint_as_float(__float_as_uint(x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u)) + 0x1000u);
        lo[4 * c + 2] = __uint_as_float(__float_as_uint(x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u)) + 0x1000u);
        lo[4 * c + 3] = __uint_as_float(__float_as_uint(x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u)) + 0x1000u);
      }
      tc::tmem_st16(t_dst0 + s * kMxN, lo);
      tc::tmem_st16(t_dst0 + s * kMxN + 16, lo + 16);
      tc::tmem_wait_st();
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.lo_ready[s]);
      if (tid == 0) MX_T(0, t, 3);
    }
  }
  // ================================================================ epilogue
  // warps w and w+4 share TMEM lane quadrant w & 3 (rows 32 (w&3) .. +31 of the tile) and take the 32-float column halves
  else if (warp < kMxConvWarps + kMxEpiWarps) {
    const int quad = warp & 3, half = (warp >> 2) & 1;
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % kMxRing;
      const int k = first + (t / n_btiles) * stride, b = (t % n_btiles) * kMxM + quad * 32 + lane;
      if (warp == kMxConvWarps && lane == 0) MX_T(1, t, 0);
      mbar_wait(&sm.d_full[s], (t / kMxRing) & 1);
      tc::fence_after_thread_sync();
      if (warp == kMxConvWarps && lane == 0) MX_T(1, t, 1);
      float v[32];
      tc::tmem_ld32(tmem + kMxColD + (static_cast<uint32_t>(quad * 32) << 16) + s * kMxN + half * 32, v);
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.d_free[s]);
      if (b < batch && ym_img == nullptr) {
        float4* dst = ym + (static_cast<size_t>(k) * batch + b) * (kC / 2) + half * 8;
#pragma unroll
        for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      } else if (b < batch) {
        // Operand image of block_fused_kernel's GEMM1 (fno_block_fused.cu): per sample [hi|lo][ky/4][ky%4][48 rows][32 o]
        // fp32, row = 24 (kxi & 1) + 2 (kxi >> 1) + (re|im), 32-byte chunks XOR-swizzled with (row & 3); tf32 hi / lo
        // split here so the consumer is pure bulk copy + MMA.  This thread holds o = 16 half .. 16 half + 15, (re, im).
        const int kxi = k / kM2, ky = k % kM2;
        unsigned char* img = ym_img + static_cast<size_t>(b) * 147456 + (ky >> 2) * 24576 + (ky & 3) * 6144;
#pragma unroll
        for (int ri = 0; ri < 2; ++ri) {
          const int row = 24 * (kxi & 1) + 2 * (kxi >> 1) + ri;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) tc::split_tf32(v[2 * (8 * c + j) + ri], hi[j], lo[j]);
            unsigned char* dst = img + row * 128 + (((2 * half + c) ^ (row & 3)) << 5);
            mx_store32(dst, hi);            // one 32-byte chunk = one full sector per store instruction
            mx_store32(dst + 73728, lo);
          }
        }
      }
      if (warp == kMxConvWarps && lane == 0) MX_T(1, t, 2);
    }
  }
  // ================================================================ producer
  else if (warp == kMxProdWarp) {
    if (lane == 0) {
      int t = 0;
      for (int j = 0; j < n_modes; ++j) {
        const int k = first + j * stride;
        if (j >= 2) {   // modes 0 and 1 were requested in the prologue
          mbar_wait(&sm.b_free[j & 1], ((j >> 1) - 1) & 1);
          mbar_expect_tx(&sm.b_full[j & 1], kMxBBytes);
          bulk_g2s(sm.b[j & 1], wop + static_cast<size_t>(k) * kMxOperandFloats, kMxBBytes, &sm.b_full[j & 1]);
        }
        for (int bt = 0; bt < n_btiles; ++bt, ++t) {
          const int s = t % kMxRing;
          if (t >= kMxRing) mbar_wait(&sm.a_free[s], ((t / kMxRing) - 1) & 1);
          MX_T(2, t, 0);
          mbar_expect_tx(&sm.a_full[s], kMxABytes);
          const int row0 = k * batch + bt * kMxM;   // rows past this mode's samples are never stored by the epilogue
          fz_tma_load_2d(sm.a[s], &x_map, 0, row0, &sm.a_full[s]);
          fz_tma_load_2d(sm.a[s] + kMxABytes / 2, &x_map, 32, row0, &sm.a_full[s]);
        }
      }
    }
    __syncwarp();
  }
  // ================================================================ MMA issue
  else if (warp == kMxMmaWarp) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc_tf32(kMxM, kMxN);
      int t = 0;
#pragma unroll 1
      for (int j = 0; j < n_modes; ++j) {
        mbar_wait(&sm.b_full[j & 1], (j >> 1) & 1);
        const uint32_t b_hi = tc::smem_addr(sm.b[j & 1]), b_lo = b_hi + kMxBFloats * 4;
#pragma unroll 1
        for (int bt = 0; bt < n_btiles; ++bt, ++t) {
          const int s = t % kMxRing;
          MX_T(3, t, 0);
          mbar_wait(&sm.a_full[s], (t / kMxRing) & 1);
          if (t >= kMxRing) mbar_wait(&sm.d_free[s], ((t / kMxRing) - 1) & 1);
          tc::fence_after_thread_sync();
          MX_T(3, t, 1);
          const uint32_t d = tmem + kMxColD + s * kMxN, a_s = tc::smem_addr(sm.a[s]);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {   // A_raw x B_hi, A_raw x B_lo
            const uint32_t pb = pass ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < kMxK / 8; ++ks)   // K = 8 per MMA: 32 bytes inside the 128-byte swizzle row of a K half
              fz_mma_tf32_ss(d, fz_desc_sw128(a_s + (ks >> 2) * (kMxABytes / 2) + (ks & 3) * 32, 0, 1024),
                             tc::make_smem_desc(pb + ks * 2 * kMxLboB, kMxLboB, 128), idesc, (pass | ks) ? 1u : 0u);
          }
          MX_T(3, t, 2);
          mbar_wait(&sm.lo_ready[s], (t / kMxRing) & 1);
          tc::fence_after_thread_sync();
          MX_T(3, t, 3);
#pragma unroll
          for (int ks = 0; ks < kMxK / 8; ++ks)   // A_lo (tensor memory) x B_hi
            fz_mma_tf32_ts(d, tmem + kMxColLo + s * kMxN + ks * 8, tc::make_smem_desc(b_hi + ks * 2 * kMxLboB, kMxLboB, 128), idesc, 1u);
          tc::mma_commit(&sm.a_free[s]);
          tc::mma_commit(&sm.d_full[s]);
          if (bt == n_btiles - 1) tc::mma_commit(&sm.b_free[j & 1]);
          MX_T(3, t, 4);
        }
      }
    }
    __syncwarp();
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  MX_CTA(2);
  if (warp == kMxMmaWarp) tc::tmem_dealloc<kMxTmemCols>(tmem);
}

#ifdef FNO_FZ_TRACE
extern "C" int fno_debug_mix_trace(void* p) {
  long long* q = static_cast<long long*>(p);
  return cudaMemcpyToSymbol(g_mx_trace, &q, sizeof(q)) == cudaSuccess ? 0 : 2;
}
#endif

// tensor map of the mode-major spectrum as rows of 64 floats: [288 * batch rows][64], box {32 floats, 128 rows}
static cudaError_t mx_make_map(const void* xm, int batch, CUtensorMap* out) {
  static FzEncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) return e;
    if (!p) return cudaErrorNotSupported;
    fn = reinterpret_cast<FzEncodeFn>(p);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(kMxK), static_cast<cuuint64_t>(kModes) * batch};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(kMxK) * 4};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(kMxM)}, estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(xm), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// ym_img != nullptr: write the per-sample GEMM1 operand image (see the epilogue) instead of the mode-major ym.
cudaError_t launch_mode_mix(const void* xm, const void* wop, void* ym, void* ym_img, int batch, cudaStream_t stream) {
  auto kern = mode_mix_tc_kernel;
  constexpr size_t smem = sizeof(MxSmem);
  static PerDeviceLaunch pd;
  int n_sm = 0;
  cudaError_t e0 = per_device_setup(kern, smem, pd, &n_sm);
  if (e0 != cudaSuccess) return e0;
  if (reinterpret_cast<uintptr_t>(xm) & 15) return cudaErrorMisalignedAddress;
  CUtensorMap map;
  e0 = mx_make_map(xm, batch, &map);
  if (e0 != cudaSuccess) return e0;
  const int n_btiles = (batch + kMxM - 1) / kMxM;
  const int grid = kModes < n_sm ? kModes : n_sm;
  return launch_chained(kern, dim3(grid), dim3(kMxThreads), smem, stream, map, static_cast<const float*>(wop),
                        static_cast<float4*>(ym), static_cast<unsigned char*>(ym_img), batch, n_btiles);
}

// ------------------------------------------------------------------------------------------------
// B operand image of the mix, built from the packed weights Wk[k][i][o] (pack_spectral_kernel below):
// per mode 2 x 4096 floats (tf32 hi image, then lo image), element (n = 2o + part, kk = 2i + ri) at
// tc::kmajor_offset(n, kk, 64).
// ------------------------------------------------------------------------------------------------
__global__ void pack_mix_operand_kernel(const float2* __restrict__ wk, float* __restrict__ wop) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over k*1024 + i*32 + o
  if (idx >= kModes * kC * kC) return;
  const int k = idx / (kC * kC), i = (idx / kC) % kC, o = idx % kC;
  const float2 w = wk[idx];
  float* img = wop + static_cast<size_t>(k) * kMxOperandFloats;
  const float val[2][2] = {{w.x, -w.y}, {w.y, w.x}};  // [part of the output][re|im of the input]
#pragma unroll
  for (int part = 0; part < 2; ++part)
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
      float hi, lo;
      tc::split_tf32(val[part][ri], hi, lo);
      const uint32_t off = tc::kmajor_offset(2 * o + part, 2 * i + ri, kMxN) / 4;
      img[off] = hi;
      img[kMxBFloats + off] = lo;
    }
}

cudaError_t launch_pack_mix_operand(const void* wk, void* wop, cudaStream_t stream) {
  const int n = kModes * kC * kC;
  pack_mix_operand_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const float2*>(wk), static_cast<float*>(wop));
  return cudaGetLastError();
}

size_t mix_operand_bytes() { return static_cast<size_t>(kModes) * kMxOperandFloats * sizeof(float); }

// The same image straight from the reference parameter layout (weights1/2: (Cin, Cout, 12, 12) complex64), one
// launch per layer and direction: a thread produces one 16-byte operand chunk (row n, 4 consecutive kk) of both the
// hi and the lo image, so the writes are coalesced; the 2.36 MB of weights are gathered through L2.  This is what
// runs after every optimizer step.
__global__ void pack_mix_operand_direct_kernel(const float2* __restrict__ w1, const float2* __restrict__ w2,
                                               float* __restrict__ wop, int conj_transpose) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over k * 1024 + (kk/4) * 64 + n: the operand's own order
  if (idx >= kModes * 1024) return;
  const int k = idx >> 10, kq = (idx >> 6) & 15, n = idx & 63;
  const int kxi = k / kM2, ky = k % kM2;
  const float2* src = (kxi < kM1) ? w1 : w2;
  const int kk_mode = (kxi % kM1) * kM2 + ky;
  const int a = n >> 1, part = n & 1;  // output channel (o) of the mix and its re|im
  float v[4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = 2 * kq + j;  // input channel (i) of the mix
    // forward: Wk[k][i = c][o = a] = W[c][a][k];  adjoint: Wk[k][i = c][o = a] = conj(W[a][c][k])
    const int wi = conj_transpose ? a : c, wo = conj_transpose ? c : a;
    float2 w = __ldg(src + (static_cast<size_t>(wi) * kC + wo) * (kM1 * kM2) + kk_mode);
    if (conj_transpose) w.y = -w.y;
    v[2 * j + 0] = part ? w.y : w.x;    // (o,re): [Wre, -Wim]   (o,im): [Wim, Wre]
    v[2 * j + 1] = part ? w.x : -w.y;
  }
  float4 hi, lo;
  tc::split_tf32(v[0], hi.x, lo.x);
  tc::split_tf32(v[1], hi.y, lo.y);
  tc::split_tf32(v[2], hi.z, lo.z);
  tc::split_tf32(v[3], hi.w, lo.w);
  float* img = wop + static_cast<size_t>(k) * kMxOperandFloats;
  const uint32_t off = tc::kmajor_offset(n, 4 * kq, kMxN) / 4;
  *reinterpret_cast<float4*>(img + off) = hi;
  *reinterpret_cast<float4*>(img + kMxBFloats + off) = lo;
}

cudaError_t launch_pack_mix_operand_direct(const void* w1, const void* w2, void* wop, int conj_transpose,
                                           cudaStream_t stream) {
  const int n = kModes * 1024;
  pack_mix_operand_direct_kernel<<<(n + 255) / 256, 256, 0, stream>>>(
      static_cast<const float2*>(w1), static_cast<const float2*>(w2), static_cast<float*>(wop), conj_transpose);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Weight packing: reference parameter layout (Cin, Cout, 12, 12) complex64 x2 (weights1, weights2;
// reference fno2d.py:31-51)  ->  Wk[k][i][o], k = kxi*12 + ky, kxi<12 from weights1 else weights2.
// conj_transpose=1 writes Wk[k][o][i] = conj(W[i][o][k]) (the operand of the adjoint mix:
// Xbar[b,i,k] = sum_o G[b,o,k] conj(W[i,o,k])).
// ------------------------------------------------------------------------------------------------
__global__ void pack_spectral_kernel(const float2* __restrict__ w1, const float2* __restrict__ w2,
                                     float2* __restrict__ wk, int conj_transpose) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over k*1024 + a*32 + c
  if (idx >= kModes * kC * kC) return;
  const int k = idx / (kC * kC);
  const int a = (idx / kC) % kC, c = idx % kC;
  const int i = conj_transpose ? c : a;
  const int o = conj_transpose ? a : c;
  const int kxi = k / kM2, ky = k % kM2;
  const float2* src = (kxi < kM1) ? w1 : w2;
  const int kk = (kxi % kM1) * kM2 + ky;
  float2 v = src[(static_cast<size_t>(i) * kC + o) * (kM1 * kM2) + kk];
  if (conj_transpose) v.y = -v.y;
  wk[idx] = v;
}

cudaError_t launch_pack_spectral(const void* w1, const void* w2, void* wk, int conj_transpose, cudaStream_t stream) {
  const int n = kModes * kC * kC;
  pack_spectral_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const float2*>(w1), static_cast<const float2*>(w2),
                                                           static_cast<float2*>(wk), conj_transpose);
  return cudaGetLastError();
}

// inverse of the pack for gradients: gWk[k][i][o] -> gw1/gw2 (Cin, Cout, 12, 12)
__global__ void unpack_spectral_kernel(const float2* __restrict__ gwk, float2* __restrict__ gw1,
                                       float2* __restrict__ gw2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over (i*32+o)*144 + kk for both halves
  if (idx >= 2 * kC * kC * kM1 * kM2) return;
  const int half = idx / (kC * kC * kM1 * kM2);
  const int r = idx % (kC * kC * kM1 * kM2);
  const int io = r / (kM1 * kM2), kk = r % (kM1 * kM2);
  const int k = (half * kM1 + kk / kM2) * kM2 + kk % kM2;
  const float2 v = gwk[static_cast<size_t>(k) * kC * kC + io];
  (half ? gw2 : gw1)[r] = v;
}

cudaError_t launch_unpack_spectral(const void* gwk, void* gw1, void* gw2, cudaStream_t stream) {
  const int n = 2 * kC * kC * kM1 * kM2;
  unpack_spectral_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const float2*>(gwk), static_cast<float2*>(gw1),
                                                             static_cast<float2*>(gw2));
  return cudaGetLastError();
}

}  // namespace fno
