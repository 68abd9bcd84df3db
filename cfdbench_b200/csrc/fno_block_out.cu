// K3 -- zero-padded inverse transform + 1x1 conv + bias + exact GELU, fused:
//   out[b][o][h][w] = act( irfft2(pad(Y))[b][o][h][w] + sum_i W0[o][i] x[b][i][h][w] + bias[o] )
//
// Replaces torch.fft.irfft2 on the zero-filled spectrum, nn.Conv2d(32,32,1), the add and nn.GELU of
// the reference FnoBlock (src/models/fno/fno2d.py:81, 104-112) -- four full-tensor round trips there,
// one read of x and one write of out here.
//
// Grid (8, B): CTA (r, b) produces rows h = 8h'+r of sample b, all 32 output channels.
//  phase A  thread (ky, o): 24 kept kx bins of Y[b][.][ky][o] -> inverse DFT along kx evaluated only at
//           the CTA's 8 rows (DIF fold mod 8, codelet icfft64_in24_r<r>) -> Zs[h'][ky][o] in smem.
//  phase B  warp h', lane o: C2R along ky (codelet c2r64_in12; Im of the ky=0 bin is dropped exactly as
//           irfft2 does) leaves the 64 pixels of row h of channel o in 64 registers; the 1x1 conv then
//           accumulates into the same registers from the x row tile [32 ch][64 px] that the TMA engine
//           bulk-copied into smem (all lanes read the same address -> broadcast LDS.128, FFMA2 math),
//           then bias + GELU + store.  The zero-padded inverse never leaves registers.
// EPI selects the epilogue: forward inference, forward training (also stores the pre-activation),
// backward (multiply by GELU'(pre) of the previous block) or plain (gradient w.r.t. the lift output).
#include "fft_codelets.cuh"
#include "fno_common.cuh"

namespace fno {

constexpr int kOutThreads = 256;
constexpr int kOutRows = 8;  // rows per CTA == warps per CTA == INV_R of the codelet generator

enum : int { kEpiGelu = 0, kEpiGeluSavePre = 1, kEpiMulDgelu = 2, kEpiPlain = 3 };

template <typename TAct>
struct OutSmem {
  alignas(128) TAct xs[kOutRows][kC][kW];        // conv input rows, one tile per warp
  alignas(16) float2 zs[kOutRows][kM2][kC];      // Z[h'][ky][o]
  alignas(16) float w0t[kC][kC];                 // w0t[i][o] = W0[o][i]
  alignas(16) float bias[kC];
  alignas(8) uint64_t bar[kOutRows];
};

template <int R>
__device__ __forceinline__ void inv_kx(const float* yre, const float* yim, float* ore, float* oim) {
  if constexpr (R == 0) fno_codelets::icfft64_in24_r0<float>(yre, yim, ore, oim);
  if constexpr (R == 1) fno_codelets::icfft64_in24_r1<float>(yre, yim, ore, oim);
  if constexpr (R == 2) fno_codelets::icfft64_in24_r2<float>(yre, yim, ore, oim);
  if constexpr (R == 3) fno_codelets::icfft64_in24_r3<float>(yre, yim, ore, oim);
  if constexpr (R == 4) fno_codelets::icfft64_in24_r4<float>(yre, yim, ore, oim);
  if constexpr (R == 5) fno_codelets::icfft64_in24_r5<float>(yre, yim, ore, oim);
  if constexpr (R == 6) fno_codelets::icfft64_in24_r6<float>(yre, yim, ore, oim);
  if constexpr (R == 7) fno_codelets::icfft64_in24_r7<float>(yre, yim, ore, oim);
}

__device__ __forceinline__ void load_row_pairs(const float* p, int n, float2& a, float2& b) {
  const float4 v = *reinterpret_cast<const float4*>(p + 4 * n);
  a = make_float2(v.x, v.y);
  b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void load_row_pairs(const __nv_bfloat16* p, int n, float2& a, float2& b) {
  const uint2 v = *reinterpret_cast<const uint2*>(p + 4 * n);  // 4 bf16
  a = make_float2(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u));
  b = make_float2(__uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
}

__device__ __forceinline__ void store_row(float* dst, const float2* v) {
#pragma unroll
  for (int n = 0; n < 16; ++n)
    reinterpret_cast<float4*>(dst)[n] = make_float4(v[2 * n].x, v[2 * n].y, v[2 * n + 1].x, v[2 * n + 1].y);
}
__device__ __forceinline__ void store_row(__nv_bfloat16* dst, const float2* v) {
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    uint4 pk;
    __nv_bfloat162 t;
    t = __float22bfloat162_rn(v[4 * n + 0]); pk.x = *reinterpret_cast<uint32_t*>(&t);
    t = __float22bfloat162_rn(v[4 * n + 1]); pk.y = *reinterpret_cast<uint32_t*>(&t);
    t = __float22bfloat162_rn(v[4 * n + 2]); pk.z = *reinterpret_cast<uint32_t*>(&t);
    t = __float22bfloat162_rn(v[4 * n + 3]); pk.w = *reinterpret_cast<uint32_t*>(&t);
    reinterpret_cast<uint4*>(dst)[n] = pk;
  }
}

template <typename TAct, int EPI>
__global__ void __launch_bounds__(kOutThreads, 2)
    block_out_kernel(const float2* __restrict__ ym, const TAct* __restrict__ x, const float* __restrict__ w0t,
                     const float* __restrict__ bias, TAct* __restrict__ out, float* __restrict__ pre_out,
                     const float* __restrict__ pre_in, float s0, float s1) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  OutSmem<TAct>& sm = *reinterpret_cast<OutSmem<TAct>*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = blockIdx.x, b = blockIdx.y;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kOutRows; ++i) mbar_init(&sm.bar[i], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kC * kC; i += kOutThreads) (&sm.w0t[0][0])[i] = w0t[i];
  if (tid < kC) sm.bias[tid] = (bias != nullptr) ? bias[tid] : 0.f;
  __syncthreads();

  // ---- TMA: row h = 8*warp + r of all 32 input channels -> sm.xs[warp] (32 bulk copies of one row)
  {
    constexpr uint32_t row_bytes = kW * sizeof(TAct);
    if (lane == 0) mbar_expect_tx(&sm.bar[warp], kC * row_bytes);
    __syncwarp();
    const int h = kOutRows * warp + r;
    bulk_g2s(&sm.xs[warp][lane][0], x + ((static_cast<size_t>(b) * kC + lane) * kH + h) * kW, row_bytes,
             &sm.bar[warp]);
  }

  // ---- phase A: inverse along kx for this CTA's 8 rows ------------------------------------------
  {
    const int o = lane;
    const float2* ym_b = ym + static_cast<size_t>(b) * kModes * kC;
#pragma unroll 1
    for (int ky = warp; ky < kM2; ky += kOutRows) {
      float yre[kKX], yim[kKX], ore[8], oim[8];
#pragma unroll
      for (int kxi = 0; kxi < kKX; ++kxi) {
        const float2 v = __ldg(ym_b + (kxi * kM2 + ky) * kC + o);
        yre[kxi] = v.x;
        yim[kxi] = v.y;
      }
      switch (r) {
        case 0: inv_kx<0>(yre, yim, ore, oim); break;
        case 1: inv_kx<1>(yre, yim, ore, oim); break;
        case 2: inv_kx<2>(yre, yim, ore, oim); break;
        case 3: inv_kx<3>(yre, yim, ore, oim); break;
        case 4: inv_kx<4>(yre, yim, ore, oim); break;
        case 5: inv_kx<5>(yre, yim, ore, oim); break;
        case 6: inv_kx<6>(yre, yim, ore, oim); break;
        default: inv_kx<7>(yre, yim, ore, oim); break;
      }
      const float s = (ky == 0) ? s0 : s1;
#pragma unroll
      for (int hp = 0; hp < 8; ++hp) sm.zs[hp][ky][o] = make_float2(ore[hp] * s, oim[hp] * s);
    }
  }
  __syncthreads();

  // ---- phase B: C2R along ky into registers, 1x1 conv on top, epilogue ---------------------------
  const int h = kOutRows * warp + r;
  float2 acc[32];
  {
    float zre[kM2], zim[kM2], y[64];
#pragma unroll
    for (int k = 0; k < kM2; ++k) {
      const float2 v = sm.zs[warp][k][lane];
      zre[k] = v.x;
      zim[k] = v.y;
    }
    fno_codelets::c2r64_in12<float>(zre, zim, y);
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = make_float2(y[2 * n], y[2 * n + 1]);
  }
  mbar_wait(&sm.bar[warp], 0);
#pragma unroll 4
  for (int i = 0; i < kC; ++i) {
    const float wv = sm.w0t[i][lane];
    const float2 ww = make_float2(wv, wv);
    const TAct* xr = &sm.xs[warp][i][0];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      float2 p0, p1;
      load_row_pairs(xr, n, p0, p1);
      acc[2 * n] = __ffma2_rn(p0, ww, acc[2 * n]);
      acc[2 * n + 1] = __ffma2_rn(p1, ww, acc[2 * n + 1]);
    }
  }

  const size_t off = ((static_cast<size_t>(b) * kC + lane) * kH + h) * kW;
  if constexpr (EPI == kEpiGelu || EPI == kEpiGeluSavePre) {
    const float bv = sm.bias[lane];
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = make_float2(acc[n].x + bv, acc[n].y + bv);
    if constexpr (EPI == kEpiGeluSavePre) store_row(pre_out + off, acc);
#pragma unroll
    for (int n = 0; n < 32; ++n) acc[n] = gelu_erf2(acc[n]);
  } else if constexpr (EPI == kEpiMulDgelu) {
    const float4* pp = reinterpret_cast<const float4*>(pre_in + off);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const float4 pv = __ldg(pp + n);
      acc[2 * n].x *= dgelu_erf(pv.x);
      acc[2 * n].y *= dgelu_erf(pv.y);
      acc[2 * n + 1].x *= dgelu_erf(pv.z);
      acc[2 * n + 1].y *= dgelu_erf(pv.w);
    }
  }
  store_row(out + off, acc);
}

template <typename TAct, int EPI>
static cudaError_t launch_one(const void* ym, const void* x, const float* w0t, const float* bias, void* out,
                              float* pre_out, const float* pre_in, int batch, float s0, float s1,
                              cudaStream_t stream) {
  auto kern = block_out_kernel<TAct, EPI>;
  constexpr size_t smem = sizeof(OutSmem<TAct>);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid(kOutRows, batch);
  kern<<<grid, kOutThreads, smem, stream>>>(static_cast<const float2*>(ym), static_cast<const TAct*>(x), w0t, bias,
                                            static_cast<TAct*>(out), pre_out, pre_in, s0, s1);
  return cudaGetLastError();
}

template <typename TAct>
cudaError_t launch_block_out(int epi, const void* ym, const void* x, const float* w0t, const float* bias, void* out,
                             float* pre_out, const float* pre_in, int batch, float s0, float s1,
                             cudaStream_t stream) {
  switch (epi) {
    case kEpiGelu: return launch_one<TAct, kEpiGelu>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    case kEpiGeluSavePre:
      return launch_one<TAct, kEpiGeluSavePre>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    case kEpiMulDgelu:
      return launch_one<TAct, kEpiMulDgelu>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    case kEpiPlain: return launch_one<TAct, kEpiPlain>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    default: return cudaErrorInvalidValue;
  }
}

template cudaError_t launch_block_out<float>(int, const void*, const void*, const float*, const float*, void*, float*,
                                             const float*, int, float, float, cudaStream_t);
template cudaError_t launch_block_out<__nv_bfloat16>(int, const void*, const void*, const float*, const float*, void*,
                                                     float*, const float*, int, float, float, cudaStream_t);

}  // namespace fno
