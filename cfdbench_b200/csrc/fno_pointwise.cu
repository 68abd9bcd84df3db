// Lift (channel assembly + fc0) and project (fc1 + GELU + fc2 + mask) kernels.
//
// lift_kernel replaces 3x torch.cat + repeat + the per-call numpy->H2D coordinate grid + Conv2d(5+p,32,1)
// of the reference (src/models/fno/fno2d.py:195-217, 244-255) with one pass that reads u, v, mask and
// writes the 32 lifted channels: the coordinate and case-parameter channels are folded into a per-sample
// per-channel constant and two rank-1 terms,
//   a0[c] = Wu[c] u + Wv[c] v + Wm[c] mask + Wx[c] x(h) + Wy[c] y(w) + (b[c] + sum_p Wp[c][p] params[p]).
//
// project_kernel replaces Conv2d(32,128,1) + GELU + Conv2d(128,2,1) + "* mask" (fno2d.py:228-233); the
// (B,128,64,64) hidden tensor (537 MB at B=256) only ever exists as two registers per thread.
#include "fno_common.cuh"

namespace fno {

// ------------------------------------------------------------------------------------------ lift
constexpr int kLiftThreads = 256;

template <typename TAct>
__global__ void __launch_bounds__(kLiftThreads)
    lift_kernel(const float* __restrict__ inputs,  // [B][2][64][64]
                const float* __restrict__ mask,    // [B][64][64]
                const float* __restrict__ params,  // [B][p]
                const float* __restrict__ w,       // [32][5+p]   order: u, v, mask, x, y, params...
                const float* __restrict__ bias,    // [32]
                const float* __restrict__ gx, const float* __restrict__ gy,  // [64] coordinate tables
                TAct* __restrict__ out, int p) {
  __shared__ float sw[kC][5];
  __shared__ float scb[kC];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int nin = 5 + p;
  if (tid < kC) {
    float cb = bias[tid];
    for (int q = 0; q < p; ++q) cb = fmaf(w[tid * nin + 5 + q], params[b * p + q], cb);
    scb[tid] = cb;
#pragma unroll
    for (int q = 0; q < 5; ++q) sw[tid][q] = w[tid * nin + q];
  }
  __syncthreads();
  const int pix = (blockIdx.x * kLiftThreads + tid) * 4;  // 4 consecutive pixels of one row
  const int h = pix >> 6, w0 = pix & 63;
  const float4 u = *reinterpret_cast<const float4*>(inputs + (static_cast<size_t>(b) * 2 + 0) * kHW + pix);
  const float4 v = *reinterpret_cast<const float4*>(inputs + (static_cast<size_t>(b) * 2 + 1) * kHW + pix);
  const float4 m = *reinterpret_cast<const float4*>(mask + static_cast<size_t>(b) * kHW + pix);
  const float xh = gx[h];
  const float4 yw = *reinterpret_cast<const float4*>(gy + w0);
  TAct* dst = out + static_cast<size_t>(b) * kC * kHW + pix;
#pragma unroll 4
  for (int c = 0; c < kC; ++c) {
    const float wu = sw[c][0], wv = sw[c][1], wm = sw[c][2], wx = sw[c][3], wy = sw[c][4];
    const float base = fmaf(wx, xh, scb[c]);
    float4 r;
    r.x = fmaf(wu, u.x, fmaf(wv, v.x, fmaf(wm, m.x, fmaf(wy, yw.x, base))));
    r.y = fmaf(wu, u.y, fmaf(wv, v.y, fmaf(wm, m.y, fmaf(wy, yw.y, base))));
    r.z = fmaf(wu, u.z, fmaf(wv, v.z, fmaf(wm, m.z, fmaf(wy, yw.z, base))));
    r.w = fmaf(wu, u.w, fmaf(wv, v.w, fmaf(wm, m.w, fmaf(wy, yw.w, base))));
    if constexpr (sizeof(TAct) == 4) {
      *reinterpret_cast<float4*>(dst + static_cast<size_t>(c) * kHW) = r;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(dst + static_cast<size_t>(c) * kHW) = pk;
    }
  }
}

template <typename TAct>
cudaError_t launch_lift(const float* inputs, const float* mask, const float* params, const float* w,
                        const float* bias, const float* gx, const float* gy, void* out, int batch, int p,
                        cudaStream_t stream) {
  if (p < 0 || p > kMaxCaseParams) return cudaErrorInvalidValue;
  dim3 grid(kHW / (kLiftThreads * 4), batch);
  lift_kernel<TAct><<<grid, kLiftThreads, 0, stream>>>(inputs, mask, params, w, bias, gx, gy, static_cast<TAct*>(out), p);
  return cudaGetLastError();
}
template cudaError_t launch_lift<float>(const float*, const float*, const float*, const float*, const float*,
                                        const float*, const float*, void*, int, int, cudaStream_t);
template cudaError_t launch_lift<__nv_bfloat16>(const float*, const float*, const float*, const float*, const float*,
                                                const float*, const float*, void*, int, int, cudaStream_t);

// --------------------------------------------------------------------------------------- project
constexpr int kProjThreads = 128;
constexpr int kProjPix = 256;  // pixels per CTA (4 rows), 2 per thread

template <typename TAct>
struct ProjSmem {
  alignas(128) TAct xs[kC][kProjPix];   // input tile, 32 bulk copies of 4 rows
  alignas(16) float w1[kProj][kC];      // fc1 weights (FFMA2 takes them as broadcast scalar operands)
  alignas(16) float2 w2p[kProj];        // (w2[0][j], w2[1][j])
  alignas(16) float b1[kProj];
  alignas(8) uint64_t bar;
};

template <typename TAct>
__global__ void __launch_bounds__(kProjThreads)
    project_kernel(const TAct* __restrict__ a,       // [B][32][64][64]
                   const float* __restrict__ w1,     // [128][32]
                   const float* __restrict__ b1,     // [128]
                   const float* __restrict__ w2,     // [2][128]
                   const float* __restrict__ b2,     // [2]
                   const float* __restrict__ mask,   // [B][64][64]
                   float* __restrict__ preds) {      // [B][2][64][64]
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ProjSmem<TAct>& sm = *reinterpret_cast<ProjSmem<TAct>*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int pix0 = blockIdx.x * kProjPix;

  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid < kC) {
    constexpr uint32_t bytes = kProjPix * sizeof(TAct);
    if (tid == 0) mbar_expect_tx(&sm.bar, kC * bytes);
    __syncwarp();
    bulk_g2s(&sm.xs[tid][0], a + (static_cast<size_t>(b) * kC + tid) * kHW + pix0, bytes, &sm.bar);
  }
  for (int i = tid; i < kProj * kC; i += kProjThreads) (&sm.w1[0][0])[i] = w1[i];
  for (int j = tid; j < kProj; j += kProjThreads) {
    sm.w2p[j] = make_float2(w2[j], w2[kProj + j]);
    sm.b1[j] = b1[j];
  }
  __syncthreads();
  mbar_wait(&sm.bar, 0);

  float2 x[kC];
#pragma unroll
  for (int i = 0; i < kC; ++i) {
    if constexpr (sizeof(TAct) == 4) {
      x[i] = *reinterpret_cast<const float2*>(&sm.xs[i][2 * tid]);
    } else {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(&sm.xs[i][2 * tid]);
      x[i] = make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
    }
  }
  float2 o0 = make_float2(b2[0], b2[0]), o1 = make_float2(b2[1], b2[1]);
#pragma unroll 2
  for (int j = 0; j < kProj; ++j) {
    const float bj = sm.b1[j];
    float2 acc0 = make_float2(bj, bj), acc1 = make_float2(0.f, 0.f);
    const float4* wrow = reinterpret_cast<const float4*>(&sm.w1[j][0]);
#pragma unroll
    for (int i = 0; i < kC; i += 4) {  // two accumulators to halve the dependent-FMA chain
      const float4 wv = wrow[i / 4];   // broadcast LDS.128: 4 weights
      acc0 = __ffma2_rn(x[i], make_float2(wv.x, wv.x), acc0);
      acc1 = __ffma2_rn(x[i + 1], make_float2(wv.y, wv.y), acc1);
      acc0 = __ffma2_rn(x[i + 2], make_float2(wv.z, wv.z), acc0);
      acc1 = __ffma2_rn(x[i + 3], make_float2(wv.w, wv.w), acc1);
    }
    const float2 g = gelu_erf2(make_float2(acc0.x + acc1.x, acc0.y + acc1.y));
    const float2 w2v = sm.w2p[j];
    o0 = __ffma2_rn(g, make_float2(w2v.x, w2v.x), o0);
    o1 = __ffma2_rn(g, make_float2(w2v.y, w2v.y), o1);
  }
  const int pix = pix0 + 2 * tid;
  const float2 m = *reinterpret_cast<const float2*>(mask + static_cast<size_t>(b) * kHW + pix);
  *reinterpret_cast<float2*>(preds + (static_cast<size_t>(b) * 2 + 0) * kHW + pix) = make_float2(o0.x * m.x, o0.y * m.y);
  *reinterpret_cast<float2*>(preds + (static_cast<size_t>(b) * 2 + 1) * kHW + pix) = make_float2(o1.x * m.x, o1.y * m.y);
}

template <typename TAct>
cudaError_t launch_project(const void* a, const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* mask, float* preds, int batch, cudaStream_t stream) {
  auto kern = project_kernel<TAct>;
  constexpr size_t smem = sizeof(ProjSmem<TAct>);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid(kHW / kProjPix, batch);
  kern<<<grid, kProjThreads, smem, stream>>>(static_cast<const TAct*>(a), w1, b1, w2, b2, mask, preds);
  return cudaGetLastError();
}
template cudaError_t launch_project<float>(const void*, const float*, const float*, const float*, const float*,
                                           const float*, float*, int, cudaStream_t);
template cudaError_t launch_project<__nv_bfloat16>(const void*, const float*, const float*, const float*,
                                                   const float*, const float*, float*, int, cudaStream_t);

}  // namespace fno
