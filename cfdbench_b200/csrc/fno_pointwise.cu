// Lift kernel (channel assembly + fc0).  The project stage lives in fno_project_tc.cu.
//
// lift_kernel replaces 3x torch.cat + repeat + the per-call numpy->H2D coordinate grid + Conv2d(5+p,32,1)
// of the reference (src/models/fno/fno2d.py:195-217, 244-255) with one pass that reads u, v, mask and
// writes the 32 lifted channels: the coordinate and case-parameter channels are folded into a per-sample
// per-channel constant and two rank-1 terms,
//   a0[c] = Wu[c] u + Wv[c] v + Wm[c] mask + Wx[c] x(h) + Wy[c] y(w) + (b[c] + sum_p Wp[c][p] params[p]).
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
  pdl_wait();  // first kernel of a step: inputs are the previous step's predictions
  pdl_launch_dependents();
  if (tid < kC) {
    float cb = bias[tid];
    for (int q = 0; q < p; ++q) cb = fmaf(w[tid * nin + 5 + q], params[b * p + q], cb);
    scb[tid] = cb;
#pragma unroll
    for (int q = 0; q < 5; ++q) sw[tid][q] = w[tid * nin + q];
  }
  __syncthreads();
  // 16-byte stores for both storage types: 4 fp32 or 8 bf16 consecutive pixels of one row per thread and channel
  constexpr int kPx = 16 / sizeof(TAct);
  const int pix = (blockIdx.x * kLiftThreads + tid) * kPx;
  const int h = pix >> 6, w0 = pix & 63;
  float u[kPx], v[kPx], m[kPx], yw[kPx];
#pragma unroll
  for (int j = 0; j < kPx; j += 4) {
    *reinterpret_cast<float4*>(u + j) = *reinterpret_cast<const float4*>(inputs + (static_cast<size_t>(b) * 2 + 0) * kHW + pix + j);
    *reinterpret_cast<float4*>(v + j) = *reinterpret_cast<const float4*>(inputs + (static_cast<size_t>(b) * 2 + 1) * kHW + pix + j);
    *reinterpret_cast<float4*>(m + j) = *reinterpret_cast<const float4*>(mask + static_cast<size_t>(b) * kHW + pix + j);
    *reinterpret_cast<float4*>(yw + j) = *reinterpret_cast<const float4*>(gy + w0 + j);
  }
  const float xh = gx[h];
  TAct* dst = out + static_cast<size_t>(b) * kC * kHW + pix;
#pragma unroll 4
  for (int c = 0; c < kC; ++c) {
    const float wu = sw[c][0], wv = sw[c][1], wm = sw[c][2], wx = sw[c][3], wy = sw[c][4];
    const float base = fmaf(wx, xh, scb[c]);
    float r[kPx];
#pragma unroll
    for (int j = 0; j < kPx; ++j) r[j] = fmaf(wu, u[j], fmaf(wv, v[j], fmaf(wm, m[j], fmaf(wy, yw[j], base))));
    if constexpr (sizeof(TAct) == 4) {
      *reinterpret_cast<float4*>(dst + static_cast<size_t>(c) * kHW) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
      uint4 pk;
      __nv_bfloat162 t;
      t = __floats2bfloat162_rn(r[0], r[1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(r[2], r[3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(r[4], r[5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(r[6], r[7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(c) * kHW) = pk;
    }
  }
}

template <typename TAct>
cudaError_t launch_lift(const float* inputs, const float* mask, const float* params, const float* w,
                        const float* bias, const float* gx, const float* gy, void* out, int batch, int p,
                        cudaStream_t stream) {
  if (p < 0 || p > kMaxCaseParams) return cudaErrorInvalidValue;
  dim3 grid(kHW / (kLiftThreads * (16 / static_cast<int>(sizeof(TAct)))), batch);
  return launch_chained(lift_kernel<TAct>, grid, dim3(kLiftThreads), 0, stream, inputs, mask, params, w, bias, gx, gy,
                        static_cast<TAct*>(out), p);
}
template cudaError_t launch_lift<float>(const float*, const float*, const float*, const float*, const float*,
                                        const float*, const float*, void*, int, int, cudaStream_t);
template cudaError_t launch_lift<__nv_bfloat16>(const float*, const float*, const float*, const float*, const float*,
                                                const float*, const float*, void*, int, int, cudaStream_t);

}  // namespace fno
