// SURVEY.md 8f.1 -- rollout evaluation metrics on the device.
// The reference computes, per rollout step and per case, mse / nmse / mae of the masked u channel with three
// .item() host syncs each (src/test_multistep.py:73-83, 153-177).  Here one launch reduces every (step, case) plane
// to the three sums the metrics are made of; a single D2H of steps*B*3 floats replaces 3*steps*B synchronisations.
#include "fno_common.cuh"

namespace fno {

constexpr int kMtThreads = 256;

// preds_seq [S][B][2][64][64] (channel 0 = u), label_u [S][B][64][64], mask [S][B][64][64]
// out [S][B][3] = (sum (p-l)^2, sum l^2, sum |p-l|) over the 4096 pixels, with p, l multiplied by the mask
__global__ void __launch_bounds__(kMtThreads)
    multistep_metrics_kernel(const float* __restrict__ preds_seq, const float* __restrict__ label_u,
                             const float* __restrict__ mask, float* __restrict__ out, int batch) {
  __shared__ float red[kMtThreads / 32][3];
  const int b = blockIdx.x, s = blockIdx.y;
  const size_t plane = static_cast<size_t>(s) * batch + b;
  const float4* p = reinterpret_cast<const float4*>(preds_seq + plane * 2 * kHW);
  const float4* l = reinterpret_cast<const float4*>(label_u + plane * kHW);
  const float4* m = reinterpret_cast<const float4*>(mask + plane * kHW);
  float se = 0.f, sl = 0.f, sa = 0.f;
  for (int i = threadIdx.x; i < kHW / 4; i += kMtThreads) {
    const float4 pv = __ldg(p + i), lv = __ldg(l + i), mv = __ldg(m + i);
    const float pp[4] = {pv.x * mv.x, pv.y * mv.y, pv.z * mv.z, pv.w * mv.w};
    const float ll[4] = {lv.x * mv.x, lv.y * mv.y, lv.z * mv.z, lv.w * mv.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float d = pp[c] - ll[c];
      se = fmaf(d, d, se);
      sl = fmaf(ll[c], ll[c], sl);
      sa += fabsf(d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = se;
    red[warp][1] = sl;
    red[warp][2] = sa;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kMtThreads / 32; ++w) t += red[w][threadIdx.x];  // fixed order: deterministic
    out[plane * 3 + threadIdx.x] = t;
  }
}

cudaError_t launch_multistep_metrics(const float* preds_seq, const float* label_u, const float* mask, float* out,
                                     int steps, int batch, cudaStream_t stream) {
  dim3 grid(batch, steps);
  multistep_metrics_kernel<<<grid, kMtThreads, 0, stream>>>(preds_seq, label_u, mask, out, batch);
  return cudaGetLastError();
}

}  // namespace fno
