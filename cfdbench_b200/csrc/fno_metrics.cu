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

// ------------------------------------------------------------------------------------------------
// SURVEY.md 8f.2 -- device-resident input pipeline.  The reference keeps every (input, label) frame pair of the
// training set in host tensors (src/dataset/cavity.py:326-331), and per batch runs DataLoader indexing, torch.stack,
// three channel slices, a Python loop over the case-parameter dicts and four .cuda() copies
// (collate_fn, src/train_auto.py:33-58).  With the frames resident in HBM one launch gathers a batch:
//   inputs[b] = frames_in[idx[b]][0:2], mask[b] = frames_in[idx[b]][2], label[b] = frames_out[idx[b]][0:2],
//   case_params[b] = case_table[case_ids[idx[b]]]
// ------------------------------------------------------------------------------------------------
namespace fno {

template <typename TFrame>
__device__ __forceinline__ float4 gather_ld4(const TFrame* p);
template <>
__device__ __forceinline__ float4 gather_ld4<float>(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
template <>
__device__ __forceinline__ float4 gather_ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
  return make_float4(a.x, a.y, b.x, b.y);
}

constexpr int kGbThreads = 256;

// grid (5 planes, n_idx): plane 0,1 -> inputs u,v; 2 -> mask; 3,4 -> label u,v
template <typename TFrame>
__global__ void __launch_bounds__(kGbThreads)
    gather_batch_kernel(const TFrame* __restrict__ frames_in, const TFrame* __restrict__ frames_out,
                        const float* __restrict__ case_table, const int* __restrict__ case_ids,
                        const long long* __restrict__ idx, int n_case_params, float* __restrict__ inputs,
                        float* __restrict__ label, float* __restrict__ mask, float* __restrict__ case_params) {
  const int plane = blockIdx.x, b = blockIdx.y;
  const long long i = idx[b];
  const TFrame* src = (plane < 3 ? frames_in : frames_out) + (static_cast<size_t>(i) * 3 + (plane < 3 ? plane : plane - 3)) * kHW;
  float* dst = plane < 2   ? inputs + (static_cast<size_t>(b) * 2 + plane) * kHW
               : plane == 2 ? mask + static_cast<size_t>(b) * kHW
                            : label + (static_cast<size_t>(b) * 2 + (plane - 3)) * kHW;
  for (int e = threadIdx.x * 4; e < kHW; e += kGbThreads * 4)
    *reinterpret_cast<float4*>(dst + e) = gather_ld4<TFrame>(src + e);
  if (plane == 0 && threadIdx.x < n_case_params)
    case_params[static_cast<size_t>(b) * n_case_params + threadIdx.x] =
        __ldg(case_table + static_cast<size_t>(case_ids[i]) * n_case_params + threadIdx.x);
}

cudaError_t launch_gather_batch(const void* frames_in, const void* frames_out, const float* case_table, const int* case_ids,
                                const long long* idx, int n_idx, int n_case_params, int frame_bf16, float* inputs,
                                float* label, float* mask, float* case_params, cudaStream_t stream) {
  dim3 grid(5, n_idx);
  if (frame_bf16)
    gather_batch_kernel<__nv_bfloat16><<<grid, kGbThreads, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(frames_in), static_cast<const __nv_bfloat16*>(frames_out), case_table, case_ids, idx,
        n_case_params, inputs, label, mask, case_params);
  else
    gather_batch_kernel<float><<<grid, kGbThreads, 0, stream>>>(static_cast<const float*>(frames_in),
                                                                static_cast<const float*>(frames_out), case_table, case_ids,
                                                                idx, n_case_params, inputs, label, mask, case_params);
  return cudaGetLastError();
}

}  // namespace fno
