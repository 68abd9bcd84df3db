// The two non-network pieces of a training step (reference src/train_auto.py:233-260), each one launch:
//
//  * MseLoss (reference src/models/loss.py:22-37): mse = mean((p-l)^2), rmse = sqrt(mse), mae = mean|p-l|,
//    nmse = mse / mean(l^2) over the whole batch tensor.  The reference issues 5 reduction kernels; here one kernel
//    accumulates the three sums (fixed grid, fixed-order two-stage reduction: deterministic) and the last block to
//    finish turns them into the five scalars.  The backward kernel writes dL/dpreds for any combination of upstream
//    gradients of the four dict entries (the script calls loss["nmse"].backward()).
//  * torch.optim.Adam.step (train_auto.py:213,256; complex parameters as pairs of reals, no amsgrad): all parameter
//    tensors of the model in ONE launch through a pointer table, against ~10 multi-tensor launches of the reference.
#include "fno_common.cuh"
#include "../../include/cfdbench_b200.h"

namespace fno {

// ------------------------------------------------------------------------------------------------ loss
constexpr int kLossThreads = 256;
constexpr int kLossBlocks = 296;  // 2 per SM; also the size of the partial-sum table

// out[0..4] = mse, rmse, mae, nmse, mean(l^2);  scratch: [kLossBlocks][3] floats + one uint32 ticket
__global__ void __launch_bounds__(kLossThreads)
    loss_fwd_kernel(const float* __restrict__ preds, const float* __restrict__ labels, size_t n, float* __restrict__ scratch,
                    float* __restrict__ out) {
  __shared__ float red[kLossThreads / 32][3];
  __shared__ bool last;
  float se = 0.f, sa = 0.f, sl = 0.f;
  const size_t n4 = n / 4;
  const float4* p4 = reinterpret_cast<const float4*>(preds);
  const float4* l4 = reinterpret_cast<const float4*>(labels);
  for (size_t i = static_cast<size_t>(blockIdx.x) * kLossThreads + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * kLossThreads) {
    const float4 p = __ldg(p4 + i), l = __ldg(l4 + i);
    const float d[4] = {p.x - l.x, p.y - l.y, p.z - l.z, p.w - l.w};
    const float lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      se = fmaf(d[c], d[c], se);
      sa += fabsf(d[c]);
      sl = fmaf(lv[c], lv[c], sl);
    }
  }
  if (blockIdx.x == 0)  // tail (n not a multiple of 4)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += kLossThreads) {
      const float d = preds[i] - labels[i];
      se = fmaf(d, d, se);
      sa += fabsf(d);
      sl = fmaf(labels[i], labels[i], sl);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = se;
    red[warp][1] = sa;
    red[warp][2] = sl;
  }
  __syncthreads();
  unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch + kLossBlocks * 3);
  if (threadIdx.x == 0) {
    float t[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) {
      t[0] += red[w][0];
      t[1] += red[w][1];
      t[2] += red[w][2];
    }
    scratch[blockIdx.x * 3 + 0] = t[0];
    scratch[blockIdx.x * 3 + 1] = t[1];
    scratch[blockIdx.x * 3 + 2] = t[2];
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {  // one warp sums the table in a fixed order (double accumulation)
    __threadfence();
    double t[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += 32) {
      t[0] += static_cast<double>(scratch[b * 3 + 0]);
      t[1] += static_cast<double>(scratch[b * 3 + 1]);
      t[2] += static_cast<double>(scratch[b * 3 + 2]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t[0] += __shfl_xor_sync(0xffffffffu, t[0], o);
      t[1] += __shfl_xor_sync(0xffffffffu, t[1], o);
      t[2] += __shfl_xor_sync(0xffffffffu, t[2], o);
    }
    if (threadIdx.x == 0) {
      const double inv = 1.0 / static_cast<double>(n);
      const double mse = t[0] * inv, mae = t[1] * inv, ml2 = t[2] * inv;
      out[0] = static_cast<float>(mse);
      out[1] = static_cast<float>(sqrt(mse));
      out[2] = static_cast<float>(mae);
      out[3] = static_cast<float>(mse / ml2);
      out[4] = static_cast<float>(ml2);
      *ticket = 0u;  // ready for the next call on this scratch buffer
    }
  }
}

// dpreds = g_mse 2d/N + g_rmse d/(N rmse) + g_mae sign(d)/N + g_nmse 2d/(N mean(l^2))
__global__ void __launch_bounds__(kLossThreads)
    loss_bwd_kernel(const float* __restrict__ preds, const float* __restrict__ labels, const float* __restrict__ fwd,
                    const float* __restrict__ gout, float* __restrict__ dpreds, size_t n) {
  const float inv_n = 1.f / static_cast<float>(n);
  const float g_mse = gout[0], g_rmse = gout[1], g_mae = gout[2], g_nmse = gout[3];
  const float rmse = fwd[1], ml2 = fwd[4];
  const float kd = inv_n * (2.f * g_mse + (rmse > 0.f ? g_rmse / rmse : 0.f) + 2.f * g_nmse / ml2);
  const float ka = inv_n * g_mae;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kLossThreads + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * kLossThreads) {
    const float d = preds[i] - labels[i];
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    dpreds[i] = fmaf(kd, d, ka * sgn);
  }
}

cudaError_t launch_loss_fwd(const float* preds, const float* labels, size_t n, float* scratch, float* out,
                            cudaStream_t stream) {
  loss_fwd_kernel<<<kLossBlocks, kLossThreads, 0, stream>>>(preds, labels, n, scratch, out);
  return cudaGetLastError();
}
cudaError_t launch_loss_bwd(const float* preds, const float* labels, const float* fwd, const float* gout, float* dpreds,
                            size_t n, cudaStream_t stream) {
  loss_bwd_kernel<<<kLossBlocks * 2, kLossThreads, 0, stream>>>(preds, labels, fwd, gout, dpreds, n);
  return cudaGetLastError();
}
size_t loss_scratch_bytes() { return (kLossBlocks * 3 + 4) * sizeof(float); }

// ------------------------------------------------------------------------------------------------ Adam
constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = kAdamThreads * 4;  // elements per block

struct AdamArgs {
  fno_adam_tensors t;
  int first_block[FNO_ADAM_MAX_TENSORS + 1];  // prefix sum of ceil(n_i / kAdamChunk)
  float lr, beta1, beta2, eps, weight_decay, step_size, inv_bc2_sqrt;
};

// torch.optim.Adam, single-tensor formulation (torch/optim/adam.py _single_tensor_adam):
//   g += wd p;  m = lerp(m, g, 1-b1);  v = b2 v + (1-b2) g g;  p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(kAdamThreads) adam_step_kernel(const __grid_constant__ AdamArgs a) {
  int ti = 0;
#pragma unroll 1
  while (ti + 1 < a.t.count && static_cast<int>(blockIdx.x) >= a.first_block[ti + 1]) ++ti;
  const long long n = a.t.n[ti];
  float* __restrict__ p = static_cast<float*>(a.t.param[ti]);
  const float* __restrict__ g = static_cast<const float*>(a.t.grad[ti]);
  float* __restrict__ m = static_cast<float*>(a.t.exp_avg[ti]);
  float* __restrict__ v = static_cast<float*>(a.t.exp_avg_sq[ti]);
  const long long base = static_cast<long long>(blockIdx.x - a.first_block[ti]) * kAdamChunk;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const long long i = base + r * kAdamThreads + threadIdx.x;
    if (i < n) {
      float gi = g[i];
      const float pi = p[i];
      if (a.weight_decay != 0.f) gi = fmaf(a.weight_decay, pi, gi);
      float mi = m[i], vi = v[i];
      mi = fmaf(gi - mi, 1.f - a.beta1, mi);
      vi = fmaf(1.f - a.beta2, gi * gi, vi * a.beta2);
      const float denom = sqrtf(vi) * a.inv_bc2_sqrt + a.eps;
      m[i] = mi;
      v[i] = vi;
      p[i] = pi - a.step_size * (mi / denom);
    }
  }
}

cudaError_t launch_adam_step(const fno_adam_tensors* t, float lr, float beta1, float beta2, float eps, float weight_decay,
                             long long step, cudaStream_t stream) {
  AdamArgs a;
  a.t = *t;
  int blocks = 0;
  for (int i = 0; i < t->count; ++i) {
    a.first_block[i] = blocks;
    blocks += static_cast<int>((t->n[i] + kAdamChunk - 1) / kAdamChunk);
  }
  a.first_block[t->count] = blocks;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  a.lr = lr;
  a.beta1 = beta1;
  a.beta2 = beta2;
  a.eps = eps;
  a.weight_decay = weight_decay;
  a.step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  a.inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  if (blocks == 0) return cudaSuccess;
  adam_step_kernel<<<blocks, kAdamThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace fno
