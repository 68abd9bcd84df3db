// K2 -- per-mode complex channel mix: Y[b][k][o] = sum_i X[b][k][i] * Wk[k][i][o].
//
// Replaces the two torch.einsum("bixy,ioxy->boxy") corner products plus the zero-filled
// (B,32,64,33) cfloat buffer of the reference (src/models/fno/fno2d.py:54-57, 65-78).
//
// Weight reuse is the whole point of having this as its own phase (SURVEY.md 7 "hard parts"): the
// 2.36 MB of spectral weights per layer are read once per *batch tile*, not once per sample.  A warp
// owns one mode k and keeps Wk[k][:,o] for its lane's output channel o in registers (32 complex =
// 64 regs); the 32 input-channel values of each sample are staged through shared memory and
// broadcast.  Also used for the backward pass with the conj-transposed pack (see fno_pack.cu).
#include "fno_common.cuh"

namespace fno {

constexpr int kMixWarps = 4;
constexpr int kMixThreads = kMixWarps * 32;
constexpr int kMixChunk = 8;        // samples staged per warp iteration
constexpr int kMixTile = 128;       // samples per CTA (32 per warp: the 8 KB weight column load is amortised)

__global__ void __launch_bounds__(kMixThreads)
    mode_mix_kernel(const float2* __restrict__ xm, const float2* __restrict__ wk, float2* __restrict__ ym,
                    int batch) {
  __shared__ __align__(16) float2 xs[kMixWarps][kMixChunk][kC];
  const int k = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // this lane's column of the 32x32 complex weight block of mode k
  float2 w[kC];
  const float2* wk_k = wk + static_cast<size_t>(k) * kC * kC;
#pragma unroll
  for (int i = 0; i < kC; ++i) w[i] = __ldg(wk_k + i * kC + lane);
  pdl_wait();  // the weight column above is not produced by the chain; xm is
  pdl_launch_dependents();

  const int per_warp = kMixTile / kMixWarps;
  const int b_begin = blockIdx.y * kMixTile + warp * per_warp;
  const int b_end = min(b_begin + per_warp, batch);

  // software pipeline: the next chunk's rows are in flight while the current one is multiplied
  auto load_chunk = [&](float2* st, int b0) {
#pragma unroll
    for (int s = 0; s < kMixChunk; ++s)
      st[s] = (b0 + s < b_end) ? __ldg(xm + (static_cast<size_t>(b0 + s) * kModes + k) * kC + lane)
                               : make_float2(0.f, 0.f);
  };
  float2 stage[kMixChunk];
  if (b_begin < b_end) load_chunk(stage, b_begin);

  for (int b0 = b_begin; b0 < b_end; b0 += kMixChunk) {
    const int nb = min(kMixChunk, b_end - b0);
    __syncwarp();  // previous chunk's broadcast reads are done
#pragma unroll
    for (int s = 0; s < kMixChunk; ++s) xs[warp][s][lane] = stage[s];
    __syncwarp();
    if (b0 + kMixChunk < b_end) load_chunk(stage, b0 + kMixChunk);

    // acc_a += xr * (wr, wi);  acc_b += xi * (wr, wi);  y = (a.x - b.y, a.y + b.x)
    float2 acc_a[kMixChunk], acc_b[kMixChunk];
#pragma unroll
    for (int s = 0; s < kMixChunk; ++s) acc_a[s] = acc_b[s] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kC; i += 2) {
#pragma unroll
      for (int s = 0; s < kMixChunk; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(&xs[warp][s][i]);  // broadcast: (xr0, xi0, xr1, xi1)
        acc_a[s] = __ffma2_rn(make_float2(v.x, v.x), w[i], acc_a[s]);
        acc_b[s] = __ffma2_rn(make_float2(v.y, v.y), w[i], acc_b[s]);
        acc_a[s] = __ffma2_rn(make_float2(v.z, v.z), w[i + 1], acc_a[s]);
        acc_b[s] = __ffma2_rn(make_float2(v.w, v.w), w[i + 1], acc_b[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < kMixChunk; ++s)
      if (s < nb)
        ym[(static_cast<size_t>(b0 + s) * kModes + k) * kC + lane] =
            make_float2(acc_a[s].x - acc_b[s].y, acc_a[s].y + acc_b[s].x);
  }
}

cudaError_t launch_mode_mix(const void* xm, const void* wk, void* ym, int batch, cudaStream_t stream) {
  dim3 grid(kModes, (batch + kMixTile - 1) / kMixTile);
  return launch_chained(mode_mix_kernel, grid, dim3(kMixThreads), 0, stream, static_cast<const float2*>(xm),
                        static_cast<const float2*>(wk), static_cast<float2*>(ym), batch);
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
