// Backward-only kernels of the FNO training step (what torch.autograd derives for the reference's
// Fno2d.forward when train_auto.py:255 calls loss["nmse"].backward()).  The data-gradient path of a
// Fourier block reuses the forward kernels (K1 with the c_ky/4096 output scale, K2 with the
// conj-transposed weight pack, K3 with W0 un-transposed and the MUL_DGELU / PLAIN epilogues); this
// file holds what is new in the backward direction:
//   project_bwd_kernel   d(fc1,GELU,fc2,mask): recomputes the 128-wide hidden layer per pixel, emits
//                        dpre of the last block (or d a_L), dz1 for the fc1 weight gradient, and the
//                        fc2 / bias gradients
//   chan_outer_kernel    G[j][i] += sum_{b,pix} P[b][j][pix] Q[b][i][pix]  (1x1-conv weight gradients)
//   spectral_wgrad_kernel  gWk[k][i][o] = sum_b conj(X[b][k][i]) G[b][k][o]   
//   lift_bwd_kernel      gradients of fc0 (spatial feature columns + folded per-sample constants)
#include "fno_common.cuh"

namespace fno {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// GELU and its derivative sharing one erfc evaluation
__device__ __forceinline__ void gelu_both(float x, float& g, float& dg) {
  const float ax = fabsf(x);
  const float e = 0.5f * erfc_abs_scaled(ax);  // 0.5 erfc(|x|/sqrt2)
  g = fmaxf(x, 0.f) - ax * e;
  const float cdf = x >= 0.f ? 1.f - e : e;
  const float pdf = 0.3989422804014327f * ex2_approx(-0.7213475204444817f * x * x);
  dg = fmaf(x, pdf, cdf);
}

// ------------------------------------------------------------------------------------ project bwd
constexpr int kPbThreads = 128;
constexpr int kPbPix = 256;
constexpr int kPbOut = 3 * kProj + 2;   // per-CTA partial row: g_w2 (2 x 128) | g_b1 (128) | g_b2 (2)

// out[i] += sum over parts of partial[part][i] in a FIXED order: the second, deterministic half of every small-gradient
// reduction (the first half = one plain store per CTA).  Replaces float atomics, whose summation order -- and therefore the
// last bits of every gradient and the whole training trajectory -- changed from run to run.
// One launch serves up to three consecutive column segments of the partial rows, each with its own destination (e.g. the
// w2 | b1 | b2 gradients of project_bwd's 386-column rows).  Block = 32 columns x 32 row groups: thread (tx, ty) adds rows
// ty, ty + 32, .. with all its loads in flight (the first version had 8 row groups and ~37 dependent loads per thread:
// 14.5 us per launch, 16 launches per training step), the 32 group sums are then added in index order.
struct ReduceSeg {
  float* out[3];
  int n[3];
  int accumulate;   // 0: out = sum (no memset of the gradient needed), 1: out += sum (further batch chunks)
};
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partial, int n_parts, int row_stride,
                                                               ReduceSeg seg) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  const int n_out = seg.n[0] + seg.n[1] + seg.n[2];
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (i < n_out) {
    int c = ty, u = 0;
    for (; c + 96 < n_parts; c += 128) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += partial[static_cast<size_t>(c + 32 * k) * row_stride + i];
    }
    for (; c < n_parts; c += 32, ++u) s[u & 3] += partial[static_cast<size_t>(c) * row_stride + i];
  }
  red[ty][tx] = (s[0] + s[1]) + (s[2] + s[3]);
  __syncthreads();
  if (ty == 0 && i < n_out) {
    float t = red[0][tx];
#pragma unroll
    for (int r = 1; r < 32; ++r) t += red[r][tx];
    float* dst = i < seg.n[0] ? seg.out[0] + i
                 : (i < seg.n[0] + seg.n[1] ? seg.out[1] + (i - seg.n[0]) : seg.out[2] + (i - seg.n[0] - seg.n[1]));
    *dst = seg.accumulate ? *dst + t : t;
  }
}
cudaError_t launch_reduce_partials(const float* partial, int n_parts, int row_stride, float* out0, int n0, float* out1, int n1,
                                   float* out2, int n2, int accumulate, cudaStream_t stream) {
  const int n_out = n0 + n1 + n2;
  if (n_parts <= 0 || n_out <= 0) return cudaSuccess;
  ReduceSeg seg;
  seg.out[0] = out0; seg.out[1] = out1; seg.out[2] = out2;
  seg.n[0] = n0; seg.n[1] = n1; seg.n[2] = n2;
  seg.accumulate = accumulate;
  reduce_partials_kernel<<<(n_out + 31) / 32, 1024, 0, stream>>>(partial, n_parts, row_stride, seg);
  return cudaGetLastError();
}

template <typename TAct>
struct PbSmem {
  alignas(128) TAct xs[kC][kPbPix];
  alignas(16) float w1[kProj][kC];
  alignas(16) float b1[kProj];
  alignas(16) float w2[2][kProj];
  alignas(16) float acc[kPbThreads / 32][3][kProj];   // per-warp partials of (g_w2[0][j], g_w2[1][j], g_b1[j]): fixed order
  alignas(16) float acc_b2[kPbThreads / 32][2];
  alignas(8) uint64_t bar;
};

template <typename TAct>
__global__ void __launch_bounds__(kPbThreads)
    project_bwd_kernel(const TAct* __restrict__ a,        // [B][32][4096]  a_L
                       const float* __restrict__ dpreds,  // [B][2][4096]
                       const float* __restrict__ mask,    // [B][4096]
                       const float* __restrict__ pre,     // [B][32][4096] pre-activation of the last block (or null)
                       const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                       float* __restrict__ d_out,         // [B][32][4096]: dpre_{L-1} (pre != null) or d a_L
                       float* __restrict__ dz1,           // [B][128][4096]
                       float* __restrict__ partial) {   // [CTA][kPbOut]: (g_w2 256 | g_b1 128 | g_b2 2), reduced in CTA order
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PbSmem<TAct>& sm = *reinterpret_cast<PbSmem<TAct>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.y;
  const int pix0 = blockIdx.x * kPbPix;

  if (tid == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid < kC) {
    constexpr uint32_t bytes = kPbPix * sizeof(TAct);
    if (tid == 0) mbar_expect_tx(&sm.bar, kC * bytes);
    __syncwarp();
    bulk_g2s(&sm.xs[tid][0], a + (static_cast<size_t>(b) * kC + tid) * kHW + pix0, bytes, &sm.bar);
  }
  for (int i = tid; i < kProj * kC; i += kPbThreads) (&sm.w1[0][0])[i] = w1[i];
  for (int j = tid; j < kProj; j += kPbThreads) {
    sm.b1[j] = b1[j];
    sm.w2[0][j] = w2[j];
    sm.w2[1][j] = w2[kProj + j];
  }
  __syncthreads();
  mbar_wait(&sm.bar, 0);

  const int pix = pix0 + 2 * tid;
  float2 x[kC], da[kC];
#pragma unroll
  for (int i = 0; i < kC; ++i) {
    if constexpr (sizeof(TAct) == 4) {
      x[i] = *reinterpret_cast<const float2*>(&sm.xs[i][2 * tid]);
    } else {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(&sm.xs[i][2 * tid]);
      x[i] = make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
    }
    da[i] = make_float2(0.f, 0.f);
  }
  const float2 m = *reinterpret_cast<const float2*>(mask + static_cast<size_t>(b) * kHW + pix);
  float2 d0 = *reinterpret_cast<const float2*>(dpreds + (static_cast<size_t>(b) * 2 + 0) * kHW + pix);
  float2 d1 = *reinterpret_cast<const float2*>(dpreds + (static_cast<size_t>(b) * 2 + 1) * kHW + pix);
  d0.x *= m.x; d0.y *= m.y; d1.x *= m.x; d1.y *= m.y;

  float* dz_b = dz1 + static_cast<size_t>(b) * kProj * kHW + pix;
#pragma unroll 1
  for (int j = 0; j < kProj; ++j) {
    float wr[kC];
    const float4* wrow = reinterpret_cast<const float4*>(&sm.w1[j][0]);
#pragma unroll
    for (int q = 0; q < kC / 4; ++q) {
      const float4 t = wrow[q];
      wr[4 * q] = t.x; wr[4 * q + 1] = t.y; wr[4 * q + 2] = t.z; wr[4 * q + 3] = t.w;
    }
    const float bj = sm.b1[j];
    float2 z0 = make_float2(bj, bj), z1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kC; i += 2) {
      z0 = __ffma2_rn(x[i], make_float2(wr[i], wr[i]), z0);
      z1 = __ffma2_rn(x[i + 1], make_float2(wr[i + 1], wr[i + 1]), z1);
    }
    const float2 z = make_float2(z0.x + z1.x, z0.y + z1.y);
    float gx_, gy_, dgx, dgy;
    gelu_both(z.x, gx_, dgx);
    gelu_both(z.y, gy_, dgy);
    const float w20 = sm.w2[0][j], w21 = sm.w2[1][j];
    const float2 dz = make_float2((w20 * d0.x + w21 * d1.x) * dgx, (w20 * d0.y + w21 * d1.y) * dgy);
#pragma unroll
    for (int i = 0; i < kC; ++i) da[i] = __ffma2_rn(make_float2(wr[i], wr[i]), dz, da[i]);
    *reinterpret_cast<float2*>(dz_b + static_cast<size_t>(j) * kHW) = dz;
    // fc2 weight gradient and fc1 bias gradient: warp partials -> smem accumulators
    const float p0 = warp_sum(d0.x * gx_ + d0.y * gy_);
    const float p1 = warp_sum(d1.x * gx_ + d1.y * gy_);
    const float pb = warp_sum(dz.x + dz.y);
    if (lane == 0) {   // no atomics anywhere in the gradient path: every slot has one writer, every sum a fixed order
      sm.acc[tid >> 5][0][j] = p0;
      sm.acc[tid >> 5][1][j] = p1;
      sm.acc[tid >> 5][2][j] = pb;
    }
  }
  // fc2 bias gradient
  {
    const float s0 = warp_sum(d0.x + d0.y), s1 = warp_sum(d1.x + d1.y);
    if (lane == 0) {
      sm.acc_b2[tid >> 5][0] = s0;
      sm.acc_b2[tid >> 5][1] = s1;
    }
  }
  // d a_L (optionally times GELU'(pre_{L-1}))
  const size_t base = static_cast<size_t>(b) * kC * kHW + pix;
#pragma unroll
  for (int i = 0; i < kC; ++i) {
    float2 v = da[i];
    if (pre != nullptr) {
      const float2 pv = *reinterpret_cast<const float2*>(pre + base + static_cast<size_t>(i) * kHW);
      v.x *= dgelu_erf(pv.x);
      v.y *= dgelu_erf(pv.y);
    }
    *reinterpret_cast<float2*>(d_out + base + static_cast<size_t>(i) * kHW) = v;
  }
  __syncthreads();
  float* prow = partial + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * kPbOut;
  for (int j = tid; j < kProj; j += kPbThreads) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < kPbThreads / 32; ++w) { t0 += sm.acc[w][0][j]; t1 += sm.acc[w][1][j]; t2 += sm.acc[w][2][j]; }
    prow[j] = t0;
    prow[kProj + j] = t1;
    prow[2 * kProj + j] = t2;
  }
  if (tid < 2) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kPbThreads / 32; ++w) t += sm.acc_b2[w][tid];
    prow[3 * kProj + tid] = t;
  }
}

template <typename TAct>
cudaError_t launch_project_bwd(const void* a, const float* dpreds, const float* mask, const float* pre,
                               const float* w1, const float* b1, const float* w2, float* d_out, float* dz1,
                               float* partial, int batch, cudaStream_t stream) {
  auto kern = project_bwd_kernel<TAct>;
  constexpr size_t smem = sizeof(PbSmem<TAct>);
  static PerDeviceLaunch pd;
  cudaError_t e0 = per_device_setup(kern, smem, pd);
  if (e0 != cudaSuccess) return e0;
  dim3 grid(kHW / kPbPix, batch);
  kern<<<grid, kPbThreads, smem, stream>>>(static_cast<const TAct*>(a), dpreds, mask, pre, w1, b1, w2, d_out, dz1,
                                           partial);
  return cudaGetLastError();
}
int project_bwd_parts(int batch) { return (kHW / kPbPix) * batch; }   // partial rows written by one launch
int project_bwd_row() { return kPbOut; }
template cudaError_t launch_project_bwd<float>(const void*, const float*, const float*, const float*, const float*,
                                               const float*, const float*, float*, float*, float*,
                                               int, cudaStream_t);
template cudaError_t launch_project_bwd<__nv_bfloat16>(const void*, const float*, const float*, const float*,
                                                       const float*, const float*, const float*, float*, float*,
                                                       float*, int, cudaStream_t);

// ------------------------------------------------------------------------------------- chan outer
// out[j][i] += sum_{b,pix} P[b][j][pix] * Q[b][i][pix];   rowsum[j] += sum_{b,pix} P[b][j][pix]
constexpr int kCoThreads = 256;

// 4 consecutive pixels of a staged row as fp32 (bf16 rows are widened by a shift)
template <typename T>
__device__ __forceinline__ float4 co_ld4(const unsigned char* row, int px);
template <>
__device__ __forceinline__ float4 co_ld4<float>(const unsigned char* row, int px) {
  return *reinterpret_cast<const float4*>(row + px * 4);
}
template <>
__device__ __forceinline__ float4 co_ld4<__nv_bfloat16>(const unsigned char* row, int px) {
  const uint2 v = *reinterpret_cast<const uint2*>(row + px * 2);
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                     __uint_as_float(v.y & 0xffff0000u));
}

// G[j][i] = sum_{b,pix} P[b][j][pix] Q[b][i][pix] and the row sums of P (the bias gradient): the 1x1-convolution weight
// gradients.  Round 1's version gave every thread a 2x2 (4x4) output tile: one 16-byte shared-memory read per 4 fma, so
// the kernel ran at the shared-memory rate, a quarter of the fma rate (68 / 98 us per launch at B = 64, 30 % of a
// training step).  Here a thread owns an 8 x 4 output tile (1.5 B of shared memory per fma) and the pixels of a chunk are
// split among NG thread groups whose partial tiles are added in a fixed order at the end; chunks arrive through a
// two-stage cp.async ring, so the loads of the next chunk overlap the arithmetic of the current one.
// Rows are assigned interleaved (j = a * NJ/8 + tj, i = c * NI/4 + ti) so that the rows a warp reads at the same time are
// consecutive and, with the 16-byte row skew, fall into different banks.
template <typename TP, typename TQ, int NJ, int NI>
struct CoCfg {
  static constexpr int TJ = 8, TI = 4;
  static constexpr int G = NJ * NI / (TJ * TI);          // threads per pixel group
  static constexpr int NG = kCoThreads / G;              // pixel groups
  static constexpr int PIX = (NJ >= 128) ? 64 : 128;     // pixels per chunk
  static constexpr int PXG = PIX / NG;                   // pixels per group and chunk
  static constexpr int NTI = NI / TI, NTJ = NJ / TJ;
  static constexpr int PITCH_P = PIX * static_cast<int>(sizeof(TP)) + 16, PITCH_Q = PIX * static_cast<int>(sizeof(TQ)) + 16;
  static constexpr int STAGE = NJ * PITCH_P + NI * PITCH_Q;
  static constexpr int OUT = NJ * NI + NJ;
  static constexpr size_t SMEM = (2 * STAGE > NG * OUT * 4) ? 2 * STAGE : NG * OUT * 4;
  static_assert(G * NG == kCoThreads && PXG % 4 == 0 && NJ % TJ == 0 && NI % TI == 0, "tile mismatch");
};

template <typename TP, typename TQ, int NJ, int NI>
__global__ void __launch_bounds__(kCoThreads)
    chan_outer_kernel(const TP* __restrict__ P, const TQ* __restrict__ Q, float* __restrict__ partial, int batch) {
  // partial[CTA][NJ*NI + NJ]: this CTA's share of G[j][i] and of the row sums sum_pix P[j][pix] (the bias gradient)
  using Cfg = CoCfg<TP, TQ, NJ, NI>;
  constexpr int TJ = Cfg::TJ, TI = Cfg::TI, PIX = Cfg::PIX;
  extern __shared__ __align__(16) unsigned char co_smem[];
  const int tid = threadIdx.x;
  const int grp = tid / Cfg::G, t = tid % Cfg::G;
  const int tj = t / Cfg::NTI, ti = t % Cfg::NTI;
  float acc[TJ][TI];
  float rs[TJ];
#pragma unroll
  for (int a = 0; a < TJ; ++a) {
    rs[a] = 0.f;
#pragma unroll
    for (int c = 0; c < TI; ++c) acc[a][c] = 0.f;
  }
  const int chunks = kHW / PIX;
  const int items = batch * chunks;
  auto stage_load = [&](int it, int st) {   // all threads: 16-byte cp.async pieces of the item's P and Q rows
    const int b = it / chunks, p0 = (it % chunks) * PIX;
    unsigned char* base = co_smem + st * Cfg::STAGE;
    constexpr int CP = PIX * static_cast<int>(sizeof(TP)) / 16, CQ = PIX * static_cast<int>(sizeof(TQ)) / 16;
    for (int e = tid; e < NJ * CP; e += kCoThreads) {
      const int j = e / CP, q = e % CP;
      const char* src = reinterpret_cast<const char*>(P + (static_cast<size_t>(b) * NJ + j) * kHW + p0) + q * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(base + j * Cfg::PITCH_P + q * 16)), "l"(src) : "memory");
    }
    for (int e = tid; e < NI * CQ; e += kCoThreads) {
      const int i = e / CQ, q = e % CQ;
      const char* src = reinterpret_cast<const char*>(Q + (static_cast<size_t>(b) * NI + i) * kHW + p0) + q * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(base + NJ * Cfg::PITCH_P + i * Cfg::PITCH_Q + q * 16)),
                   "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int n = 0;
  if (static_cast<int>(blockIdx.x) < items) stage_load(blockIdx.x, 0);
  for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
    const int nxt = it + gridDim.x;
    if (nxt < items) {
      stage_load(nxt, (n + 1) & 1);   // the buffer was released by the barrier at the end of the previous iteration
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const unsigned char* ps = co_smem + (n & 1) * Cfg::STAGE;
    const unsigned char* qs = ps + NJ * Cfg::PITCH_P;
#pragma unroll 2
    for (int q = 0; q < Cfg::PXG / 4; ++q) {
      const int px = grp * Cfg::PXG + 4 * q;
      float4 pv[TJ], qv[TI];
#pragma unroll
      for (int a = 0; a < TJ; ++a) pv[a] = co_ld4<TP>(ps + (a * Cfg::NTJ + tj) * Cfg::PITCH_P, px);
#pragma unroll
      for (int c = 0; c < TI; ++c) qv[c] = co_ld4<TQ>(qs + (c * Cfg::NTI + ti) * Cfg::PITCH_Q, px);
#pragma unroll
      for (int a = 0; a < TJ; ++a) {
        if (ti == 0) rs[a] += (pv[a].x + pv[a].y) + (pv[a].z + pv[a].w);
#pragma unroll
        for (int c = 0; c < TI; ++c)
          acc[a][c] = fmaf(pv[a].x, qv[c].x, fmaf(pv[a].y, qv[c].y, fmaf(pv[a].z, qv[c].z, fmaf(pv[a].w, qv[c].w, acc[a][c]))));
      }
    }
    __syncthreads();   // everybody is done with this stage: the next iteration may refill it
  }
  // the NG pixel groups' tiles, added in group order (fixed summation tree)
  float* red = reinterpret_cast<float*>(co_smem);
#pragma unroll
  for (int a = 0; a < TJ; ++a) {
    const int j = a * Cfg::NTJ + tj;
#pragma unroll
    for (int c = 0; c < TI; ++c) red[grp * Cfg::OUT + j * NI + c * Cfg::NTI + ti] = acc[a][c];
    if (ti == 0) red[grp * Cfg::OUT + NJ * NI + j] = rs[a];
  }
  __syncthreads();
  float* prow = partial + static_cast<size_t>(blockIdx.x) * Cfg::OUT;
  for (int e = tid; e < Cfg::OUT; e += kCoThreads) {
    float v = red[e];
#pragma unroll
    for (int g = 1; g < Cfg::NG; ++g) v += red[g * Cfg::OUT + e];
    prow[e] = v;
  }
}

// returns the number of partial rows written (= grid size) through *n_parts
template <typename TP, typename TQ, int NJ, int NI>
cudaError_t launch_chan_outer(const void* P, const void* Q, float* partial, int* n_parts, int batch, cudaStream_t stream) {
  using Cfg = CoCfg<TP, TQ, NJ, NI>;
  auto kern = chan_outer_kernel<TP, TQ, NJ, NI>;
  constexpr size_t smem = Cfg::SMEM;
  static PerDeviceLaunch pd;
  cudaError_t e0 = per_device_setup(kern, smem, pd);
  if (e0 != cudaSuccess) return e0;
  const int items = batch * (kHW / Cfg::PIX);
  const int grid = items < 296 ? items : 296;   // 2 CTAs per SM; also the row count of the partial buffer
  kern<<<grid, kCoThreads, smem, stream>>>(static_cast<const TP*>(P), static_cast<const TQ*>(Q), partial, batch);
  *n_parts = grid;
  return cudaGetLastError();
}
template cudaError_t launch_chan_outer<float, float, 128, 32>(const void*, const void*, float*, int*, int, cudaStream_t);
template cudaError_t launch_chan_outer<float, __nv_bfloat16, 128, 32>(const void*, const void*, float*, int*, int, cudaStream_t);
template cudaError_t launch_chan_outer<float, float, 32, 32>(const void*, const void*, float*, int*, int, cudaStream_t);
template cudaError_t launch_chan_outer<float, __nv_bfloat16, 32, 32>(const void*, const void*, float*, int*, int, cudaStream_t);

// --------------------------------------------------------------------------------- spectral wgrad
// gWk[k][i][o] = sum_b conj(X[k][b][i]) * G[k][b][o];  one CTA per mode k, lane = o, warps split the batch.
constexpr int kSwWarps = 4;

__global__ void __launch_bounds__(kSwWarps * 32)
    spectral_wgrad_kernel(const float2* __restrict__ xm, const float2* __restrict__ gm, float2* __restrict__ gwk,
                          int batch) {
  __shared__ __align__(16) float2 xs[kSwWarps][kC];
  __shared__ __align__(16) float2 red[kSwWarps][kC][kC + 1];
  const int k = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2 acc_a[kC], acc_b[kC];  // a += xr*(gr,gi); b += xi*(gr,gi);  conj(x)*g = (a.x + b.y, a.y - b.x)
#pragma unroll
  for (int i = 0; i < kC; ++i) acc_a[i] = acc_b[i] = make_float2(0.f, 0.f);
  // the loads of the next three samples of this warp are in flight while the current one is multiplied (without the
  // prefetch every sample paid a full DRAM round trip: 44.7 us per launch at B = 256 for 38 MB)
  constexpr int kAhead = 3;
  float2 xq[kAhead], gq[kAhead];
#pragma unroll
  for (int u = 0; u < kAhead; ++u) {
    const int bu = warp + u * kSwWarps;
    if (bu < batch) {
      const size_t o = (static_cast<size_t>(k) * batch + bu) * kC + lane;
      xq[u] = __ldg(xm + o);
      gq[u] = __ldg(gm + o);
    }
  }
  for (int b = warp; b < batch; b += kSwWarps) {
    const int bn = b + kAhead * kSwWarps;
    float2 xn = make_float2(0.f, 0.f), gn = xn;
    if (bn < batch) {
      const size_t o = (static_cast<size_t>(k) * batch + bn) * kC + lane;
      xn = __ldg(xm + o);
      gn = __ldg(gm + o);
    }
    float2 xv, g;
    // rotating register queue with compile-time indices
    xv = xq[0]; g = gq[0];
#pragma unroll
    for (int u = 0; u < kAhead - 1; ++u) { xq[u] = xq[u + 1]; gq[u] = gq[u + 1]; }
    xq[kAhead - 1] = xn; gq[kAhead - 1] = gn;
    __syncwarp();
    xs[warp][lane] = xv;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kC; i += 2) {
      const float4 v = *reinterpret_cast<const float4*>(&xs[warp][i]);
      acc_a[i] = __ffma2_rn(make_float2(v.x, v.x), g, acc_a[i]);
      acc_b[i] = __ffma2_rn(make_float2(v.y, v.y), g, acc_b[i]);
      acc_a[i + 1] = __ffma2_rn(make_float2(v.z, v.z), g, acc_a[i + 1]);
      acc_b[i + 1] = __ffma2_rn(make_float2(v.w, v.w), g, acc_b[i + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < kC; ++i) red[warp][i][lane] = make_float2(acc_a[i].x + acc_b[i].y, acc_a[i].y - acc_b[i].x);
  __syncthreads();
  for (int e = threadIdx.x; e < kC * kC; e += kSwWarps * 32) {
    const int i = e / kC, o = e % kC;
    float2 s = red[0][i][o];
#pragma unroll
    for (int w = 1; w < kSwWarps; ++w) {
      s.x += red[w][i][o].x;
      s.y += red[w][i][o].y;
    }
    gwk[static_cast<size_t>(k) * kC * kC + e] = s;
  }
}

cudaError_t launch_spectral_wgrad(const void* xm, const void* gm, void* gwk, int batch, cudaStream_t stream) {
  spectral_wgrad_kernel<<<kModes, kSwWarps * 32, 0, stream>>>(static_cast<const float2*>(xm), static_cast<const float2*>(gm),
                                                             static_cast<float2*>(gwk), batch);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------- lift bwd
// g_w[c][q] (q<5: u,v,mask,x,y; q>=5: case params), g_b[c] from d a0.  One CTA per (channel c, sample
// slice); per-sample plane sums T[b][c] carry the bias and case-parameter columns.
constexpr int kLbThreads = 256;

__global__ void __launch_bounds__(kLbThreads)
    lift_bwd_kernel(const float* __restrict__ da0,     // [B][32][4096]
                    const float* __restrict__ inputs,  // [B][2][4096]
                    const float* __restrict__ mask,    // [B][4096]
                    const float* __restrict__ params,  // [B][p]
                    const float* __restrict__ gx, const float* __restrict__ gy, float* __restrict__ partial,
                    int batch, int p) {   // partial[blockIdx.y][c][nin + 1]: weight row of channel c, then its bias
  // Every thread accumulates its share of all six sums over ALL its samples (the case-parameter columns are linear in the
  // per-sample plane sum, so its per-thread part is weighted by params[b][q] on the fly); one block reduction at the end.
  // (Round 1 reduced over the block after every sample: two barriers per 16 KB plane, 80 us at B = 256 for 134 MB.)
  __shared__ float red[kLbThreads / 32][6 + kMaxCaseParams];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nin = 5 + p;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gp_acc[kMaxCaseParams];
#pragma unroll
  for (int q = 0; q < kMaxCaseParams; ++q) gp_acc[q] = 0.f;
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const float* d = da0 + (static_cast<size_t>(b) * kC + c) * kHW;
    float dplane = 0.f;   // this thread's share of the plane sum of sample b
#pragma unroll 4
    for (int px = tid * 4; px < kHW; px += kLbThreads * 4) {
      const float4 dv = *reinterpret_cast<const float4*>(d + px);
      const float4 u = *reinterpret_cast<const float4*>(inputs + (static_cast<size_t>(b) * 2 + 0) * kHW + px);
      const float4 v = *reinterpret_cast<const float4*>(inputs + (static_cast<size_t>(b) * 2 + 1) * kHW + px);
      const float4 m = *reinterpret_cast<const float4*>(mask + static_cast<size_t>(b) * kHW + px);
      const float4 yw = *reinterpret_cast<const float4*>(gy + (px & 63));
      const float xh = gx[px >> 6];
      const float dsum = (dv.x + dv.y) + (dv.z + dv.w);
      s[0] += dv.x * u.x + dv.y * u.y + dv.z * u.z + dv.w * u.w;
      s[1] += dv.x * v.x + dv.y * v.y + dv.z * v.z + dv.w * v.w;
      s[2] += dv.x * m.x + dv.y * m.y + dv.z * m.z + dv.w * m.w;
      s[3] += dsum * xh;
      s[4] += dv.x * yw.x + dv.y * yw.y + dv.z * yw.z + dv.w * yw.w;
      dplane += dsum;
    }
    s[5] += dplane;
#pragma unroll
    for (int q = 0; q < kMaxCaseParams; ++q)
      if (q < p) gp_acc[q] = fmaf(dplane, params[b * p + q], gp_acc[q]);
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const float r = warp_sum(s[q]);
    if (lane == 0) red[warp][q] = r;
  }
#pragma unroll
  for (int q = 0; q < kMaxCaseParams; ++q) {
    const float r = warp_sum(gp_acc[q]);
    if (lane == 0) red[warp][6 + q] = r;
  }
  __syncthreads();
  if (tid < 6 + kMaxCaseParams) {
    float t = 0.f;
    for (int w = 0; w < kLbThreads / 32; ++w) t += red[w][tid];   // warp order: fixed
    float* prow = partial + (static_cast<size_t>(blockIdx.y) * kC + c) * (nin + 1);
    if (tid < 5) prow[tid] = t;
    else if (tid == 5) prow[nin] = t;
    else if (tid - 6 < p) prow[5 + tid - 6] = t;
  }
}

// g_w[c][q] += sum_y partial[y][c][q], g_b[c] += sum_y partial[y][c][nin]   (fixed order)
__global__ void lift_bwd_reduce_kernel(const float* __restrict__ partial, int n_parts, int nin, float* __restrict__ g_w,
                                       float* __restrict__ g_b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kC * (nin + 1)) return;
  const int c = i / (nin + 1), q = i % (nin + 1);
  float s = 0.f;
  for (int y = 0; y < n_parts; ++y) s += partial[(static_cast<size_t>(y) * kC + c) * (nin + 1) + q];
  if (q < nin) g_w[c * nin + q] = s;   // the only contribution: no memset of the gradient needed
  else g_b[c] = s;
}

cudaError_t launch_lift_bwd(const float* da0, const float* inputs, const float* mask, const float* params,
                            const float* gx, const float* gy, float* g_w, float* g_b, float* partial, int batch, int p,
                            cudaStream_t stream) {
  if (p < 0 || p > kMaxCaseParams) return cudaErrorInvalidValue;
  dim3 grid(kC, batch < 16 ? batch : 16);
  lift_bwd_kernel<<<grid, kLbThreads, 0, stream>>>(da0, inputs, mask, params, gx, gy, partial, batch, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int n = kC * (5 + p + 1);
  lift_bwd_reduce_kernel<<<(n + 127) / 128, 128, 0, stream>>>(partial, static_cast<int>(grid.y), 5 + p, g_w, g_b);
  return cudaGetLastError();
}

}  // namespace fno
