// K2 on the tensor cores -- per-mode complex channel mix  Y[b][k][o] = sum_i X[b][k][i] * Wk[k][i][o]
// replacing the two torch.einsum("bixy,ioxy->boxy") corner products and the zero-filled (B,32,64,33) cfloat
// buffer of the reference (src/models/fno/fno2d.py:54-57, 65-78).
//
// For one mode k the mix is a real GEMM over 128 samples:
//     D[128 samples][64 = (o, re|im)] = A[128][64 = (i, re|im)] * B^T,
//     B[(o,re)][(i,re)] = Wre, B[(o,re)][(i,im)] = -Wim, B[(o,im)][(i,re)] = Wim, B[(o,im)][(i,im)] = Wre,
// i.e. exactly the interleaved complex64 rows X[b][k][:] and Y[b][k][:] as they sit in memory: a thread loads
// its sample's 256-byte row, splits it into tf32 hi/lo (3xTF32, round-to-nearest) and stores it as one row of the
// K-major A operand; 24 UMMAs (M=128, N=64, K=8) accumulate in TMEM; the thread reads its output row back and
// writes 256 contiguous bytes.  The 2.36 MB of spectral weights per layer are read once per 128 samples.
// Persistent CTA = two independent 128-thread pipelines (thread = sample row of the tile).
// With the conj-transposed pack (fno_mode_mix.cu) the same kernel is the adjoint mix of the backward pass.
#include "fno_common.cuh"
#include "tc_common.cuh"

namespace fno {

constexpr int kMxThreads = 256;
constexpr int kMxGroup = 128;
constexpr int kMxM = 128;                       // samples per tile
constexpr int kMxK = 2 * kC;                    // 64 real (i, re|im)
constexpr int kMxN = 2 * kC;                    // 64 real (o, re|im)
constexpr uint32_t kMxLboA = (kMxM / 8) * 128;  // 2048
constexpr uint32_t kMxLboB = (kMxN / 8) * 128;  // 1024

struct MxSmem {
  alignas(128) float a_hi[2][kMxM * kMxK];  // [pipeline] 2 x 32 KB
  alignas(128) float a_lo[2][kMxM * kMxK];
  alignas(128) float b_hi[2][kMxN * kMxK];  // [pipeline] 2 x 16 KB
  alignas(128) float b_lo[2][kMxN * kMxK];
  alignas(8) uint64_t mma_bar[2];
  uint32_t tmem_base;
};

template <int GRP>
__device__ __forceinline__ void mx_group_barrier() {
  asm volatile("bar.sync %0, %1;" ::"n"(GRP + 1), "n"(kMxGroup) : "memory");
}

template <int GRP>
__device__ __forceinline__ void mx_pipeline(MxSmem& sm, const float4* __restrict__ xm, const float4* __restrict__ wk,
                                            float4* __restrict__ ym, int batch, int n_btiles, int n_tiles) {
  const int gtid = threadIdx.x & (kMxGroup - 1), gwarp = gtid >> 5;
  const uint32_t tmem_acc = sm.tmem_base + GRP * kMxN;
  constexpr uint32_t idesc = tc::make_idesc_tf32(kMxM, kMxN);
  float* a_hi = sm.a_hi[GRP];
  float* a_lo = sm.a_lo[GRP];
  float* b_hi = sm.b_hi[GRP];
  float* b_lo = sm.b_lo[GRP];

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_cta = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;
  const int n_mine = (n_cta + 1 - GRP) / 2;

  for (int it = 0; it < n_mine; ++it) {
    const int tile = first + (2 * it + GRP) * stride;
    const int k = tile / n_btiles;
    const int b = (tile % n_btiles) * kMxM + gtid;  // this thread's sample
    const bool valid = b < batch;
    const size_t row = (static_cast<size_t>(valid ? b : 0) * kModes + k) * (kMxK / 4);  // float4 index of X[b][k][0]

    // ---- A operand: this sample's 64 floats (32 complex, interleaved) -> one K-major row, hi/lo
#pragma unroll
    for (int q = 0; q < kMxK / 4; ++q) {
      float4 v = valid ? __ldg(xm + row + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 hi, lo;
      tc::split_tf32(v.x, hi.x, lo.x);
      tc::split_tf32(v.y, hi.y, lo.y);
      tc::split_tf32(v.z, hi.z, lo.z);
      tc::split_tf32(v.w, hi.w, lo.w);
      const uint32_t off = tc::kmajor_offset(gtid, 4 * q, kMxM) / 4;
      *reinterpret_cast<float4*>(a_hi + off) = hi;
      *reinterpret_cast<float4*>(a_lo + off) = lo;
    }
    // ---- B operand from Wk[k][i][o] (complex): rows n = 2o (re), 2o+1 (im); columns kk = 2i (re), 2i+1 (im)
    const float4* wk_k = wk + static_cast<size_t>(k) * (kC * kC / 2);  // float4 = 2 complex (o, o+1)
#pragma unroll
    for (int rep = 0; rep < (kC * kC / 2) / kMxGroup; ++rep) {
      const int e = rep * kMxGroup + gtid;  // e -> (i = e / 16, o = 2 (e % 16))
      const int i = e >> 4, o = (e & 15) * 2;
      const float4 w = __ldg(wk_k + e);     // (wr0, wi0, wr1, wi1)
      const float wr[2] = {w.x, w.z}, wi[2] = {w.y, w.w};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float rh, rl, ih, il;
        tc::split_tf32(wr[c], rh, rl);
        tc::split_tf32(wi[c], ih, il);
        const uint32_t off_re = tc::kmajor_offset(2 * (o + c), 2 * i, kMxN) / 4;      // row (o,re)
        const uint32_t off_im = tc::kmajor_offset(2 * (o + c) + 1, 2 * i, kMxN) / 4;  // row (o,im)
        *reinterpret_cast<float2*>(b_hi + off_re) = make_float2(rh, -ih);
        *reinterpret_cast<float2*>(b_lo + off_re) = make_float2(rl, -il);
        *reinterpret_cast<float2*>(b_hi + off_im) = make_float2(ih, rh);
        *reinterpret_cast<float2*>(b_lo + off_im) = make_float2(il, rl);
      }
    }
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    mx_group_barrier<GRP>();
    tc::fence_after_thread_sync();
    if (gwarp == 0) {
      if (tc::elect_one()) {
        const uint32_t a_s[3] = {tc::smem_addr(a_hi), tc::smem_addr(a_lo), tc::smem_addr(a_hi)};
        const uint32_t b_s[3] = {tc::smem_addr(b_hi), tc::smem_addr(b_hi), tc::smem_addr(b_lo)};
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t da0 = tc::make_smem_desc(a_s[pass], kMxLboA, 128);
          const uint64_t db0 = tc::make_smem_desc(b_s[pass], kMxLboB, 128);
#pragma unroll
          for (int ks = 0; ks < kMxK / 8; ++ks) {
            const uint64_t da = da0 + ((ks * 2 * kMxLboA) >> 4), db = db0 + ((ks * 2 * kMxLboB) >> 4);
            if (pass == 0 && ks == 0) tc::mma_tf32_imm<false>(tmem_acc, da, db, idesc);
            else tc::mma_tf32_imm<true>(tmem_acc, da, db, idesc);
          }
        }
        tc::mma_commit(&sm.mma_bar[GRP]);
      }
      __syncwarp();
    }
    // ---- epilogue: this sample's output row
    mbar_wait(&sm.mma_bar[GRP], it & 1);
    tc::fence_after_thread_sync();
#pragma unroll
    for (int chunk = 0; chunk < 2; ++chunk) {
      float v[32];
      tc::tmem_ld32(tmem_acc + (static_cast<uint32_t>(gwarp * 32) << 16) + chunk * 32, v);
      if (valid) {
        float4* dst = ym + row + chunk * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
    tc::fence_before_thread_sync();
    mx_group_barrier<GRP>();  // TMEM reads done before the next tile's first MMA overwrites the accumulator
  }
}

__global__ void __launch_bounds__(kMxThreads, 1)
    mode_mix_tc_kernel(const float4* __restrict__ xm, const float4* __restrict__ wk, float4* __restrict__ ym, int batch,
                       int n_btiles, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // no pointer arithmetic: keeps LDS/STS addressing
  MxSmem& sm = *reinterpret_cast<MxSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 127u) != 0) __trap();
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&sm.mma_bar[0], 1);
    mbar_init(&sm.mma_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<2 * kMxN>(&sm.tmem_base);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  if (tid < kMxGroup) mx_pipeline<0>(sm, xm, wk, ym, batch, n_btiles, n_tiles);
  else mx_pipeline<1>(sm, xm, wk, ym, batch, n_btiles, n_tiles);
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<2 * kMxN>(sm.tmem_base);
}

cudaError_t launch_mode_mix_tc(const void* xm, const void* wk, void* ym, int batch, cudaStream_t stream) {
  constexpr size_t smem = sizeof(MxSmem);
  static bool configured = false;
  static int n_sm = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mode_mix_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int n_btiles = (batch + kMxM - 1) / kMxM;
  const int n_tiles = kModes * n_btiles;
  const int grid = n_tiles < 2 * n_sm ? (n_tiles + 1) / 2 : n_sm;
  mode_mix_tc_kernel<<<grid, kMxThreads, smem, stream>>>(static_cast<const float4*>(xm), static_cast<const float4*>(wk),
                                                         static_cast<float4*>(ym), batch, n_btiles, n_tiles);
  return cudaGetLastError();
}

}  // namespace fno
