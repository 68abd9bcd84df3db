// Project kernel for bf16 activation storage, warp-specialised:  preds = (fc2 . GELU . fc1)(a_L) * mask
// replacing Conv2d(32,128,1) + GELU + Conv2d(128,2,1) + "* mask" (reference src/models/fno/fno2d.py:228-233).
//
// project_tc_kernel (fno_project_tc.cu, still the fp32-storage path) runs prefetch -> split -> store -> barrier -> MMA ->
// epilogue back to back inside each of its two pipelines: 85 us per launch at B=256, while its GELU epilogue alone -- the
// floor: 134 M erf-GELUs per launch, 12 FMA-pipe instructions per pair -- needs ~47 us.  Here nothing is staged through
// registers and every role has its own warps:
//   producer (1 lane)   activation tiles [32 ch][128 px] bf16 by TMA (tensor map, 128B swizzle) into a 4-slot ring;
//   MMA issue (1 lane)  D[128 px][128 hidden] = X W1^T, kind::f16: x is exact in bf16, W1 enters as three bf16 pieces
//                       (24 significant bits), 6 MMAs; + 3 MMAs "ones x b1" that add the bias; 4 accumulators = the
//                       whole tensor memory, slot s of the x ring <-> accumulator s, ONE ready barrier per tile
//                       (TMA bytes + the 4 epilogue warps that drained the accumulator);
//   epilogue (16 warps) four groups of four warps (one warp per TMEM lane quadrant); group g owns slot g, i.e. tiles
//                       T = g (mod 4).  A thread owns a pixel and all 128 hidden units of it: 4 x tcgen05.ld, 64 GELU
//                       pairs, fc2 as two FFMA2 per pair, mask, two coalesced stores -- no cross-warp reduction, fixed
//                       summation order (deterministic).
// The (B,128,64,64) hidden tensor of the reference (537 MB at B=256) exists only in tensor memory.
#include "fno_common.cuh"
#include "tc_common.cuh"
#include "tc_tma.cuh"

namespace fno {

constexpr int kPwEpiWarps = 16;
constexpr int kPwMmaWarp = 16, kPwProdWarp = 17;
constexpr int kPwThreads = 18 * 32;
constexpr int kPwR = 4;                       // ring slots = accumulators (4 x 128 columns = all of tensor memory)
constexpr int kPwM = 128;                     // pixels per tile
constexpr int kPwTilesPerSample = kHW / kPwM;
constexpr uint32_t kPwXBytes = 8192;
constexpr uint32_t kPwW1Piece = kProj * kC * 2;      // 8 KB: bf16 [n = hidden 128][k = channel 32], K-major
constexpr uint32_t kPwB1Piece = kProj * 16 * 2;      // 4 KB: bf16 [n = hidden 128][k = 16], only k = 0 non-zero
constexpr uint32_t kPwOnes = kPwM * 16 * 2;          // 4 KB: bf16 [m = 128][k = 16], column 0 = 1

struct PwSmem {
  alignas(1024) unsigned char x[kPwR][kPwXBytes];
  alignas(1024) unsigned char w1[3][kPwW1Piece];
  alignas(1024) unsigned char b1[3][kPwB1Piece];
  alignas(1024) unsigned char ones[kPwOnes];
  alignas(16) float4 w2q[kProj / 2];   // (w2[0][j], w2[1][j], w2[0][j+1], w2[1][j+1])
  alignas(8) uint64_t ready[kPwR], d_full[kPwR], x_free[kPwR];
  uint32_t tmem_base;
};

// three bf16 pieces of an fp32 value: v = p0 + p1 + p2 up to 2^-24 |v|
__device__ __forceinline__ void pw_split3(float v, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
  p0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(p0);
  p1 = __float2bfloat16_rn(r1);
  p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
}
// byte offset of element (row, k) in a K-major un-swizzled bf16 operand with `rows` rows (8 x 16-byte core matrices)
__host__ __device__ constexpr uint32_t pw_kmajor16(int row, int k, int rows) {
  return static_cast<uint32_t>(((k >> 3) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

__global__ void __launch_bounds__(kPwThreads, 1)
    project_ws_kernel(const __grid_constant__ CUtensorMap x_map, const float* __restrict__ w1, const float* __restrict__ b1,
                      const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ mask,
                      float* __restrict__ preds, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  PwSmem& sm = *reinterpret_cast<PwSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;

  // ---------------------------------------------------------------- prologue (weights only)
  if (tid == 0) {
    for (int i = 0; i < kPwR; ++i) {
      mbar_init(&sm.ready[i], 1 + 4);   // producer's expect_tx arrival + the 4 epilogue warps of the slot's group
      mbar_init(&sm.d_full[i], 1);
      mbar_init(&sm.x_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == kPwMmaWarp) tc::tmem_alloc<512>(&sm.tmem_base);
  // all loads of a thread first (one L2 round trip instead of eight dependent ones), then the conversions
  constexpr int kW1PerThread = (kProj * kC + kPwThreads - 1) / kPwThreads;
  float w1v[kW1PerThread];
#pragma unroll
  for (int it = 0; it < kW1PerThread; ++it) {
    const int e = tid + it * kPwThreads;
    w1v[it] = e < kProj * kC ? __ldg(w1 + e) : 0.f;
  }
#pragma unroll
  for (int it = 0; it < kW1PerThread; ++it) {   // w1[j][i] -> B[n = j][k = i]
    const int e = tid + it * kPwThreads;
    if (e >= kProj * kC) break;
    const int j = e / kC, i = e % kC;
    __nv_bfloat16 p0, p1, p2;
    pw_split3(w1v[it], p0, p1, p2);
    const uint32_t off = pw_kmajor16(j, i, kProj);
    *reinterpret_cast<__nv_bfloat16*>(sm.w1[0] + off) = p0;
    *reinterpret_cast<__nv_bfloat16*>(sm.w1[1] + off) = p1;
    *reinterpret_cast<__nv_bfloat16*>(sm.w1[2] + off) = p2;
  }
  for (int e = tid; e < kProj * 16; e += kPwThreads) {   // bias operand: B[n = j][k] = b1[j] pieces at k = 0
    const int j = e >> 4, k = e & 15;
    __nv_bfloat16 p0 = __float2bfloat16_rn(0.f), p1 = p0, p2 = p0;
    if (k == 0) pw_split3(b1[j], p0, p1, p2);
    const uint32_t off = pw_kmajor16(j, k, kProj);
    *reinterpret_cast<__nv_bfloat16*>(sm.b1[0] + off) = p0;
    *reinterpret_cast<__nv_bfloat16*>(sm.b1[1] + off) = p1;
    *reinterpret_cast<__nv_bfloat16*>(sm.b1[2] + off) = p2;
  }
  for (int e = tid; e < kPwM * 16; e += kPwThreads)
    *reinterpret_cast<__nv_bfloat16*>(sm.ones + pw_kmajor16(e >> 4, e & 15, kPwM)) = __float2bfloat16_rn((e & 15) == 0 ? 1.f : 0.f);
  if (tid < kProj / 2) sm.w2q[tid] = make_float4(w2[2 * tid], w2[kProj + 2 * tid], w2[2 * tid + 1], w2[kProj + 2 * tid + 1]);
  const float b2x = b2[0], b2y = b2[1];
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem = sm.tmem_base;
  pdl_wait();   // the activations come from the previous kernel of the chain
  pdl_launch_dependents();
  // all accumulators start out free: the epilogue groups' share of every ready barrier's first phase
  if (warp < kPwEpiWarps && lane == 0) mbar_arrive(&sm.ready[warp >> 2]);

  // ================================================================ epilogue
  if (warp < kPwEpiWarps) {
    const int q = warp & 3, g = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    for (int it = g; it < n_mine; it += kPwR) {
      const int tile = first + it * stride;
      const int b = tile / kPwTilesPerSample, pix = (tile % kPwTilesPerSample) * kPwM + q * 32 + lane;
      mbar_wait(&sm.d_full[g], (it / kPwR) & 1);
      tc::fence_after_thread_sync();
      float2 acc = make_float2(0.f, 0.f);   // (out channel 0, out channel 1)
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        float v[32];
        tc::tmem_ld32(tmem + g * kProj + chunk * 32 + lane_base, v);
        if (chunk == 3) {   // the whole accumulator is in registers (or consumed): the slot may be refilled
          tc::fence_before_thread_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.ready[g]);
        }
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {   // 8 pairs at a time: 8 independent polynomial chains in flight
          float2 gl[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) gl[i] = make_float2(v[c0 + 2 * i], v[c0 + 2 * i + 1]);   // the accumulator includes b1
          gelu_erf2_deg5_batch<8>(gl);   // degree-5 erfc fit, fno_common.cuh
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 wq = sm.w2q[(chunk * 32 + c0 + 2 * i) >> 1];
            acc = __ffma2_rn(make_float2(gl[i].x, gl[i].x), make_float2(wq.x, wq.y), acc);
            acc = __ffma2_rn(make_float2(gl[i].y, gl[i].y), make_float2(wq.z, wq.w), acc);
          }
        }
      }
      const float mk = __ldg(mask + static_cast<size_t>(b) * kHW + pix);
      preds[(static_cast<size_t>(b) * 2 + 0) * kHW + pix] = (b2x + acc.x) * mk;
      preds[(static_cast<size_t>(b) * 2 + 1) * kHW + pix] = (b2y + acc.y) * mk;
    }
  }
  // ================================================================ MMA issue
  else if (warp == kPwMmaWarp) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc_x = fz_idesc_bf16(kPwM, kProj) | kAMajorMN;   // A = x tile, MN-major (pixels contiguous)
      constexpr uint32_t idesc_b = fz_idesc_bf16(kPwM, kProj);               // A = ones, K-major
      const uint32_t ones_s = tc::smem_addr(sm.ones);
#pragma unroll 1
      for (int it = 0; it < n_mine; ++it) {
        const int s = it % kPwR;
        mbar_wait(&sm.ready[s], (it / kPwR) & 1);   // x tile landed and the accumulator is drained
        tc::fence_after_thread_sync();
        const uint32_t d = tmem + s * kProj;
        const uint32_t x_s = tc::smem_addr(sm.x[s]);
#pragma unroll
        for (int pc = 0; pc < 3; ++pc) {
          const uint32_t w_s = tc::smem_addr(sm.w1[pc]);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)   // K = 32 channels = 2 x 16; B's K-direction core-matrix stride is 2048 B
            fz_mma_f16_ss(d, fz_desc_sw128(x_s + ks * 2048, 4096, 1024), tc::make_smem_desc(w_s + ks * 4096, 2048, 128),
                          idesc_x, (pc | ks) ? 1u : 0u);
          fz_mma_f16_ss(d, tc::make_smem_desc(ones_s, 2048, 128), tc::make_smem_desc(tc::smem_addr(sm.b1[pc]), 2048, 128),
                        idesc_b, 1u);
        }
        tc::mma_commit(&sm.x_free[s]);
        tc::mma_commit(&sm.d_full[s]);
      }
    }
    __syncwarp();
  }
  // ================================================================ producer
  else if (warp == kPwProdWarp) {
    if (lane == 0) {
      for (int it = 0; it < n_mine; ++it) {
        const int s = it % kPwR;
        const int tile = first + it * stride;
        const int b = tile / kPwTilesPerSample, px0 = (tile % kPwTilesPerSample) * kPwM;
        if (it >= kPwR) mbar_wait(&sm.x_free[s], ((it - kPwR) / kPwR) & 1);   // the previous user's MMAs have read the slot
        mbar_expect_tx(&sm.ready[s], kPwXBytes);
        fz_tma_load_2d(sm.x[s], &x_map, px0, b * kC, &sm.ready[s]);
        fz_tma_load_2d(sm.x[s] + 4096, &x_map, px0 + 64, b * kC, &sm.ready[s]);
      }
    }
    __syncwarp();
  }

  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == kPwMmaWarp) tc::tmem_dealloc<512>(tmem);
}

cudaError_t launch_project_ws(const void* a_bf16, const float* w1, const float* b1, const float* w2, const float* b2,
                              const float* mask, float* preds, int batch, cudaStream_t stream) {
  static bool configured[64] = {};
  static int n_sm[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!configured[dev]) {
    e = cudaFuncSetAttribute(project_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PwSmem));
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  if (reinterpret_cast<uintptr_t>(a_bf16) & 15) return cudaErrorMisalignedAddress;
  CUtensorMap map;
  e = fz_make_map(a_bf16, batch, &map);
  if (e != cudaSuccess) return e;
  const int n_tiles = batch * kPwTilesPerSample;
  const int grid = n_tiles < n_sm[dev] ? n_tiles : n_sm[dev];
  return launch_chained(project_ws_kernel, dim3(grid), dim3(kPwThreads), sizeof(PwSmem), stream, map, w1, b1, w2, b2, mask,
                        preds, n_tiles);
}

}  // namespace fno
