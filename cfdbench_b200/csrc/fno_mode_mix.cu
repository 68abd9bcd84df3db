// K2 -- per-mode complex channel mix on the tensor cores:  Y[k][b][o] = sum_i X[k][b][i] * Wk[k][i][o].
//
// Replaces the two torch.einsum("bixy,ioxy->boxy") corner products plus the zero-filled
// (B,32,64,33) cfloat buffer of the reference (src/models/fno/fno2d.py:54-57, 65-78).
//
// Modes are stored mode-major (xm[k][b][c], written that way by dft_fwd_kernel), so the 128 rows of a tile are one
// contiguous 32 KB block; with the sample-major layout the same rows sat 73,728 B apart and both this kernel and its
// CUDA-core predecessor were bound by that access pattern (~20 us per launch at B=256 whatever the arithmetic).
// For one mode k the mix over a tile of 128 samples is a real GEMM on the interleaved complex64 rows exactly
// as they sit in memory:
//     D[128 samples][64 = (o, re|im)] = A[128][64 = (i, re|im)] * B_k^T         (tcgen05.mma kind::tf32, K = 64)
//     B_k[(o,re)][(i,re)] = Wre,  B_k[(o,re)][(i,im)] = -Wim,  B_k[(o,im)][(i,re)] = Wim,  B_k[(o,im)][(i,im)] = Wre
// run as 3xTF32 (hi*hi + lo*hi + hi*lo).  The B operand of every mode is prepared once per weight update
// (pack_mix_operand_kernel: real-expanded, split into tf32 hi/lo, laid out as the K-major UMMA image) and arrives
// with ONE 32 KB bulk copy per tile; the activations-side rows are prefetched into registers one tile ahead
// (two 256-byte rows per warp instruction), split and stored as the A operand.  The A operand's K-direction core
// matrix stride (LBO) is skewed by 16 B so that the 16-byte stores of lanes running along K are bank-conflict free.
// Persistent CTA = two independent 256-thread pipelines, accumulators double-buffered in TMEM; the epilogue writes
// 128 contiguous bytes per thread.  With the conj-transposed pack the same kernel is the adjoint mix of the
// backward pass (Xbar[b,i,k] = sum_o G[b,o,k] conj(W[i,o,k])).
//
// (The CUDA-core version of this phase -- weights in registers, FFMA2 -- is in the history up to commit 73b34cc:
// 20.7 us per launch at B=256 against the ~6 us its 40 MB of HBM traffic need.)
#include "fno_common.cuh"
#include "tc_common.cuh"

namespace fno {

constexpr int kMxThreads = 512;  // two independent 256-thread tile pipelines
constexpr int kMxGroup = 256;
constexpr int kMxM = 128;        // samples per tile
constexpr int kMxK = 2 * kC;     // 64 real (i, re|im)
constexpr int kMxN = 2 * kC;     // 64 real (o, re|im)
constexpr uint32_t kMxLboA = (kMxM / 8) * 128 + 16;  // 2064: skewed K-direction core-matrix stride of A
constexpr uint32_t kMxLboB = (kMxN / 8) * 128;       // 1024
constexpr int kMxAFloats = (kMxK / 4) * kMxLboA / 4; // 8256
constexpr int kMxBFloats = kMxN * kMxK;              // 4096 per image (hi or lo)
constexpr int kMxOperandFloats = 2 * kMxBFloats;     // per mode: hi image then lo image (32 KB)
constexpr int kMxReps = (kMxM * kMxK / 4) / kMxGroup;  // 8 float4 per thread per tile

struct MxSmem {
  alignas(128) float a_hi[2][kMxAFloats];       // [pipeline] 2 x 33,024 B
  alignas(128) float a_lo[2][kMxAFloats];
  alignas(128) float b[2][kMxOperandFloats];    // [pipeline] 2 x 32 KB, bulk-copied per tile
  alignas(8) uint64_t mma_bar[2][2];
  alignas(8) uint64_t b_bar[2];
  uint32_t tmem_base;
};

// float index of A element (row m, column kk) with kk a multiple of 4
__device__ __forceinline__ uint32_t mx_a_offset(int m, int kk) {
  return (static_cast<uint32_t>(kk >> 2) * kMxLboA + static_cast<uint32_t>(m >> 3) * 128u + static_cast<uint32_t>(m & 7) * 16u) >> 2;
}

struct MxRegs {
  float4 v[kMxReps];  // task = rep*256 + gtid -> (row m = task >> 4, 16-byte chunk = task & 15)
};

__device__ __forceinline__ void mx_prefetch(MxRegs& r, const float4* __restrict__ xm, int k, int b0, int batch, int gtid) {
#pragma unroll
  for (int rep = 0; rep < kMxReps; ++rep) {
    const int task = rep * kMxGroup + gtid;
    const int m = task >> 4, ch = task & 15;
    r.v[rep] = (b0 + m < batch) ? __ldg(xm + (static_cast<size_t>(k) * batch + b0 + m) * (kC / 2) + ch)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__device__ __forceinline__ void mx_split_store(const MxRegs& r, float* a_hi, float* a_lo, int gtid) {
#pragma unroll
  for (int rep = 0; rep < kMxReps; ++rep) {
    const int task = rep * kMxGroup + gtid;
    const int m = task >> 4, ch = task & 15;
    float4 hi, lo;
    tc::split_tf32(r.v[rep].x, hi.x, lo.x);
    tc::split_tf32(r.v[rep].y, hi.y, lo.y);
    tc::split_tf32(r.v[rep].z, hi.z, lo.z);
    tc::split_tf32(r.v[rep].w, hi.w, lo.w);
    const uint32_t off = mx_a_offset(m, 4 * ch);
    *reinterpret_cast<float4*>(a_hi + off) = hi;
    *reinterpret_cast<float4*>(a_lo + off) = lo;
  }
}

// 256-bit global store (sm_100: st.global.v8): the image epilogue writes 32-byte chunks of 32 different samples per warp
// instruction, so the instruction count -- not the bytes -- is what the LSU queue sees (lg_throttle 5.0 per issue with
// 16-byte stores, profiles/ncu_r02*.md).
__device__ __forceinline__ void mx_store32(void* dst, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

template <int GRP>
__device__ __forceinline__ void mx_group_barrier() {
  asm volatile("bar.sync %0, %1;" ::"n"(GRP + 1), "n"(kMxGroup) : "memory");
}

template <int GRP>
__device__ __forceinline__ void mx_pipeline(MxSmem& sm, const float4* __restrict__ xm, const float* __restrict__ wop,
                                            float4* __restrict__ ym, unsigned char* __restrict__ ym_img, int batch,
                                            int n_btiles, int n_tiles) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int gtid = tid & (kMxGroup - 1), gwarp = tc::warp_index_uniform() & 7;
  const uint32_t tmem_base = sm.tmem_base + GRP * (2 * kMxN);
  constexpr uint32_t idesc = tc::make_idesc_tf32(kMxM, kMxN);

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_cta = (first < n_tiles) ? (n_tiles - first + stride - 1) / stride : 0;
  const int n_mine = (n_cta + 1 - GRP) / 2;
  auto tile_of = [&](int it) { return first + (2 * it + GRP) * stride; };  // tile = k * n_btiles + sample tile

  // epilogue of local tile `it`: warps w and w+4 share TMEM lane quadrant w & 3 (rows 32(w&3)..+31 of the tile) and
  // take the 32-float column halves; a thread writes 128 contiguous bytes of its sample's output row
  auto epilogue = [&](int it) {
    const int buf = it & 1;
    mbar_wait(&sm.mma_bar[GRP][buf], (it >> 1) & 1);
    tc::fence_after_thread_sync();
    const int quad = gwarp & 3, half = gwarp >> 2;
    const int tile = tile_of(it);
    const int k = tile / n_btiles, b = (tile % n_btiles) * kMxM + quad * 32 + lane;
    float v[32];
    tc::tmem_ld32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kMxN + half * 32, v);
    tc::fence_before_thread_sync();
    if (b < batch && ym_img == nullptr) {
      float4* dst = ym + (static_cast<size_t>(k) * batch + b) * (kC / 2) + half * 8;
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    } else if (b < batch) {
      // Operand image of block_fused_kernel's GEMM1 (fno_block_fused.cu): per sample [hi|lo][ky/4][ky%4][48 rows][32 o]
      // fp32, row = 24 (kxi & 1) + 2 (kxi >> 1) + (re|im), 32-byte chunks XOR-swizzled with (row & 3); tf32 hi / lo
      // split here so the consumer is pure bulk copy + MMA.  This thread holds o = 16 half .. 16 half + 15, (re, im).
      const int kxi = k / kM2, ky = k % kM2;
      unsigned char* img = ym_img + static_cast<size_t>(b) * 147456 + (ky >> 2) * 24576 + (ky & 3) * 6144;
#pragma unroll
      for (int ri = 0; ri < 2; ++ri) {
        const int row = 24 * (kxi & 1) + 2 * (kxi >> 1) + ri;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) tc::split_tf32(v[2 * (8 * c + j) + ri], hi[j], lo[j]);
          unsigned char* dst = img + row * 128 + (((2 * half + c) ^ (row & 3)) << 5);
          mx_store32(dst, hi);            // one 32-byte chunk = one full sector per store instruction
          mx_store32(dst + 73728, lo);
        }
      }
    }
  };

  // Two tiles of activations are in flight in registers (at B = 256 a pipeline owns only ~2 tiles, so every
  // load of the kernel is issued up front and the DRAM latency is paid once, not once per tile).
  MxRegs ring[2];
  auto prefetch_tile = [&](MxRegs& r, int it) {
    if (it < n_mine) {
      const int t = tile_of(it);
      mx_prefetch(r, xm, t / n_btiles, (t % n_btiles) * kMxM, batch, gtid);
    }
  };
  prefetch_tile(ring[0], 0);
  prefetch_tile(ring[1], 1);

  auto body = [&](int it, MxRegs& regs) {
    const int buf = it & 1;
    const int tile = tile_of(it);
    // the single-buffered operands were last read by the MMAs of tile it-1: wait for them (normally long done)
    if (it >= 1) mbar_wait(&sm.mma_bar[GRP][(it - 1) & 1], ((it - 1) >> 1) & 1);
    if (it >= 1 && gwarp == 0 && tc::elect_one()) {  // this tile's B operand: one bulk copy (tile 0: kernel prologue)
      constexpr uint32_t kBytes = kMxOperandFloats * sizeof(float);
      mbar_expect_tx(&sm.b_bar[GRP], kBytes);
      bulk_g2s(sm.b[GRP], wop + static_cast<size_t>(tile / n_btiles) * kMxOperandFloats, kBytes, &sm.b_bar[GRP]);
    }
    mx_split_store(regs, sm.a_hi[GRP], sm.a_lo[GRP], gtid);
    tc::fence_proxy_async_smem();
    tc::fence_before_thread_sync();
    mx_group_barrier<GRP>();
    tc::fence_after_thread_sync();
    // refill this register set AFTER the fence: the membar inside fence.proxy.async would otherwise wait for the loads
    prefetch_tile(regs, it + 2);
    if (gwarp == 0) {
      if (tc::elect_one()) {
        mbar_wait(&sm.b_bar[GRP], it & 1);
        const uint32_t d_tmem = tmem_base + buf * kMxN;
        const uint32_t a_s[3] = {tc::smem_addr(sm.a_hi[GRP]), tc::smem_addr(sm.a_lo[GRP]), tc::smem_addr(sm.a_hi[GRP])};
        const uint32_t b_hi = tc::smem_addr(sm.b[GRP]), b_lo = b_hi + kMxBFloats * sizeof(float);
        const uint32_t b_s[3] = {b_hi, b_hi, b_lo};
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t da0 = tc::make_smem_desc(a_s[pass], kMxLboA, 128);
          const uint64_t db0 = tc::make_smem_desc(b_s[pass], kMxLboB, 128);
#pragma unroll
          for (int ks = 0; ks < kMxK / 8; ++ks) {
            const uint64_t da = da0 + ((ks * 2 * kMxLboA) >> 4), db = db0 + ((ks * 2 * kMxLboB) >> 4);
            if (pass == 0 && ks == 0) tc::mma_tf32_imm<false>(d_tmem, da, db, idesc);
            else tc::mma_tf32_imm<true>(d_tmem, da, db, idesc);
          }
        }
        tc::mma_commit(&sm.mma_bar[GRP][buf]);
      }
      __syncwarp();
    }
    if (it >= 1) epilogue(it - 1);
  };
  for (int it = 0; it < n_mine; it += 2) {
    body(it, ring[0]);
    if (it + 1 < n_mine) body(it + 1, ring[1]);
  }
  if (n_mine >= 1) epilogue(n_mine - 1);
}

__global__ void __launch_bounds__(kMxThreads, 1)
    mode_mix_tc_kernel(const float4* __restrict__ xm, const float* __restrict__ wop, float4* __restrict__ ym,
                       unsigned char* __restrict__ ym_img, int batch, int n_btiles, int n_tiles) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // no pointer arithmetic: keeps LDS/STS addressing
  MxSmem& sm = *reinterpret_cast<MxSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 127u) != 0) __trap();
  const int tid = threadIdx.x, warp = tc::warp_index_uniform();
  const int grp = warp >> 3;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mbar_init(&sm.mma_bar[i >> 1][i & 1], 1);
    mbar_init(&sm.b_bar[0], 1);
    mbar_init(&sm.b_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<4 * kMxN>(&sm.tmem_base);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  if (tid == 0) {  // the first tile's weights of both pipelines do not depend on the previous kernel: fetch them now
    constexpr uint32_t kBytes = kMxOperandFloats * sizeof(float);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int t = blockIdx.x + g * gridDim.x;
      if (t < n_tiles) {
        mbar_expect_tx(&sm.b_bar[g], kBytes);
        bulk_g2s(sm.b[g], wop + static_cast<size_t>(t / n_btiles) * kMxOperandFloats, kBytes, &sm.b_bar[g]);
      }
    }
  }
  pdl_wait();  // xm comes from the previous kernel of the chain
  pdl_launch_dependents();
  if (grp == 0) mx_pipeline<0>(sm, xm, wop, ym, ym_img, batch, n_btiles, n_tiles);
  else mx_pipeline<1>(sm, xm, wop, ym, ym_img, batch, n_btiles, n_tiles);
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<4 * kMxN>(sm.tmem_base);
}

// ym_img != nullptr: write the per-sample GEMM1 operand image (see the epilogue) instead of the mode-major ym.
cudaError_t launch_mode_mix(const void* xm, const void* wop, void* ym, void* ym_img, int batch, cudaStream_t stream) {
  auto kern = mode_mix_tc_kernel;
  constexpr size_t smem = sizeof(MxSmem);
  static PerDeviceLaunch pd;
  int n_sm = 0;
  cudaError_t e0 = per_device_setup(kern, smem, pd, &n_sm);
  if (e0 != cudaSuccess) return e0;
  const int n_btiles = (batch + kMxM - 1) / kMxM;
  const int n_tiles = kModes * n_btiles;
  const int grid = n_tiles < 2 * n_sm ? (n_tiles + 1) / 2 : n_sm;
  return launch_chained(kern, dim3(grid), dim3(kMxThreads), smem, stream, static_cast<const float4*>(xm),
                        static_cast<const float*>(wop), static_cast<float4*>(ym), static_cast<unsigned char*>(ym_img), batch,
                        n_btiles, n_tiles);
}

// ------------------------------------------------------------------------------------------------
// B operand image of the mix, built from the packed weights Wk[k][i][o] (pack_spectral_kernel below):
// per mode 2 x 4096 floats (tf32 hi image, then lo image), element (n = 2o + part, kk = 2i + ri) at
// tc::kmajor_offset(n, kk, 64).
// ------------------------------------------------------------------------------------------------
__global__ void pack_mix_operand_kernel(const float2* __restrict__ wk, float* __restrict__ wop) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over k*1024 + i*32 + o
  if (idx >= kModes * kC * kC) return;
  const int k = idx / (kC * kC), i = (idx / kC) % kC, o = idx % kC;
  const float2 w = wk[idx];
  float* img = wop + static_cast<size_t>(k) * kMxOperandFloats;
  const float val[2][2] = {{w.x, -w.y}, {w.y, w.x}};  // [part of the output][re|im of the input]
#pragma unroll
  for (int part = 0; part < 2; ++part)
#pragma unroll
    for (int ri = 0; ri < 2; ++ri) {
      float hi, lo;
      tc::split_tf32(val[part][ri], hi, lo);
      const uint32_t off = tc::kmajor_offset(2 * o + part, 2 * i + ri, kMxN) / 4;
      img[off] = hi;
      img[kMxBFloats + off] = lo;
    }
}

cudaError_t launch_pack_mix_operand(const void* wk, void* wop, cudaStream_t stream) {
  const int n = kModes * kC * kC;
  pack_mix_operand_kernel<<<(n + 255) / 256, 256, 0, stream>>>(static_cast<const float2*>(wk), static_cast<float*>(wop));
  return cudaGetLastError();
}

size_t mix_operand_bytes() { return static_cast<size_t>(kModes) * kMxOperandFloats * sizeof(float); }

// The same image straight from the reference parameter layout (weights1/2: (Cin, Cout, 12, 12) complex64), one
// launch per layer and direction: a thread produces one 16-byte operand chunk (row n, 4 consecutive kk) of both the
// hi and the lo image, so the writes are coalesced; the 2.36 MB of weights are gathered through L2.  This is what
// runs after every optimizer step.
__global__ void pack_mix_operand_direct_kernel(const float2* __restrict__ w1, const float2* __restrict__ w2,
                                               float* __restrict__ wop, int conj_transpose) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over k * 1024 + (kk/4) * 64 + n: the operand's own order
  if (idx >= kModes * 1024) return;
  const int k = idx >> 10, kq = (idx >> 6) & 15, n = idx & 63;
  const int kxi = k / kM2, ky = k % kM2;
  const float2* src = (kxi < kM1) ? w1 : w2;
  const int kk_mode = (kxi % kM1) * kM2 + ky;
  const int a = n >> 1, part = n & 1;  // output channel (o) of the mix and its re|im
  float v[4];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = 2 * kq + j;  // input channel (i) of the mix
    // forward: Wk[k][i = c][o = a] = W[c][a][k];  adjoint: Wk[k][i = c][o = a] = conj(W[a][c][k])
    const int wi = conj_transpose ? a : c, wo = conj_transpose ? c : a;
    float2 w = __ldg(src + (static_cast<size_t>(wi) * kC + wo) * (kM1 * kM2) + kk_mode);
    if (conj_transpose) w.y = -w.y;
    v[2 * j + 0] = part ? w.y : w.x;    // (o,re): [Wre, -Wim]   (o,im): [Wim, Wre]
    v[2 * j + 1] = part ? w.x : -w.y;
  }
  float4 hi, lo;
  tc::split_tf32(v[0], hi.x, lo.x);
  tc::split_tf32(v[1], hi.y, lo.y);
  tc::split_tf32(v[2], hi.z, lo.z);
  tc::split_tf32(v[3], hi.w, lo.w);
  float* img = wop + static_cast<size_t>(k) * kMxOperandFloats;
  const uint32_t off = tc::kmajor_offset(n, 4 * kq, kMxN) / 4;
  *reinterpret_cast<float4*>(img + off) = hi;
  *reinterpret_cast<float4*>(img + kMxBFloats + off) = lo;
}

cudaError_t launch_pack_mix_operand_direct(const void* w1, const void* w2, void* wop, int conj_transpose,
                                           cudaStream_t stream) {
  const int n = kModes * 1024;
  pack_mix_operand_direct_kernel<<<(n + 255) / 256, 256, 0, stream>>>(
      static_cast<const float2*>(w1), static_cast<const float2*>(w2), static_cast<float*>(wop), conj_transpose);
  return cudaGetLastError();
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
