// K1 on the tensor cores (bf16 activation storage) -- truncated forward 2-D DFT of activation planes:
//     x[b][c][64][64] (bf16)  ->  Xm[k][b][c]  (288 kept modes, complex64, mode-major).
//
// Replaces torch.fft.rfft2 + the two corner slices of the reference (src/models/fno/fno2d.py:62,73-78).
// Both 1-D transforms are GEMMs against constant twiddle matrices, on a batch of 4 planes:
//
//  stage A (along w):  G[(p,h)][(q,ri)] = sum_w x_p[h][w] * TA[(q,ri)][w],  q = 0..11,  TA = (cos, -sin)(2 pi q w/64)
//      tcgen05.mma kind::f16, M = 128 (2 planes x 64 rows), N = 32 (24 used), K = 64.  The A operand IS the bf16 plane:
//      its 16-byte row chunks are dropped into the K-major core-matrix layout by cp.async (no register pass, exact);
//      TA is split into three bf16 terms (t1 + t2 + t3 carries 24 mantissa bits), so 3 x 4 MMAs per plane pair.
//  stage B (along h):  F[(kxi,part)][(p,q)] = sum_{(ri,h)} A2[(kxi,part)][(ri,h)] * G_p[h][q][ri],  kxi = 0..23
//      tcgen05.mma kind::tf32 as 3xTF32, M = 64 (48 used), N = 48 (4 planes x 12), K = 128.  A2 = (c, s | -s, c) is a
//      constant; the B operand is stage A's accumulator, read from TMEM by the thread that owns row (p,h), split into
//      tf32 hi/lo and scattered K-major (4-byte stores, conflict-free through a skewed K stride).
//      Because x is real, G[h][-q] = conj(G[h][q]): only q >= 0 is computed and stage B produces all 24 kept kx rows
//      (kx = 0..11 and 52..63) of the 12 kept columns directly.
//  epilogue: rows (kxi,re) / (kxi,im) sit in adjacent TMEM lanes (M = 64 places row r in lane 32 (r/16) + r%16,
//      tools/tc_probe2.cu); lane pairs exchange halves with two shuffles per column and write 16 bytes each.
//
// One persistent 512-thread CTA per SM; per batch the stage-B MMAs of batch i and the stage-A MMAs of batch i+1 run
// while the threads fetch batch i+2 (cp.async) and write batch i-1's modes.  Thread work per plane is ~400 warp
// instructions against ~2250 for the register-FFT kernel (fno_dft_fwd.cu, still used for fp32 storage and for the
// fp32 gradients of the backward pass).
#include "fno_common.cuh"
#include "tc_common.cuh"
#include <math.h>
#include <stddef.h>
#include <string.h>

namespace fno {

constexpr int kTdWorkers = 512;                            // 16 worker warps: fetch, split, epilogue
constexpr int kTdThreads = kTdWorkers + 32;                // + one warp that only issues MMAs
constexpr int kTdPlanes = 4;                               // planes per batch
constexpr uint32_t kTdLboX = (128 / 8) * 128 + 16;         // 2064: A operand of stage A (2 planes), skewed K stride
constexpr uint32_t kTdXGroupBytes = 8 * kTdLboX;           // 16,512 B per plane pair (64 bf16 = 8 K chunks)
constexpr uint32_t kTdXBufBytes = 2 * kTdXGroupBytes;      // 33,024 B per batch
constexpr int kTdNA = 32;                                  // stage A N (24 used)
// Rows 24..31 of TA (and rows 48..63 of A2 below) only pad N (M) up to what the instruction accepts; their results are
// never read.  The K stride is therefore that of the USED rows, so a padded row group aliases the start of the next K
// column (finite table values) instead of costing shared memory.
constexpr uint32_t kTdLboTA = (24 / 8) * 128;              // 384
constexpr uint32_t kTdTABytes = 8 * kTdLboTA + 128;        // 3,200 B per bf16 term
constexpr int kTdM2 = 64, kTdK2 = 128, kTdN2 = kTdPlanes * kM2;  // stage B: 64 x 48 x 128, column n2 = 4 q + p
constexpr uint32_t kTdLboA2 = (48 / 8) * 128;              // 768
constexpr uint32_t kTdA2Bytes = (kTdK2 / 4) * kTdLboA2 + 256;  // 24,832 B per image
constexpr uint32_t kTdLboB2 = (kTdN2 / 8) * 128 + 16;      // 784: skewed so that lanes running along h do not collide
constexpr uint32_t kTdB2Bytes = (kTdK2 / 4) * kTdLboB2;    // 25,088 B per image
constexpr uint32_t kTdTableBytes = 3 * kTdTABytes + 2 * kTdA2Bytes;  // 59,264 B constant block (one bulk copy)

struct TdSmem {
  alignas(128) unsigned char x[2][kTdXBufBytes];   // stage-A A operands, double buffered
  alignas(128) unsigned char ta[3][kTdTABytes];    // stage-A B operand: bf16 terms t1, t2, t3       } one contiguous
  alignas(128) unsigned char a2[2][kTdA2Bytes];    // stage-B A operand: tf32 hi, lo                 } table image
  alignas(128) unsigned char b2[2][2][kTdB2Bytes]; // stage-B B operand: [buffer][tf32 hi, lo]
  alignas(8) uint64_t mma_a_bar[2];
  alignas(8) uint64_t mma_b_bar[2];
  alignas(8) uint64_t ready_bar[2];   // workers -> MMA warp: planes landed / stage-B operand staged (16 warp arrivals)
  alignas(8) uint64_t table_bar;
  uint32_t tmem_base;
};
static_assert(offsetof(TdSmem, a2) == offsetof(TdSmem, ta) + 3 * kTdTABytes, "table image must be contiguous");

__host__ __device__ constexpr uint32_t td_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
template <bool kAccumulate>
__device__ __forceinline__ void td_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  if constexpr (kAccumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
  } else {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
  }
}
// 32 lanes x 4 consecutive 32-bit columns
__device__ __forceinline__ void td_tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r0);
  v[1] = __uint_as_float(r1);
  v[2] = __uint_as_float(r2);
  v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void td_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte chunks of 4 planes -> the two K-major A operands of stage A.  Thread = (row h, chunk kc) of every plane:
// a warp reads 512 contiguous bytes per plane; all addresses are a per-thread constant plus compile-time offsets.
__device__ __forceinline__ void td_fetch(uint32_t dst_thread, const __nv_bfloat16* __restrict__ src_thread) {
#pragma unroll
  for (int p = 0; p < kTdPlanes; ++p) {
    const uint32_t dst = dst_thread + (p >> 1) * kTdXGroupBytes + (p & 1) * (64 / 8) * 128;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src_thread + p * kHW) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kTdThreads, 1)
    dft_fwd_tc_kernel(const __nv_bfloat16* __restrict__ x, float2* __restrict__ xm, const unsigned char* __restrict__ table,
                      int n_batches, int batch, float s0, float s1) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TdSmem& sm = *reinterpret_cast<TdSmem*>(smem_raw);
  if ((smem_u32(smem_raw) & 127u) != 0) __trap();
  const int tid = threadIdx.x, lane = tid & 31, warp = tc::warp_index_uniform();

  if (tid == 0) {
    mbar_init(&sm.mma_a_bar[0], 1);
    mbar_init(&sm.mma_a_bar[1], 1);
    mbar_init(&sm.mma_b_bar[0], 1);
    mbar_init(&sm.mma_b_bar[1], 1);
    mbar_init(&sm.ready_bar[0], kTdWorkers / 32);
    mbar_init(&sm.ready_bar[1], kTdWorkers / 32);
    mbar_init(&sm.table_bar, 1);
    fence_mbar_init();
    mbar_expect_tx(&sm.table_bar, kTdTableBytes);
    bulk_g2s(sm.ta, table, kTdTableBytes, &sm.table_bar);
  }
  if (warp == 0) tc::tmem_alloc<256>(&sm.tmem_base);
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_a = sm.tmem_base;          // stage-A accumulators: [buffer][group] x 32 columns
  const uint32_t tmem_b = sm.tmem_base + 128;    // stage-B accumulators: [buffer] x 64 columns (48 used)
  pdl_wait();  // the table is a constant; x comes from the previous kernel of the chain
  pdl_launch_dependents();

  const int first = blockIdx.x, stride = gridDim.x;
  const int n_mine = (first < n_batches) ? (n_batches - first + stride - 1) / stride : 0;
  auto plane0_of = [&](int i) { return static_cast<size_t>(first + i * stride) * kTdPlanes; };

  // fetch constants of this thread: row h = tid >> 3, chunk kc = tid & 7
  const int f_h = tid >> 3, f_kc = tid & 7;
  const uint32_t f_dst = f_kc * kTdLboX + (f_h >> 3) * 128 + (f_h & 7) * 16;
  const __nv_bfloat16* f_src = x + f_h * kW + f_kc * 8;
  const uint32_t x_s0 = smem_u32(sm.x[0]), x_s1 = smem_u32(sm.x[1]);

  // stage-A MMAs of local batch i (planes in x[i & 1]) -> tmem_a buffer i & 1; called by one elected thread
  auto issue_stage_a = [&](int i) {
    constexpr uint32_t idesc = td_idesc_bf16(128, kTdNA);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const uint32_t d_tmem = tmem_a + (i & 1) * 64 + g * 32;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint64_t da0 = tc::make_smem_desc(((i & 1) ? x_s1 : x_s0) + g * kTdXGroupBytes, kTdLboX, 128);
        const uint64_t db0 = tc::make_smem_desc(smem_u32(sm.ta[t]), kTdLboTA, 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // K = 16 per MMA = two 8-element core matrices
          const uint64_t da = da0 + ((ks * 2 * kTdLboX) >> 4), db = db0 + ((ks * 2 * kTdLboTA) >> 4);
          if (t == 0 && ks == 0) td_mma_bf16<false>(d_tmem, da, db, idesc);
          else td_mma_bf16<true>(d_tmem, da, db, idesc);
        }
      }
    }
    tc::mma_commit(&sm.mma_a_bar[i & 1]);
  };
  auto issue_stage_b = [&](int i) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(kTdM2, kTdN2);
    const uint32_t d_tmem = tmem_b + (i & 1) * 64;
    const uint32_t a_s[3] = {smem_u32(sm.a2[0]), smem_u32(sm.a2[1]), smem_u32(sm.a2[0])};
    const uint32_t b_hi = smem_u32(sm.b2[0][0]) + (i & 1) * 2 * kTdB2Bytes;
    const uint32_t b_s[3] = {b_hi, b_hi, b_hi + kTdB2Bytes};
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t da0 = tc::make_smem_desc(a_s[pass], kTdLboA2, 128);
      const uint64_t db0 = tc::make_smem_desc(b_s[pass], kTdLboB2, 128);
#pragma unroll
      for (int ks = 0; ks < kTdK2 / 8; ++ks) {
        const uint64_t da = da0 + ((ks * 2 * kTdLboA2) >> 4), db = db0 + ((ks * 2 * kTdLboB2) >> 4);
        if (pass == 0 && ks == 0) tc::mma_tf32_imm<false>(d_tmem, da, db, idesc);
        else tc::mma_tf32_imm<true>(d_tmem, da, db, idesc);
      }
    }
    tc::mma_commit(&sm.mma_b_bar[i & 1]);
  };

  // split-phase constants: warp = (colhalf, group g, lane quadrant); thread = row (p, h) of stage A's result and the
  // 12 columns (q = 6 colhalf + j/2, ri = j & 1) of it.  B2 element (n2 = 4 q + p, k2 = 64 ri + h): the byte offset is
  // a per-thread constant plus a compile-time function of j.
  const int sp_quad = warp & 3, sp_g = (warp >> 2) & 1, sp_half = warp >> 3;
  const int sp_p = 2 * sp_g + (sp_quad >> 1), sp_h = (sp_quad & 1) * 32 + lane;
  const uint32_t sp_tmem = tmem_a + sp_g * 32 + sp_half * 12 + (static_cast<uint32_t>(sp_quad * 32) << 16);
  unsigned char* sp_b2 = sm.b2[0][0] + (sp_h >> 2) * kTdLboB2 + (3 * sp_half) * 128 + sp_p * 16 + (sp_h & 3) * 4;

  // modes of a batch: TMEM lanes hold rows m2 = 2 kxi + part; the four warps that share a lane quadrant take three
  // columns q each (12 consecutive accumulator columns n2 = 4 q + p)
  const int ep_quad = warp & 3, ep_q0 = 3 * (warp >> 2);
  const int ep_m2 = ep_quad * 16 + (lane & 15), ep_kxi = ep_m2 >> 1, ep_part = ep_m2 & 1;
  const uint32_t ep_tmem = tmem_b + 4 * ep_q0 + (static_cast<uint32_t>(ep_quad * 32) << 16);
  auto epilogue_load = [&](int i, float* v) {  // issues the TMEM loads only; the caller waits
    if (ep_quad == 3) return;                   // rows 48..63 of the M = 64 accumulator are padding
    const uint32_t taddr = ep_tmem + (i & 1) * 64;
    td_tmem_ld4(taddr, v);
    td_tmem_ld4(taddr + 4, v + 4);
    td_tmem_ld4(taddr + 8, v + 8);
  };
  auto epilogue_store = [&](int i, const float* v) {
    if (ep_quad == 3) return;
    const size_t plane0 = plane0_of(i);
    const int b = static_cast<int>(plane0 / kC), c0 = static_cast<int>(plane0 % kC);
    float2* dst = xm + (static_cast<size_t>(ep_kxi * kM2 + ep_q0) * batch + b) * kC + c0 + 2 * ep_part;
    const size_t q_stride = static_cast<size_t>(batch) * kC;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float s = (ep_q0 + j == 0) ? s0 : s1;
      const float w0 = v[4 * j] * s, w1 = v[4 * j + 1] * s, w2 = v[4 * j + 2] * s, w3 = v[4 * j + 3] * s;
      // even lane (re row) keeps planes 0,1 and needs their im; odd lane (im row) keeps planes 2,3 and needs their re
      const float got0 = __shfl_xor_sync(0xffffffffu, ep_part ? w0 : w2, 1);
      const float got1 = __shfl_xor_sync(0xffffffffu, ep_part ? w1 : w3, 1);
      const float4 o = ep_part ? make_float4(got0, w2, got1, w3)    // (re2, im2, re3, im3)
                               : make_float4(w0, got0, w1, got1);   // (re0, im0, re1, im1)
      if (lane < 16) *reinterpret_cast<float4*>(dst + j * q_stride) = o;
    }
  };

  // The MMA instructions of a batch take ~2.7k cycles to ISSUE (the issuing thread is held while the tensor core
  // fetches each instruction's shared-memory operands, tools/tc_latency.cu), so they get a warp of their own: with
  // the issue inside a worker warp every other warp waited for it at the next CTA barrier.  Hand-offs are mbarriers:
  //   ready event e_i (ready_bar[i & 1], one arrival per worker warp): planes of batch i have landed and, for i >= 1,
  //   the stage-B operand of batch i-1 is staged;   mma_a_bar / mma_b_bar: tcgen05.commit of stage A / stage B.
  // Tensor queue order: A(0) | A(1) B(0) | A(2) B(1) | ...
  mbar_wait(&sm.table_bar, 0);
  if (warp == kTdWorkers / 32) {
    // ------------------------------------------------------------------------------------------ MMA warp
    for (int i = 0; i <= n_mine && n_mine > 0; ++i) {
      mbar_wait(&sm.ready_bar[i & 1], (i >> 1) & 1);
      tc::fence_after_thread_sync();
      if (tc::elect_one()) {
        if (i < n_mine) issue_stage_a(i);
        if (i >= 1) issue_stage_b(i - 1);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------------------------------ worker warps
    auto signal_ready = [&](int e) {  // this warp's smem writes (cp.async landed, STS) -> visible to the tensor core
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.ready_bar[e & 1]);
    };
    if (n_mine > 0) td_fetch(x_s0 + f_dst, f_src + plane0_of(0) * kHW);
    if (n_mine > 1) td_fetch(x_s1 + f_dst, f_src + plane0_of(1) * kHW);
    if (n_mine > 0) {
      if (n_mine > 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      signal_ready(0);
    }
    for (int i = 0; i < n_mine; ++i) {
      // stage A of this batch complete: its accumulator is readable and the plane buffer x[i & 1] is free
      mbar_wait(&sm.mma_a_bar[i & 1], (i >> 1) & 1);
      tc::fence_after_thread_sync();
      if (i + 2 < n_mine) td_fetch(((i & 1) ? x_s1 : x_s0) + f_dst, f_src + plane0_of(i + 2) * kHW);
      // G of this thread's row (p, h): split and scatter into the stage-B operand of this batch (buffer i & 1 was
      // last read by stage B of batch i-2, whose completion the epilogue wait of the previous iteration covered)
      {
        float g[12];
        const uint32_t taddr = sp_tmem + (i & 1) * 64;
        td_tmem_ld4(taddr, g);
        td_tmem_ld4(taddr + 4, g + 4);
        td_tmem_ld4(taddr + 8, g + 8);
        td_tmem_wait_ld();
        unsigned char* hi_p = sp_b2 + (i & 1) * 2 * kTdB2Bytes;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          float hi, lo;
          tc::split_tf32(g[j], hi, lo);
          constexpr uint32_t kRiStep = 16 * kTdLboB2;  // k2 += 64
          const uint32_t off = (j & 1) * kRiStep + (j >> 2) * 128 + ((j >> 1) & 1) * 64;
          *reinterpret_cast<float*>(hi_p + off) = hi;
          *reinterpret_cast<float*>(hi_p + kTdB2Bytes + off) = lo;
        }
      }
      // planes of batch i+1 have landed (everything but the fetch issued above)
      if (i + 2 < n_mine) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      signal_ready(i + 1);
      if (i >= 1) {  // modes of the previous batch (its stage B was queued one iteration ago)
        mbar_wait(&sm.mma_b_bar[(i - 1) & 1], ((i - 1) >> 1) & 1);
        tc::fence_after_thread_sync();
        float f[12];
        epilogue_load(i - 1, f);
        td_tmem_wait_ld();
        tc::fence_before_thread_sync();
        epilogue_store(i - 1, f);
      }
    }
    if (n_mine >= 1) {
      mbar_wait(&sm.mma_b_bar[(n_mine - 1) & 1], ((n_mine - 1) >> 1) & 1);
      tc::fence_after_thread_sync();
      float f[12];
      epilogue_load(n_mine - 1, f);
      td_tmem_wait_ld();
      epilogue_store(n_mine - 1, f);
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(sm.tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Constant block: TA as three bf16 terms [n = 2q + ri][w] (rows 24..31 zero), then A2 hi / lo
// [m2 = 2 kxi + part][k2 = ri*64 + h] (rows 48..63 zero), all in K-major core-matrix order.  Built once per
// device from float64.
// ------------------------------------------------------------------------------------------------
static uint16_t td_bf16_bits(double v) {  // round to nearest even
  float f = static_cast<float>(v);
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(r >> 16);
}
static double td_bf16_value(uint16_t b) {
  const uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return static_cast<double>(f);
}
static float td_round_tf32(double v) {
  float f = static_cast<float>(v);
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&f, &u, 4);
  return f;
}

static unsigned char* g_td_table[64] = {nullptr};

static cudaError_t td_ensure_table(const unsigned char** out, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (g_td_table[dev] == nullptr) {
    static unsigned char host[kTdTableBytes];
    memset(host, 0, sizeof(host));
    const double two_pi = 2.0 * 3.14159265358979323846;
    for (int q = 0; q < kM2; ++q)
      for (int ri = 0; ri < 2; ++ri)
        for (int w = 0; w < 64; ++w) {
          const double ang = two_pi * ((q * w) % 64) / 64.0;
          double rest = ri ? -sin(ang) : cos(ang);
          const int n = 2 * q + ri;
          const size_t off = static_cast<size_t>(w >> 3) * kTdLboTA + (n >> 3) * 128 + (n & 7) * 16 + (w & 7) * 2;  // n < 24
          for (int t = 0; t < 3; ++t) {
            const uint16_t bits = td_bf16_bits(rest);
            memcpy(host + t * kTdTABytes + off, &bits, 2);
            rest -= td_bf16_value(bits);
          }
        }
    unsigned char* a2 = host + 3 * kTdTABytes;
    for (int kxi = 0; kxi < kKX; ++kxi) {
      const int kx = kxi < kM1 ? kxi : kxi + (kH - kKX);
      for (int part = 0; part < 2; ++part)
        for (int ri = 0; ri < 2; ++ri)
          for (int h = 0; h < 64; ++h) {
            const double ang = two_pi * ((kx * h) % 64) / 64.0;
            const double c = cos(ang), s = sin(ang);
            // Fre = sum c Gre + s Gim;  Fim = sum -s Gre + c Gim
            const double val = part == 0 ? (ri == 0 ? c : s) : (ri == 0 ? -s : c);
            const int m2 = 2 * kxi + part, k2 = ri * 64 + h;
            const size_t off = static_cast<size_t>(k2 >> 2) * kTdLboA2 + (m2 >> 3) * 128 + (m2 & 7) * 16 + (k2 & 3) * 4;
            const float hi = td_round_tf32(val);
            const float lo = td_round_tf32(val - static_cast<double>(hi));
            memcpy(a2 + off, &hi, 4);
            memcpy(a2 + kTdA2Bytes + off, &lo, 4);
          }
    }
    unsigned char* d = nullptr;
    e = cudaMalloc(&d, sizeof(host));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d, host, sizeof(host), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);  // `host` is static: make sure the copy has consumed it
    if (e != cudaSuccess) return e;
    g_td_table[dev] = d;
  }
  *out = g_td_table[dev];
  return cudaSuccess;
}

void dft_fwd_tc_release(int dev) {
  if (dev >= 0 && dev < 64 && g_td_table[dev] != nullptr) {
    cudaFree(g_td_table[dev]);
    g_td_table[dev] = nullptr;
  }
}

cudaError_t launch_dft_fwd_tc(const void* x, void* xm, int batch, float s0, float s1, cudaStream_t stream) {
  auto kern = dft_fwd_tc_kernel;
  constexpr size_t smem = sizeof(TdSmem);
  static PerDeviceLaunch pd;
  int n_sm = 0;
  cudaError_t e0 = per_device_setup(kern, smem, pd, &n_sm);
  if (e0 != cudaSuccess) return e0;
  const unsigned char* table = nullptr;
  cudaError_t e = td_ensure_table(&table, stream);
  if (e != cudaSuccess) return e;
  const int n_batches = batch * kC / kTdPlanes;
  const int grid = n_batches < n_sm ? n_batches : n_sm;
  return launch_chained(kern, dim3(grid), dim3(kTdThreads), smem, stream, static_cast<const __nv_bfloat16*>(x),
                        static_cast<float2*>(xm), table, n_batches, batch, s0, s1);
}

}  // namespace fno
