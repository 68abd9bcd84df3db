// tcgen05 / TMEM primitives (sm_100a) used by the tensor-core kernels: raw PTX wrappers and the
// shared-memory / instruction descriptor encodings for kind::tf32 with un-swizzled K-major operands.
//
// Operand layout ("canonical K-major, SWIZZLE_NONE"): the matrix is tiled in core matrices of
// 8 rows (M or N) x 16 bytes (4 tf32 along K), each stored as 128 contiguous bytes (row r at r*16 B).
// SBO = byte distance between core matrices adjacent along M/N, LBO = along K.  One MMA consumes
// K = 8 (two core matrices, LBO apart); advancing K by 8 adds 2*LBO to the descriptor start address.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fno {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- TMEM allocation (one full warp executes these) -------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, Blackwell version field = 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);              // bits [0,14)  start address >> 4
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;     // bits [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;     // bits [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                              // bits [46,48) descriptor version (sm_100)
  return d;                                                         // base_offset 0, lbo_mode 0, layout NONE
}
// 32-bit instruction descriptor: D f32, A/B tf32, both K-major, dense, no negate.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
  return (1u << 4)                 // c_format = F32
         | (2u << 7)               // a_format = TF32
         | (2u << 10)              // b_format = TF32
         | (static_cast<uint32_t>(n >> 3) << 17)   // n_dim
         | (static_cast<uint32_t>(m >> 4) << 24);  // m_dim
}

// Warp index broadcast from lane 0: tells the compiler the value is warp-uniform, so branches on it are uniform
// branches and the code under them may use the uniform datapath (UR address math, ULDC descriptors) instead of
// re-deriving every uniform operand with R2UR (14 % of block_tc's executed instructions before this).
__device__ __forceinline__ int warp_index_uniform() {
  return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
}

// One elected lane of a converged warp.  Code guarded by this (and fed with warp-uniform values) is compiled
// to the uniform datapath: back-to-back UTCHMMA with UR operands instead of an R2UR + ELECT loop per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// compile-time accumulate flag variants (keep the predicate an immediate)
template <bool kAccumulate>
__device__ __forceinline__ void mma_tf32_imm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  if constexpr (kAccumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
  }
}

// Same with the A operand in TENSOR MEMORY (row m in lane m, 8 consecutive 32-bit columns per K = 8 step; M = 64 uses
// the same 16-lanes-per-quadrant placement as the accumulator, tools/tc_probe4.cu).  An MMA whose two operands
// come from shared memory is bound by their fetch -- (M + N) * 32 B at ~128 B/clk: 45.6 cycles for M=128, N=32 --
// whereas this form fetches only B: 22.7 cycles (tools/tc_latency.cu, tools/tc_probe3.cu).
template <bool kAccumulate>
__device__ __forceinline__ void mma_tf32_tmem_a_imm(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  if constexpr (kAccumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc)
        : "memory");
  }
}
// registers -> TMEM: 32 lanes x 16 consecutive 32-bit columns per warp (its own lane quadrant)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
               : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 consecutive fp32 columns per warp ------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand staging helpers -------------------------------------------------------------------------
// byte offset of element (row, k) inside a K-major un-swizzled operand tile with `rows` rows:
// core matrix (row/8, k/4) at ((k/4) * (rows/8) + row/8) * 128
__host__ __device__ constexpr uint32_t kmajor_offset(int row, int k, int rows) {
  return static_cast<uint32_t>(((k >> 2) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 3) * 4);
}
// hi/lo split for 3xTF32.  The tensor core TRUNCATES the low 13 mantissa bits of what it reads, which
// would bias every product; so both parts are rounded to nearest tf32 here (cvt.rna) and the hardware
// truncation is then a no-op: hi = rna(x), lo = rna(x - hi), |x - hi - lo| <= 2^-22 |x|, unbiased.
// Round-to-nearest (ties away from zero, like cvt.rna.tf32.f32) as two integer ops on the bit pattern: add half an
// ulp of the 10-bit mantissa, clear the 13 low bits.  ptxas expands cvt.rna into ~5 instructions because it also
// preserves NaN/Inf payloads; activations and weights here are finite, and Inf/NaN still map to Inf/NaN (the
// exponent field is untouched unless the mantissa carry overflows a value within 2^-11 of FLT_MAX).
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}

}  // namespace tc
}  // namespace fno
