// K3 (tensor-core version) -- zero-padded inverse transform + 1x1 conv + bias + exact GELU, fused:
//   out[b][o][h][w] = act( irfft2(pad(Y))[b][o][h][w] + sum_i W0[o][i] x[b][i][h][w] + bias[o] )
// replacing irfft2 + Conv2d(32,32,1) + add + GELU of the reference FnoBlock
// (src/models/fno/fno2d.py:81,104-111).
//
// Persistent kernel, one CTA per SM, work item = (sample b, row residue r): the 8 rows h = 8h'+r.
//  phase A (CUDA cores)  inverse DFT along kx evaluated at the item's 8 rows (codelets
//                        icfft64_in24_r<r>), result Z[h'][ky][o] split into tf32 hi/lo and written
//                        straight into the B-operand layout of the tensor core.
//  per tile of 2 rows (128 pixels = UMMA M):
//     TMA bulk copies bring x[b][:, h, :] (32 ch x 2 rows) into smem; a split pass rewrites it as the
//     K-major A operand (hi/lo tf32); one thread issues tcgen05.mma (kind::tf32, 3xTF32 split):
//        D[128 px][32 o] = [E(+)E | X] * [Z_a ; Z_b ; W0^T]        K = 24 + 24 + 32
//     where E[w][(ky,re/im)] = (cos, -sin)(2 pi ky w/64) is the C2R stage of the inverse transform as a
//     constant matrix (its ky=0 imaginary column is zero: that is irfft2 dropping Im of the DC column),
//     block-diagonal over the tile's two rows.  The accumulator lives in TMEM (32 columns per tile,
//     4 tiles in flight); the epilogue reads it back with tcgen05.ld (thread = pixel), adds bias,
//     applies GELU (or the backward epilogues) and stores coalesced along w.
//  The MMAs of tile t+1 overlap the epilogue of tile t; the TMA of tile t+2 overlaps both.
#include "fft_codelets.cuh"
#include "fno_common.cuh"
#include "tc_common.cuh"
#include <math.h>
#include <string.h>

namespace fno {

enum : int { kEpiGelu = 0, kEpiGeluSavePre = 1, kEpiMulDgelu = 2, kEpiPlain = 3 };

constexpr int kTcThreads = 256;
constexpr int kTcRows = 8;                 // rows per work item (== residues of the kx codelets)
constexpr int kTcTiles = kTcRows / 2;      // 2 rows = 128 pixels per MMA tile
constexpr int kTcM = 128;
constexpr int kKE = 48;                    // E-part K: 2 rows x 12 ky x (re, im)
constexpr int kKConv = 32;                    // conv-part K: input channels
constexpr uint32_t kLboA = (kTcM / 8) * 128;   // 2048: K-direction core-matrix stride of a 128-row operand
constexpr uint32_t kLboB = (kC / 8) * 128;     // 512 : ... of a 32-row operand
constexpr int kETabFloats = 2 * kTcM * kKE;    // hi image then lo image

template <typename TAct>
struct TcSmem {
  alignas(128) float e_hi[kTcM * kKE];             // A operand, E part (constant)       24,576 B
  alignas(128) float e_lo[kTcM * kKE];
  alignas(128) float ax_hi[2][kTcM * kKConv];         // A operand, conv part, 2 stages     2 x 16,384 B
  alignas(128) float ax_lo[2][kTcM * kKConv];         // (unused for bf16 activations: they are tf32-exact)
  alignas(128) float zb_hi[kTcTiles][kC * kKE];    // B operand, E part, one per tile    4 x 6,144 B
  alignas(128) float zb_lo[kTcTiles][kC * kKE];
  alignas(128) float wb_hi[kC * kKConv];              // B operand, conv part               4,096 B
  alignas(128) float wb_lo[kC * kKConv];
  alignas(128) TAct raw[2][2][kC][kW];             // TMA landing zone: [stage][row][ch][w]
  alignas(16) float bias[kC];
  alignas(8) uint64_t raw_bar[2];
  alignas(8) uint64_t mma_bar[kTcTiles];
  uint32_t tmem_base;
};

template <int R>
__device__ __forceinline__ void inv_kx_tc(const float* yre, const float* yim, float* ore, float* oim) {
  if constexpr (R == 0) fno_codelets::icfft64_in24_r0<float>(yre, yim, ore, oim);
  if constexpr (R == 1) fno_codelets::icfft64_in24_r1<float>(yre, yim, ore, oim);
  if constexpr (R == 2) fno_codelets::icfft64_in24_r2<float>(yre, yim, ore, oim);
  if constexpr (R == 3) fno_codelets::icfft64_in24_r3<float>(yre, yim, ore, oim);
  if constexpr (R == 4) fno_codelets::icfft64_in24_r4<float>(yre, yim, ore, oim);
  if constexpr (R == 5) fno_codelets::icfft64_in24_r5<float>(yre, yim, ore, oim);
  if constexpr (R == 6) fno_codelets::icfft64_in24_r6<float>(yre, yim, ore, oim);
  if constexpr (R == 7) fno_codelets::icfft64_in24_r7<float>(yre, yim, ore, oim);
}

__device__ __forceinline__ float act_to_float(float v) { return v; }
__device__ __forceinline__ float act_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store_act(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_act(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename TAct, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1)
    block_out_tc_kernel(const float2* __restrict__ ym, const TAct* __restrict__ x, const float* __restrict__ w0t,
                        const float* __restrict__ bias, const float* __restrict__ etab, TAct* __restrict__ out,
                        float* __restrict__ pre_out, const float* __restrict__ pre_in, float s0, float s1,
                        int n_items) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];  // no pointer arithmetic: keeps LDS/STS addressing
  TcSmem<TAct>& sm = *reinterpret_cast<TcSmem<TAct>*>(smem_raw);
  if ((smem_u32(smem_raw) & 127u) != 0) __trap();
  constexpr bool kBf16 = sizeof(TAct) == 2;
  constexpr uint32_t kRowBytes = kW * sizeof(TAct);
  constexpr uint32_t kTileBytes = 2 * kC * kRowBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    mbar_init(&sm.raw_bar[0], 1);
    mbar_init(&sm.raw_bar[1], 1);
#pragma unroll
    for (int t = 0; t < kTcTiles; ++t) mbar_init(&sm.mma_bar[t], 1);
    fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc<kTcTiles * kC>(&sm.tmem_base);
  for (int e = tid; e < kTcM * kKE; e += kTcThreads) {
    sm.e_hi[e] = etab[e];
    sm.e_lo[e] = etab[kTcM * kKE + e];
  }
  for (int e = tid; e < kC * kKConv; e += kTcThreads) {  // B[n = o][k = i] = W0[o][i] = w0t[i][o]
    const int i = e / kC, o = e % kC;
    float hi, lo;
    tc::split_tf32(w0t[e], hi, lo);
    const uint32_t off = tc::kmajor_offset(o, i, kC) / 4;
    sm.wb_hi[off] = hi;
    sm.wb_lo[off] = lo;
  }
  if (tid < kC) sm.bias[tid] = (bias != nullptr) ? bias[tid] : 0.f;
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = sm.tmem_base;

  const int first = blockIdx.x;
  const int n_mine = (first < n_items) ? (n_items - first + gridDim.x - 1) / gridDim.x : 0;
  const int n_tiles_total = n_mine * kTcTiles;

  // TMA of global tile index q (= item_iter * 4 + t) into raw stage q & 1, issued by one whole warp
  auto issue_tile_load = [&](int q) {
    const int item = first + (q / kTcTiles) * gridDim.x;
    const int b = item >> 3, r = item & 7, t = q % kTcTiles;
    const int s = q & 1;
    if (lane == 0) mbar_expect_tx(&sm.raw_bar[s], kTileBytes);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int h = 8 * (2 * t + j) + r;
      bulk_g2s(&sm.raw[s][j][lane][0], x + ((static_cast<size_t>(b) * kC + lane) * kH + h) * kW, kRowBytes,
               &sm.raw_bar[s]);
    }
  };
  if (warp == 1) {
    if (n_tiles_total > 0) issue_tile_load(0);
    if (n_tiles_total > 1) issue_tile_load(1);
  }

  constexpr uint32_t idesc = tc::make_idesc_tf32(kTcM, kC);

  // epilogue of tile t of the item (b, r): TMEM -> registers -> global.  Warps w and w+4 share TMEM lane
  // quadrant w & 3 (pixels 32(w&3)..+31 of the tile) and take output channels 0..15 / 16..31.
  auto epilogue = [&](int b, int r, int t, uint32_t parity) {
    mbar_wait(&sm.mma_bar[t], parity);
    tc::fence_after_thread_sync();
    const int quad = warp & 3, half = warp >> 2;
    const int m = quad * 32 + lane;            // pixel within the tile
    const int j = m >> 6, w = m & 63;
    const int h = 8 * (2 * t + j) + r;
    float v[16];
    {
      uint32_t rr[16];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * kC + half * 16;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
            "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(rr[c]);
    }
    const size_t base = ((static_cast<size_t>(b) * kC + half * 16) * kH + h) * kW + w;
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
      float2 p = make_float2(v[c], v[c + 1]);
      const size_t o0 = base + static_cast<size_t>(c) * kHW, o1 = o0 + kHW;
      if constexpr (EPI == kEpiGelu || EPI == kEpiGeluSavePre) {
        p.x += sm.bias[half * 16 + c];
        p.y += sm.bias[half * 16 + c + 1];
        if constexpr (EPI == kEpiGeluSavePre) {
          pre_out[o0] = p.x;
          pre_out[o1] = p.y;
        }
        p = gelu_erf2(p);
      } else if constexpr (EPI == kEpiMulDgelu) {
        p.x *= dgelu_erf(__ldg(pre_in + o0));
        p.y *= dgelu_erf(__ldg(pre_in + o1));
      }
      store_act(out + o0, p.x);
      store_act(out + o1, p.y);
    }
    tc::fence_before_thread_sync();
  };

  // ---------------------------------------------------------------- persistent loop over work items
  for (int it = 0; it < n_mine; ++it) {
    const int item = first + it * gridDim.x;
    const int b = item >> 3, r = item & 7;
    const uint32_t item_parity = it & 1;

    // ---- phase A: inverse along kx at rows 8h'+r -> B operand (E part) of the item's 4 tiles
    {
      const int o = lane;
      const float2* ym_b = ym + static_cast<size_t>(b) * kModes * kC;
#pragma unroll 1
      for (int ky = warp; ky < kM2; ky += kTcThreads / 32) {
        float yre[24], yim[24], ore[8], oim[8];
#pragma unroll
        for (int kxi = 0; kxi < 24; ++kxi) {
          const float2 v = __ldg(ym_b + (kxi * kM2 + ky) * kC + o);
          yre[kxi] = v.x;
          yim[kxi] = v.y;
        }
        switch (r) {
          case 0: inv_kx_tc<0>(yre, yim, ore, oim); break;
          case 1: inv_kx_tc<1>(yre, yim, ore, oim); break;
          case 2: inv_kx_tc<2>(yre, yim, ore, oim); break;
          case 3: inv_kx_tc<3>(yre, yim, ore, oim); break;
          case 4: inv_kx_tc<4>(yre, yim, ore, oim); break;
          case 5: inv_kx_tc<5>(yre, yim, ore, oim); break;
          case 6: inv_kx_tc<6>(yre, yim, ore, oim); break;
          default: inv_kx_tc<7>(yre, yim, ore, oim); break;
        }
        const float s = (ky == 0) ? s0 : s1;
#pragma unroll
        for (int hp = 0; hp < 8; ++hp) {
          const int t = hp >> 1, j = hp & 1;
          float rh, rl, ih, il;
          tc::split_tf32(ore[hp] * s, rh, rl);
          tc::split_tf32(oim[hp] * s, ih, il);
          const uint32_t off = tc::kmajor_offset(o, j * 24 + 2 * ky, kC) / 4;  // (re, im) are k-adjacent
          *reinterpret_cast<float2*>(&sm.zb_hi[t][off]) = make_float2(rh, ih);
          *reinterpret_cast<float2*>(&sm.zb_lo[t][off]) = make_float2(rl, il);
        }
      }
    }

    // ---- tiles
    for (int t = 0; t < kTcTiles; ++t) {
      const int q = it * kTcTiles + t;
      const int s = q & 1;
      mbar_wait(&sm.raw_bar[s], (q >> 1) & 1);
      // split pass: raw[s][j][i][w] -> A operand rows m = 64 j + w, k = i  (K-major, hi/lo)
#pragma unroll
      for (int rep = 0; rep < (kTcM * (kKConv / 4)) / kTcThreads; ++rep) {
        const int task = rep * kTcThreads + tid;
        const int m = task & (kTcM - 1), iq = task >> 7;
        const int j = m >> 6, w = m & 63;
        float hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float xv = act_to_float(sm.raw[s][j][4 * iq + c][w]);
          if constexpr (kBf16) {
            hi[c] = xv;  // bf16 is exactly representable in tf32
            lo[c] = 0.f;
          } else {
            tc::split_tf32(xv, hi[c], lo[c]);
          }
        }
        const uint32_t off = tc::kmajor_offset(m, 4 * iq, kTcM) / 4;
        *reinterpret_cast<float4*>(&sm.ax_hi[s][off]) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        if constexpr (!kBf16) *reinterpret_cast<float4*>(&sm.ax_lo[s][off]) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      tc::fence_proxy_async_smem();
      tc::fence_before_thread_sync();
      __syncthreads();  // operands of tile t complete; raw[s] free; epilogue(t-1) TMEM reads (prev iteration) done
      tc::fence_after_thread_sync();

      if (warp == 1 && q + 2 < n_tiles_total) issue_tile_load(q + 2);
      if (warp == 0) {
        if (tc::elect_one()) {
          // 3xTF32: pass 0 = hi*hi, pass 1 = lo*hi, pass 2 = hi*lo (A part, B part).  Everything below is
          // warp-uniform and unrolled, so the descriptors live in uniform registers.
          const uint32_t d_tmem = tmem_base + t * kC;
          const uint32_t a_e[3] = {tc::smem_addr(sm.e_hi), tc::smem_addr(sm.e_lo), tc::smem_addr(sm.e_hi)};
          const uint32_t b_z[3] = {tc::smem_addr(sm.zb_hi[t]), tc::smem_addr(sm.zb_hi[t]), tc::smem_addr(sm.zb_lo[t])};
          const uint32_t a_x[3] = {tc::smem_addr(sm.ax_hi[s]), tc::smem_addr(sm.ax_lo[s]), tc::smem_addr(sm.ax_hi[s])};
          const uint32_t b_w[3] = {tc::smem_addr(sm.wb_hi), tc::smem_addr(sm.wb_hi), tc::smem_addr(sm.wb_lo)};
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t da0 = tc::make_smem_desc(a_e[pass], kLboA, 128);
            const uint64_t db0 = tc::make_smem_desc(b_z[pass], kLboB, 128);
#pragma unroll
            for (int ks = 0; ks < kKE / 8; ++ks) {
              const uint64_t da = da0 + ((ks * 2 * kLboA) >> 4), db = db0 + ((ks * 2 * kLboB) >> 4);
              if (pass == 0 && ks == 0) tc::mma_tf32_imm<false>(d_tmem, da, db, idesc);
              else tc::mma_tf32_imm<true>(d_tmem, da, db, idesc);
            }
            if (kBf16 && pass == 1) continue;  // conv-part A has no lo component
            const uint64_t dx0 = tc::make_smem_desc(a_x[pass], kLboA, 128);
            const uint64_t dw0 = tc::make_smem_desc(b_w[pass], kLboB, 128);
#pragma unroll
            for (int ks = 0; ks < kKConv / 8; ++ks)
              tc::mma_tf32_imm<true>(d_tmem, dx0 + ((ks * 2 * kLboA) >> 4), dw0 + ((ks * 2 * kLboB) >> 4), idesc);
          }
          tc::mma_commit(&sm.mma_bar[t]);
        }
        __syncwarp();
      }
      if (t > 0) epilogue(b, r, t - 1, item_parity);
    }
    epilogue(b, r, kTcTiles - 1, item_parity);
    tc::fence_before_thread_sync();
    __syncthreads();  // all TMEM reads and zb/ax reads of this item are done before the next item rewrites them
    tc::fence_after_thread_sync();
  }

  if (warp == 0) tc::tmem_dealloc<kTcTiles * kC>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Constant A-operand image of the C2R stage: rows m = 64 j + w (j = row of the tile), columns
// k = 24 j' + 2 ky + ri;  E = cos(2 pi ky w/64) (ri=0), -sin(2 pi ky w/64) (ri=1, zero for ky=0), zero for
// j != j'.  Built once per device in float64, split into tf32 hi/lo (round-to-nearest), laid out K-major.
// ------------------------------------------------------------------------------------------------
static float round_tf32_host(double x) {
  float f = static_cast<float>(x);
  uint32_t u;
  memcpy(&u, &f, 4);
  u = (u + 0x1000u) & 0xffffe000u;  // round half away from zero on the magnitude (cvt.rna)
  memcpy(&f, &u, 4);
  return f;
}

static float* g_etab[64] = {nullptr};

cudaError_t ensure_etab(const float** out, cudaStream_t stream) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (g_etab[dev] == nullptr) {
    static float host[kETabFloats];
    for (int i = 0; i < kETabFloats; ++i) host[i] = 0.f;
    for (int m = 0; m < kTcM; ++m) {
      const int j = m >> 6, w = m & 63;
      for (int ky = 0; ky < kM2; ++ky) {
        const double ang = 2.0 * 3.14159265358979323846 * ((ky * w) % 64) / 64.0;
        const double val[2] = {cos(ang), ky == 0 ? 0.0 : -sin(ang)};
        for (int ri = 0; ri < 2; ++ri) {
          const int k = 24 * j + 2 * ky + ri;
          const float hi = round_tf32_host(val[ri]);
          const float lo = round_tf32_host(val[ri] - static_cast<double>(hi));
          const uint32_t off = tc::kmajor_offset(m, k, kTcM) / 4;
          host[off] = hi;
          host[kTcM * kKE + off] = lo;
        }
      }
    }
    float* d = nullptr;
    e = cudaMalloc(&d, sizeof(host));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(d, host, sizeof(host), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(stream);  // `host` is static: make sure the copy has consumed it
    if (e != cudaSuccess) return e;
    g_etab[dev] = d;
  }
  *out = g_etab[dev];
  return cudaSuccess;
}

template <typename TAct, int EPI>
static cudaError_t launch_one_tc(const void* ym, const void* x, const float* w0t, const float* bias, void* out,
                                 float* pre_out, const float* pre_in, int batch, float s0, float s1,
                                 cudaStream_t stream) {
  auto kern = block_out_tc_kernel<TAct, EPI>;
  constexpr size_t smem = sizeof(TcSmem<TAct>) + 128;
  static bool configured = false;
  static int n_sm = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const float* etab = nullptr;
  cudaError_t e = ensure_etab(&etab, stream);
  if (e != cudaSuccess) return e;
  const int n_items = batch * kTcRows;
  const int grid = n_items < n_sm ? n_items : n_sm;
  kern<<<grid, kTcThreads, smem, stream>>>(static_cast<const float2*>(ym), static_cast<const TAct*>(x), w0t, bias, etab,
                                           static_cast<TAct*>(out), pre_out, pre_in, s0, s1, n_items);
  return cudaGetLastError();
}

template <typename TAct>
cudaError_t launch_block_out_tc(int epi, const void* ym, const void* x, const float* w0t, const float* bias,
                                void* out, float* pre_out, const float* pre_in, int batch, float s0, float s1,
                                cudaStream_t stream) {
  switch (epi) {
    case kEpiGelu: return launch_one_tc<TAct, kEpiGelu>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    case kEpiGeluSavePre:
      return launch_one_tc<TAct, kEpiGeluSavePre>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    case kEpiMulDgelu:
      return launch_one_tc<TAct, kEpiMulDgelu>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    case kEpiPlain: return launch_one_tc<TAct, kEpiPlain>(ym, x, w0t, bias, out, pre_out, pre_in, batch, s0, s1, stream);
    default: return cudaErrorInvalidValue;
  }
}

template cudaError_t launch_block_out_tc<float>(int, const void*, const void*, const float*, const float*, void*,
                                                float*, const float*, int, float, float, cudaStream_t);
template cudaError_t launch_block_out_tc<__nv_bfloat16>(int, const void*, const void*, const float*, const float*,
                                                        void*, float*, const float*, int, float, float, cudaStream_t);

}  // namespace fno
