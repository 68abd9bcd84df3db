// extern "C" surface of libcfdbench_b200.so (declared in include/cfdbench_b200.h).
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cfdbench_b200.h"
#include "fno_common.cuh"

namespace fno {
template <typename TAct>
cudaError_t launch_dft_fwd(const void*, void*, int, float, float, cudaStream_t);
cudaError_t launch_mode_mix(const void*, const void*, void*, void*, int, cudaStream_t);
cudaError_t launch_block_fused(const void*, const void*, const float*, const float*, void*, int, cudaStream_t);
size_t ym_image_bytes(int);
cudaError_t launch_project_ws(const void*, const float*, const float*, const float*, const float*, const float*, float*, int,
                              cudaStream_t);
cudaError_t launch_pack_spectral(const void*, const void*, void*, int, cudaStream_t);
cudaError_t launch_unpack_spectral(const void*, void*, void*, cudaStream_t);
cudaError_t launch_dft_fwd_tc(const void*, void*, int, float, float, cudaStream_t);
cudaError_t launch_pack_mix_operand(const void*, void*, cudaStream_t);
cudaError_t launch_pack_mix_operand_direct(const void*, const void*, void*, int, cudaStream_t);
size_t mix_operand_bytes();
cudaError_t launch_gather_batch(const void*, const void*, const float*, const int*, const long long*, int, int, int, float*,
                                float*, float*, float*, cudaStream_t);
cudaError_t launch_loss_fwd(const float*, const float*, size_t, float*, float*, cudaStream_t);
cudaError_t launch_loss_bwd(const float*, const float*, const float*, const float*, float*, size_t, cudaStream_t);
size_t loss_scratch_bytes();
cudaError_t launch_adam_step(const fno_adam_tensors*, float, float, float, float, float, long long, cudaStream_t);
cudaError_t launch_inv_kx(const void*, void*, int, float, float, cudaStream_t);
template <typename TAct>
cudaError_t launch_block_tc(int, const void*, const void*, const float*, const float*, void*, float*, const float*, int,
                            cudaStream_t);
template <typename TAct>
cudaError_t launch_lift(const float*, const float*, const float*, const float*, const float*, const float*,
                        const float*, void*, int, int, cudaStream_t);
template <typename TAct>
cudaError_t launch_project_tc(const void*, const float*, const float*, const float*, const float*, const float*,
                              float*, int, cudaStream_t);
template <typename TAct>
cudaError_t launch_project_bwd(const void*, const float*, const float*, const float*, const float*, const float*,
                               const float*, float*, float*, float*, int, cudaStream_t);
int project_bwd_parts(int);
int project_bwd_row();
cudaError_t launch_reduce_partials(const float*, int, int, float*, int, float*, int, float*, int, int, cudaStream_t);
template <typename TAct>
cudaError_t launch_project_bwd_tc(const void*, const float*, const float*, const float*, const float*, const float*, const float*,
                                  float*, float*, float*, int*, int, cudaStream_t);
template <typename TP, typename TQ, int NJ, int NI>
cudaError_t launch_chan_outer(const void*, const void*, float*, int*, int, cudaStream_t);
cudaError_t launch_multistep_metrics(const float*, const float*, const float*, float*, int, int, cudaStream_t);
cudaError_t launch_spectral_wgrad(const void*, const void*, void*, int, cudaStream_t);
cudaError_t launch_lift_bwd(const float*, const float*, const float*, const float*, const float*, const float*,
                            float*, float*, float*, int, int, cudaStream_t);
}  // namespace fno

using namespace fno;

namespace fno {
void block_tc_release(int);
void block_fused_release(int);
void dft_fwd_tc_release(int);
}  // namespace fno
static cudaEvent_t g_chunk_events[64][2][16] = {};   // fno_rollout_host_chunked: [device][upload|compute][chunk]

static thread_local char g_err[512] = "";

static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess)
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  else
    snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}
#define FNO_CUDA(call, what)                         \
  do {                                               \
    cudaError_t _e = (call);                         \
    if (_e != cudaSuccess) return fail(kErrCuda, what, _e); \
  } while (0)
#define FNO_TRY(call)        \
  do {                       \
    int _r = (call);         \
    if (_r != kOk) return _r; \
  } while (0)

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
static inline bool bad_dtype(int d) { return d != FNO_ACT_F32 && d != FNO_ACT_BF16; }

extern "C" {

int fno_version(void) { return FNO_ABI_VERSION; }

int fno_destroy(void) {
  int dev = 0;
  FNO_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
  if (dev < 0 || dev >= 64) return fail(kErrArg, "fno_destroy: device index");
  FNO_CUDA(cudaDeviceSynchronize(), "cudaDeviceSynchronize");
  block_tc_release(dev);
  block_fused_release(dev);
  dft_fwd_tc_release(dev);
  for (int k = 0; k < 2; ++k)
    for (int c = 0; c < 16; ++c)
      if (g_chunk_events[dev][k][c]) {
        cudaEventDestroy(g_chunk_events[dev][k][c]);
        g_chunk_events[dev][k][c] = nullptr;
      }
  return kOk;
}
const char* fno_last_error(void) { return g_err; }

size_t fno_act_bytes(int batch, int act_dtype) {
  return static_cast<size_t>(batch) * kC * kHW * (act_dtype == FNO_ACT_BF16 ? 2 : 4);
}
size_t fno_modes_bytes(int batch) { return static_cast<size_t>(batch) * kModes * kC * sizeof(float2); }
size_t fno_z_bytes(int batch) { return static_cast<size_t>(batch) * kH * 2 * kM2 * kC * sizeof(float); }
size_t fno_ym_image_bytes(int batch) { return ym_image_bytes(batch); }
size_t fno_bwd_partials_bytes(void) {
  return (static_cast<size_t>(296) * (kProj * kC + kProj) + static_cast<size_t>(project_bwd_parts(FNO_BWD_CHUNK)) * project_bwd_row() +
          static_cast<size_t>(16) * kC * (5 + kMaxCaseParams + 1)) * sizeof(float);
}

int fno_pack_spectral_weights(const void* w1, const void* w2, void* wk, int conj_transpose, void* stream) {
  if (!w1 || !w2 || !wk) return fail(kErrArg, "fno_pack_spectral_weights: null pointer");
  FNO_CUDA(launch_pack_spectral(w1, w2, wk, conj_transpose, S(stream)), "pack_spectral_kernel");
  return kOk;
}

size_t fno_mix_operand_bytes(void) { return mix_operand_bytes(); }

int fno_pack_mix_operand(const void* wk, void* wop, void* stream) {
  if (!wk || !wop) return fail(kErrArg, "fno_pack_mix_operand: null pointer");
  FNO_CUDA(launch_pack_mix_operand(wk, wop, S(stream)), "pack_mix_operand_kernel");
  return kOk;
}

int fno_pack_mix_operand_from_weights(const void* w1, const void* w2, void* wop, int conj_transpose, void* stream) {
  if (!w1 || !w2 || !wop) return fail(kErrArg, "fno_pack_mix_operand_from_weights: null pointer");
  FNO_CUDA(launch_pack_mix_operand_direct(w1, w2, wop, conj_transpose, S(stream)), "pack_mix_operand_direct_kernel");
  return kOk;
}

int fno_unpack_spectral_grads(const void* gwk, void* gw1, void* gw2, void* stream) {
  if (!gwk || !gw1 || !gw2) return fail(kErrArg, "fno_unpack_spectral_grads: null pointer");
  FNO_CUDA(launch_unpack_spectral(gwk, gw1, gw2, S(stream)), "unpack_spectral_kernel");
  return kOk;
}

int fno_lift_fwd(const float* inputs, const float* mask, const float* case_params, const fno_weights* w,
                 void* act_out, int batch, int act_dtype, void* stream) {
  if (!inputs || !mask || !w || !act_out || batch <= 0 || bad_dtype(act_dtype))
    return fail(kErrArg, "fno_lift_fwd: bad argument");
  if (w->n_case_params > 0 && !case_params) return fail(kErrArg, "fno_lift_fwd: case_params is null");
  cudaError_t e = act_dtype == FNO_ACT_F32
                      ? launch_lift<float>(inputs, mask, case_params, w->fc0_w, w->fc0_b, w->gx, w->gy, act_out, batch,
                                           w->n_case_params, S(stream))
                      : launch_lift<__nv_bfloat16>(inputs, mask, case_params, w->fc0_w, w->fc0_b, w->gx, w->gy,
                                                   act_out, batch, w->n_case_params, S(stream));
  FNO_CUDA(e, "lift_kernel");
  return kOk;
}

int fno_spectral_dft_fwd(const void* act_in, void* xm, int batch, int act_dtype, float s0, float s1, void* stream) {
  if (!act_in || !xm || batch <= 0 || bad_dtype(act_dtype)) return fail(kErrArg, "fno_spectral_dft_fwd: bad argument");
  if (act_dtype == FNO_ACT_BF16) {
    // bf16 planes: the two-GEMM tensor-core kernel (fno_dft_fwd_tc.cu); FNO_DFT_TC=0 selects the register-FFT kernel
    // (an A/B switch for measurements, read once)
    static const bool use_tc = [] { const char* v = getenv("FNO_DFT_TC"); return !(v && v[0] == '0'); }();
    if (use_tc) {
      FNO_CUDA(launch_dft_fwd_tc(act_in, xm, batch, s0, s1, S(stream)), "dft_fwd_tc_kernel");
      return kOk;
    }
  }
  cudaError_t e = act_dtype == FNO_ACT_F32 ? launch_dft_fwd<float>(act_in, xm, batch, s0, s1, S(stream))
                                           : launch_dft_fwd<__nv_bfloat16>(act_in, xm, batch, s0, s1, S(stream));
  FNO_CUDA(e, "dft_fwd_kernel");
  return kOk;
}

int fno_spectral_dft_fwd_tc(const void* act_in_bf16, void* xm, int batch, float s0, float s1, void* stream) {
  if (!act_in_bf16 || !xm || batch <= 0) return fail(kErrArg, "fno_spectral_dft_fwd_tc: bad argument");
  FNO_CUDA(launch_dft_fwd_tc(act_in_bf16, xm, batch, s0, s1, S(stream)), "dft_fwd_tc_kernel");
  return kOk;
}

int fno_mode_mix(const void* xm, const void* wk, void* ym, int batch, void* stream) {
  if (!xm || !wk || !ym || batch <= 0) return fail(kErrArg, "fno_mode_mix: bad argument");
  FNO_CUDA(launch_mode_mix(xm, wk, ym, nullptr, batch, S(stream)), "mode_mix_tc_kernel");
  return kOk;
}

int fno_mode_mix_image(const void* xm, const void* wk, void* ym_img, int batch, void* stream) {
  if (!xm || !wk || !ym_img || batch <= 0) return fail(kErrArg, "fno_mode_mix_image: bad argument");
  FNO_CUDA(launch_mode_mix(xm, wk, nullptr, ym_img, batch, S(stream)), "mode_mix_tc_kernel(image)");
  return kOk;
}

int fno_block_fused(const void* ym_img, const void* act_in, const float* w0t, const float* bias, void* act_out, int batch,
                    void* stream) {
  if (!ym_img || !act_in || !w0t || !act_out || batch <= 0) return fail(kErrArg, "fno_block_fused: bad argument");
  FNO_CUDA(launch_block_fused(ym_img, act_in, w0t, bias, act_out, batch, S(stream)), "block_fused_kernel");
  return kOk;
}

int fno_spectral_inv_kx(const void* ym, void* z, int batch, float s0, float s1, void* stream) {
  if (!ym || !z || batch <= 0) return fail(kErrArg, "fno_spectral_inv_kx: bad argument");
  FNO_CUDA(launch_inv_kx(ym, z, batch, s0, s1, S(stream)), "inv_kx_kernel");
  return kOk;
}

int fno_block_out(int epilogue, const void* z, const void* act_in, const float* w0t, const float* bias, void* act_out,
                  float* pre_out, const float* pre_in, int batch, int act_dtype, void* stream) {
  if (!z || !act_in || !w0t || !act_out || batch <= 0 || bad_dtype(act_dtype))
    return fail(kErrArg, "fno_block_out: bad argument");
  if (epilogue == FNO_EPI_GELU_SAVE_PRE && !pre_out) return fail(kErrArg, "fno_block_out: pre_out is null");
  if (epilogue == FNO_EPI_MUL_DGELU && !pre_in) return fail(kErrArg, "fno_block_out: pre_in is null");
  cudaError_t e = act_dtype == FNO_ACT_F32
                      ? launch_block_tc<float>(epilogue, z, act_in, w0t, bias, act_out, pre_out, pre_in, batch, S(stream))
                      : launch_block_tc<__nv_bfloat16>(epilogue, z, act_in, w0t, bias, act_out, pre_out, pre_in, batch,
                                                       S(stream));
  FNO_CUDA(e, "block_tc_kernel");
  return kOk;
}

int fno_block_fwd(const fno_weights* w, int layer, const void* act_in, void* act_out, float* pre_out,
                  const fno_workspace* ws, int batch, int act_dtype, void* stream) {
  if (!w || !ws || layer < 0 || layer >= w->n_layers) return fail(kErrArg, "fno_block_fwd: bad argument");
  FNO_TRY(fno_spectral_dft_fwd(act_in, ws->xm, batch, act_dtype, 1.f, 1.f, stream));
  if (act_dtype == FNO_ACT_BF16 && ws->ym_img && !pre_out) {   // inference, bf16 storage: fused output stage
    FNO_TRY(fno_mode_mix_image(ws->xm, w->spec_wk[layer], ws->ym_img, batch, stream));
    return fno_block_fused(ws->ym_img, act_in, w->w0t[layer], w->w0_b[layer], act_out, batch, stream);
  }
  FNO_TRY(fno_mode_mix(ws->xm, w->spec_wk[layer], ws->ym, batch, stream));
  const float inv = 1.f / static_cast<float>(kHW);
  FNO_TRY(fno_spectral_inv_kx(ws->ym, ws->z, batch, inv, 2.f * inv, stream));
  return fno_block_out(pre_out ? FNO_EPI_GELU_SAVE_PRE : FNO_EPI_GELU, ws->z, act_in, w->w0t[layer], w->w0_b[layer],
                       act_out, pre_out, nullptr, batch, act_dtype, stream);
}

int fno_project_fwd(const void* act_in, const float* mask, const fno_weights* w, float* preds, int batch,
                    int act_dtype, void* stream) {
  if (!act_in || !mask || !w || !preds || batch <= 0 || bad_dtype(act_dtype))
    return fail(kErrArg, "fno_project_fwd: bad argument");
  cudaError_t e =
      act_dtype == FNO_ACT_F32
          ? launch_project_tc<float>(act_in, w->fc1_w, w->fc1_b, w->fc2_w, w->fc2_b, mask, preds, batch, S(stream))
          : launch_project_ws(act_in, w->fc1_w, w->fc1_b, w->fc2_w, w->fc2_b, mask, preds, batch, S(stream));
  FNO_CUDA(e, "project_kernel");
  return kOk;
}

int fno_forward(const fno_weights* w, const float* inputs, const float* mask, const float* case_params,
                float* preds, const fno_workspace* ws, int batch, int act_dtype, void* stream) {
  if (!w || !ws || !ws->act[0] || !ws->act[1] || !ws->xm) return fail(kErrArg, "fno_forward: bad workspace");
  if (!(act_dtype == FNO_ACT_BF16 && ws->ym_img) && (!ws->ym || !ws->z)) return fail(kErrArg, "fno_forward: bad workspace");
  if (w->n_layers < 1 || w->n_layers > FNO_MAX_LAYERS) return fail(kErrUnsupported, "fno_forward: n_layers out of range");
  FNO_TRY(fno_lift_fwd(inputs, mask, case_params, w, ws->act[0], batch, act_dtype, stream));
  int cur = 0;
  for (int l = 0; l < w->n_layers; ++l) {
    FNO_TRY(fno_block_fwd(w, l, ws->act[cur], ws->act[cur ^ 1], nullptr, ws, batch, act_dtype, stream));
    cur ^= 1;
  }
  return fno_project_fwd(ws->act[cur], mask, w, preds, batch, act_dtype, stream);
}

int fno_rollout(const fno_weights* w, const float* inputs, const float* mask, const float* case_params,
                float* preds_seq, int steps, const fno_workspace* ws, int batch, int act_dtype, void* stream) {
  if (steps < 0 || !preds_seq) return fail(kErrArg, "fno_rollout: bad argument");
  const size_t frame = static_cast<size_t>(batch) * 2 * kHW;
  const float* cur = inputs;
  for (int s = 0; s < steps; ++s) {
    float* nxt = preds_seq + static_cast<size_t>(s) * frame;
    FNO_TRY(fno_forward(w, cur, mask, case_params, nxt, ws, batch, act_dtype, stream));
    cur = nxt;
  }
  return kOk;
}

size_t fno_rollout_host_scratch_bytes(int batch, int n_case_params, int steps) {
  const size_t b = static_cast<size_t>(batch);
  return (b * 2 * kHW + b * kHW + static_cast<size_t>(steps) * b * 2 * kHW) * sizeof(float) +
         ((b * n_case_params * sizeof(float) + 255) / 256) * 256;
}

int fno_rollout_host(const fno_weights* w, const float* inputs_host, const float* mask_host,
                     const float* case_params_host, float* preds_seq_host, int steps, const fno_workspace* ws,
                     void* dev_io, int batch, int act_dtype, void* stream) {
  if (!w || !inputs_host || !mask_host || !preds_seq_host || !dev_io || batch <= 0 || steps <= 0)
    return fail(kErrArg, "fno_rollout_host: bad argument");
  const size_t b = static_cast<size_t>(batch);
  float* d_in = static_cast<float*>(dev_io);
  float* d_mask = d_in + b * 2 * kHW;
  float* d_seq = d_mask + b * kHW;
  float* d_params = d_seq + static_cast<size_t>(steps) * b * 2 * kHW;
  cudaStream_t st = S(stream);
  FNO_CUDA(cudaMemcpyAsync(d_in, inputs_host, b * 2 * kHW * sizeof(float), cudaMemcpyHostToDevice, st), "H2D inputs");
  FNO_CUDA(cudaMemcpyAsync(d_mask, mask_host, b * kHW * sizeof(float), cudaMemcpyHostToDevice, st), "H2D mask");
  if (w->n_case_params > 0)
    FNO_CUDA(cudaMemcpyAsync(d_params, case_params_host, b * w->n_case_params * sizeof(float), cudaMemcpyHostToDevice, st),
             "H2D case_params");
  FNO_TRY(fno_rollout(w, d_in, d_mask, d_params, d_seq, steps, ws, batch, act_dtype, stream));
  FNO_CUDA(cudaMemcpyAsync(preds_seq_host, d_seq, static_cast<size_t>(steps) * b * 2 * kHW * sizeof(float),
                           cudaMemcpyDeviceToHost, st),
           "D2H preds");
  return kOk;
}

// One step for host buffers as a three-stage pipeline over batch chunks: all host->device copies go, in chunk order,
// through ONE stream, the kernels of the chunks through a second one and the device->host copies through a third,
// chained by events.  Copies of the same direction therefore never run concurrently, while chunk c+1's upload still
// overlaps chunk c's kernels and chunk c-1's download (0.88 ms per step at B = 256 with two chunks, 0.97 ms unchunked).
int fno_rollout_host_chunked(const fno_weights* w, const float* inputs_host, const float* mask_host,
                             const float* case_params_host, float* preds_host, const fno_workspace* ws_chunks,
                             void* const* dev_io_chunks, int batch, int n_chunks, int act_dtype, void* stream_in,
                             void* stream_compute, void* stream_out) {
  constexpr int kMaxChunks = 16;
  if (!w || !inputs_host || !mask_host || !preds_host || !ws_chunks || !dev_io_chunks || batch <= 0 || n_chunks <= 0 ||
      n_chunks > kMaxChunks || batch % n_chunks != 0)
    return fail(kErrArg, "fno_rollout_host_chunked: bad argument");
  auto& ev = g_chunk_events;
  int dev = 0;
  FNO_CUDA(cudaGetDevice(&dev), "cudaGetDevice");
  if (dev < 0 || dev >= 64) return fail(kErrArg, "fno_rollout_host_chunked: device index");
  for (int c = 0; c < n_chunks; ++c)
    for (int k = 0; k < 2; ++k)
      if (!ev[dev][k][c]) FNO_CUDA(cudaEventCreateWithFlags(&ev[dev][k][c], cudaEventDisableTiming), "cudaEventCreate");
  const size_t cb = static_cast<size_t>(batch / n_chunks);
  const int p = w->n_case_params;
  cudaStream_t s_in = S(stream_in), s_cmp = S(stream_compute), s_out = S(stream_out);
  auto d_in = [&](int c) { return static_cast<float*>(dev_io_chunks[c]); };
  auto d_mask = [&](int c) { return d_in(c) + cb * 2 * kHW; };
  auto d_seq = [&](int c) { return d_mask(c) + cb * kHW; };
  auto d_params = [&](int c) { return d_seq(c) + cb * 2 * kHW; };
  for (int c = 0; c < n_chunks; ++c) {
    const size_t lo = c * cb;
    FNO_CUDA(cudaMemcpyAsync(d_in(c), inputs_host + lo * 2 * kHW, cb * 2 * kHW * sizeof(float), cudaMemcpyHostToDevice, s_in),
             "H2D inputs");
    FNO_CUDA(cudaMemcpyAsync(d_mask(c), mask_host + lo * kHW, cb * kHW * sizeof(float), cudaMemcpyHostToDevice, s_in),
             "H2D mask");
    if (p > 0)
      FNO_CUDA(cudaMemcpyAsync(d_params(c), case_params_host + lo * p, cb * p * sizeof(float), cudaMemcpyHostToDevice, s_in),
               "H2D case_params");
    FNO_CUDA(cudaEventRecord(ev[dev][0][c], s_in), "cudaEventRecord");
  }
  for (int c = 0; c < n_chunks; ++c) {
    FNO_CUDA(cudaStreamWaitEvent(s_cmp, ev[dev][0][c], 0), "cudaStreamWaitEvent");
    FNO_TRY(fno_rollout(w, d_in(c), d_mask(c), d_params(c), d_seq(c), 1, &ws_chunks[c], static_cast<int>(cb), act_dtype,
                        stream_compute));
    FNO_CUDA(cudaEventRecord(ev[dev][1][c], s_cmp), "cudaEventRecord");
  }
  for (int c = 0; c < n_chunks; ++c) {
    FNO_CUDA(cudaStreamWaitEvent(s_out, ev[dev][1][c], 0), "cudaStreamWaitEvent");
    FNO_CUDA(cudaMemcpyAsync(preds_host + c * cb * 2 * kHW, d_seq(c), cb * 2 * kHW * sizeof(float), cudaMemcpyDeviceToHost,
                             s_out),
             "D2H preds");
  }
  return kOk;
}

int fno_forward_train(const fno_weights* w, const float* inputs, const float* mask, const float* case_params,
                      float* preds, const fno_train_saved* saved, const fno_workspace* ws, int batch,
                      int act_dtype, void* stream) {
  if (!w || !saved || !ws || !ws->ym || !ws->z) return fail(kErrArg, "fno_forward_train: bad argument");
  if (w->n_layers < 1 || w->n_layers > FNO_MAX_LAYERS) return fail(kErrUnsupported, "fno_forward_train: n_layers");
  FNO_TRY(fno_lift_fwd(inputs, mask, case_params, w, saved->act[0], batch, act_dtype, stream));
  const float inv = 1.f / static_cast<float>(kHW);
  for (int l = 0; l < w->n_layers; ++l) {
    if (!saved->act[l + 1] || !saved->pre[l] || !saved->xm[l]) return fail(kErrArg, "fno_forward_train: null saved buffer");
    FNO_TRY(fno_spectral_dft_fwd(saved->act[l], saved->xm[l], batch, act_dtype, 1.f, 1.f, stream));
    FNO_TRY(fno_mode_mix(saved->xm[l], w->spec_wk[l], ws->ym, batch, stream));
    FNO_TRY(fno_spectral_inv_kx(ws->ym, ws->z, batch, inv, 2.f * inv, stream));
    FNO_TRY(fno_block_out(FNO_EPI_GELU_SAVE_PRE, ws->z, saved->act[l], w->w0t[l], w->w0_b[l], saved->act[l + 1],
                          saved->pre[l], nullptr, batch, act_dtype, stream));
  }
  return fno_project_fwd(saved->act[w->n_layers], mask, w, preds, batch, act_dtype, stream);
}

int fno_backward(const fno_weights* w, const fno_weights_bwd* wb, const float* inputs, const float* mask,
                 const float* case_params, const float* dpreds, const fno_train_saved* saved,
                 const fno_grads* g, const fno_bwd_scratch* sc, const fno_workspace* ws, int batch,
                 int act_dtype, void* stream) {
  return fno_backward_ex(w, wb, inputs, mask, case_params, dpreds, saved, g, sc, ws, batch, act_dtype, stream, nullptr);
}

int fno_backward_ex(const fno_weights* w, const fno_weights_bwd* wb, const float* inputs, const float* mask,
                    const float* case_params, const float* dpreds, const fno_train_saved* saved,
                    const fno_grads* g, const fno_bwd_scratch* sc, const fno_workspace* ws, int batch,
                    int act_dtype, void* stream, void* const* seg_events) {
  if (!w || !wb || !inputs || !mask || !dpreds || !saved || !g || !sc || !ws || batch <= 0 || bad_dtype(act_dtype))
    return fail(kErrArg, "fno_backward: bad argument");
  if (!sc->d[0] || !sc->d[1] || !sc->dz1 || !sc->gm || !sc->gwk || !sc->partials || !ws->ym || !ws->z)
    return fail(kErrArg, "fno_backward: null scratch buffer");
  cudaStream_t st = S(stream);
  const int L = w->n_layers, p = w->n_case_params;
  const bool bf = act_dtype == FNO_ACT_BF16;
  // Small gradients: every CTA stores its share as one row of sc->partials and a second launch adds the rows up in index
  // order (reduce_partials_kernel) -- no atomics, so the gradients are bit-for-bit reproducible.  The targets accumulate
  // over batch chunks / launches: clear them first.
  float* part_co = sc->partials;                                   // chan_outer: up to 296 rows x (128*32 + 128)
  float* part_pb = part_co + static_cast<size_t>(296) * (kProj * kC + kProj);   // project_bwd: 16 * chunk rows x 386
  float* part_lb = part_pb + static_cast<size_t>(project_bwd_parts(FNO_BWD_CHUNK)) * project_bwd_row();   // lift_bwd
  // no memsets: every gradient is WRITTEN by its (first) reduction -- reduce_partials with accumulate = 0,
  // lift_bwd_reduce, unpack_spectral_grads -- and only further batch chunks of the project stage accumulate
  // ---- project backward (batch chunks bound the dz1 scratch) -> d[0] = dpre_{L-1}
  const size_t act_elt = bf ? 2 : 4;
  for (int b0 = 0; b0 < batch; b0 += FNO_BWD_CHUNK) {
    const int nb = (batch - b0 < FNO_BWD_CHUNK) ? batch - b0 : FNO_BWD_CHUNK;
    const char* a_l = static_cast<const char*>(saved->act[L]) + static_cast<size_t>(b0) * kC * kHW * act_elt;
    const float* pre = saved->pre[L - 1] + static_cast<size_t>(b0) * kC * kHW;
    float* dout = sc->d[0] + static_cast<size_t>(b0) * kC * kHW;
    const float* dp = dpreds + static_cast<size_t>(b0) * 2 * kHW;
    const float* mk = mask + static_cast<size_t>(b0) * kHW;
    // the tensor-core kernel (fno_project_bwd_tc.cu); FNO_PBWD_TC=0 selects the CUDA-core kernel (A/B measurements)
    static const bool use_tc = [] { const char* v = getenv("FNO_PBWD_TC"); return !(v && v[0] == '0'); }();
    int rows = project_bwd_parts(nb);
    const int rs = project_bwd_row();
    cudaError_t e;
    if (use_tc) {
      e = bf ? launch_project_bwd_tc<__nv_bfloat16>(a_l, dp, mk, pre, w->fc1_w, w->fc1_b, w->fc2_w, dout, sc->dz1, part_pb, &rows, nb, st)
             : launch_project_bwd_tc<float>(a_l, dp, mk, pre, w->fc1_w, w->fc1_b, w->fc2_w, dout, sc->dz1, part_pb, &rows, nb, st);
    } else {
      e = bf ? launch_project_bwd<__nv_bfloat16>(a_l, dp, mk, pre, w->fc1_w, w->fc1_b, w->fc2_w, dout, sc->dz1, part_pb, nb, st)
             : launch_project_bwd<float>(a_l, dp, mk, pre, w->fc1_w, w->fc1_b, w->fc2_w, dout, sc->dz1, part_pb, nb, st);
    }
    FNO_CUDA(e, "project_bwd_kernel");
    const int accum = b0 > 0 ? 1 : 0;
    FNO_CUDA(launch_reduce_partials(part_pb, rows, rs, g->fc2_w, 2 * kProj, g->fc1_b, kProj, g->fc2_b, 2, accum, st),
             "reduce(fc2.weight | fc1.bias | fc2.bias)");
    int n_co = 0;
    e = bf ? launch_chan_outer<float, __nv_bfloat16, 128, 32>(sc->dz1, a_l, part_co, &n_co, nb, st)
           : launch_chan_outer<float, float, 128, 32>(sc->dz1, a_l, part_co, &n_co, nb, st);
    FNO_CUDA(e, "chan_outer_kernel(fc1)");
    FNO_CUDA(launch_reduce_partials(part_co, n_co, kProj * kC + kProj, g->fc1_w, kProj * kC, nullptr, 0, nullptr, 0, accum, st),
             "reduce(fc1.weight)");
  }
  auto mark = [&](int seg) -> cudaError_t {   // the gradients of segment `seg` are final from here on (stream order)
    if (seg_events == nullptr || seg_events[seg] == nullptr) return cudaSuccess;
    return cudaEventRecord(static_cast<cudaEvent_t>(seg_events[seg]), st);
  };
  FNO_CUDA(mark(0), "cudaEventRecord(fc1/fc2 gradients)");
  // ---- Fourier blocks, last to first
  const float inv = 1.f / static_cast<float>(kHW);
  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    float* dpre = sc->d[cur];
    float* dnext = sc->d[cur ^ 1];
    int n_co = 0;
    cudaError_t e = bf ? launch_chan_outer<float, __nv_bfloat16, 32, 32>(dpre, saved->act[l], part_co, &n_co, batch, st)
                       : launch_chan_outer<float, float, 32, 32>(dpre, saved->act[l], part_co, &n_co, batch, st);
    FNO_CUDA(e, "chan_outer_kernel(w0)");
    FNO_CUDA(launch_reduce_partials(part_co, n_co, kC * kC + kC, g->w0_w[l], kC * kC, g->w0_b[l], kC, nullptr, 0, 0, st),
             "reduce(w0.weight | w0.bias)");
    FNO_TRY(fno_spectral_dft_fwd(dpre, sc->gm, batch, FNO_ACT_F32, inv, 2.f * inv, stream));
    FNO_CUDA(launch_spectral_wgrad(saved->xm[l], sc->gm, sc->gwk, batch, st), "spectral_wgrad_kernel");
    FNO_TRY(fno_unpack_spectral_grads(sc->gwk, g->spec_w1[l], g->spec_w2[l], stream));
    FNO_CUDA(mark(1 + (L - 1 - l)), "cudaEventRecord(block gradients)");
    FNO_TRY(fno_mode_mix(sc->gm, wb->spec_wkT[l], ws->ym, batch, stream));
    FNO_TRY(fno_spectral_inv_kx(ws->ym, ws->z, batch, 1.f, 1.f, stream));
    FNO_TRY(fno_block_out(l > 0 ? FNO_EPI_MUL_DGELU : FNO_EPI_PLAIN, ws->z, dpre, wb->w0[l], nullptr, dnext, nullptr,
                          l > 0 ? saved->pre[l - 1] : nullptr, batch, FNO_ACT_F32, stream));
    cur ^= 1;
  }
  FNO_CUDA(launch_lift_bwd(sc->d[cur], inputs, mask, case_params, w->gx, w->gy, g->fc0_w, g->fc0_b, part_lb, batch, p, st),
           "lift_bwd_kernel");
  FNO_CUDA(mark(L + 1), "cudaEventRecord(fc0 gradients)");
  return kOk;
}

int fno_multistep_metrics(const float* preds_seq, const float* label_u, const float* mask, float* sums, int steps,
                          int batch, void* stream) {
  if (!preds_seq || !label_u || !mask || !sums || steps <= 0 || batch <= 0)
    return fail(kErrArg, "fno_multistep_metrics: bad argument");
  FNO_CUDA(launch_multistep_metrics(preds_seq, label_u, mask, sums, steps, batch, S(stream)), "multistep_metrics_kernel");
  return kOk;
}

int fno_gather_batch(const void* frames_in, const void* frames_out, const float* case_table, const int32_t* case_ids,
                     const int64_t* idx, int n_idx, int n_case_params, int frame_dtype, float* inputs, float* label,
                     float* mask, float* case_params, void* stream) {
  if (!frames_in || !frames_out || !case_ids || !idx || !inputs || !label || !mask || n_idx <= 0 || n_case_params < 0 ||
      n_case_params > kMaxCaseParams || bad_dtype(frame_dtype) || (n_case_params > 0 && (!case_table || !case_params)))
    return fail(kErrArg, "fno_gather_batch: bad argument");
  FNO_CUDA(launch_gather_batch(frames_in, frames_out, case_table, reinterpret_cast<const int*>(case_ids),
                               reinterpret_cast<const long long*>(idx), n_idx, n_case_params, frame_dtype == FNO_ACT_BF16,
                               inputs, label, mask, case_params, S(stream)),
           "gather_batch_kernel");
  return kOk;
}

size_t fno_loss_scratch_bytes(void) { return loss_scratch_bytes(); }

int fno_loss_fwd(const float* preds, const float* labels, size_t n, void* scratch, float* out, void* stream) {
  if (!preds || !labels || !scratch || !out || n == 0) return fail(kErrArg, "fno_loss_fwd: bad argument");
  if ((reinterpret_cast<uintptr_t>(preds) | reinterpret_cast<uintptr_t>(labels)) & 15)
    return fail(kErrArg, "fno_loss_fwd: preds / labels must be 16-byte aligned");
  FNO_CUDA(launch_loss_fwd(preds, labels, n, static_cast<float*>(scratch), out, S(stream)), "loss_fwd_kernel");
  return kOk;
}

int fno_loss_bwd(const float* preds, const float* labels, const float* fwd, const float* gout, float* dpreds, size_t n,
                 void* stream) {
  if (!preds || !labels || !fwd || !gout || !dpreds || n == 0) return fail(kErrArg, "fno_loss_bwd: bad argument");
  FNO_CUDA(launch_loss_bwd(preds, labels, fwd, gout, dpreds, n, S(stream)), "loss_bwd_kernel");
  return kOk;
}

int fno_adam_step(const fno_adam_tensors* t, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int64_t step, void* stream) {
  if (!t || t->count < 0 || t->count > FNO_ADAM_MAX_TENSORS || step < 1)
    return fail(kErrArg, "fno_adam_step: bad argument");
  for (int i = 0; i < t->count; ++i)
    if (!t->param[i] || !t->grad[i] || !t->exp_avg[i] || !t->exp_avg_sq[i] || t->n[i] <= 0)
      return fail(kErrArg, "fno_adam_step: null tensor or empty size");
  FNO_CUDA(launch_adam_step(t, lr, beta1, beta2, eps, weight_decay, step, S(stream)), "adam_step_kernel");
  return kOk;
}

}  // extern "C"
