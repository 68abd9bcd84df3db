/* cfdbench_b200 -- C ABI of the B200-native FNO hot path for CFDBench.
 *
 * The reference (luo-yining/CFDBench @ 6c30c62) is pure Python/PyTorch: its "FFI" for this path is the
 * set of ATen calls made by src/models/fno/fno2d.py.  Each entry point below replaces the calls cited
 * next to it.  All pointers are raw device pointers unless a name ends in _host; buffers are owned by the
 * caller (PyTorch's allocator in the Python wrapper); every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*) and returns 0 on success, non-zero otherwise
 * (fno_last_error() gives the message).  No entry point synchronises the device, none falls back to CPU.
 *
 * Fixed configuration (the reference's FNO config, src/args.py:99-103,187-197): H=W=64, hidden=32,
 * modes 12x12, fc1 width 128, out_chan=2, in_chan=2.  Activation storage `act_dtype`:
 * FNO_ACT_F32 (parity mode) or FNO_ACT_BF16 (hidden activations stored as bf16 between kernels);
 * arithmetic is fp32 in both.
 *
 * Layouts
 *   activations      [B][32][64][64]   act_dtype, NCHW contiguous
 *   frames / preds   [B][2][64][64]    float32
 *   mask             [B][64][64]       float32  (a (B,1,64,64) tensor has the same layout)
 *   case_params      [B][p]            float32
 *   modes (xm, ym)   [288][B][32]      complex64 (interleaved re,im), MODE-major; mode k = kxi*12 + ky,
 *                                      kxi 0..11 <-> kx 0..11 (weights1), kxi 12..23 <-> kx 52..63 (weights2)
 *   packed spectral  [288][32 in][32 out] complex64 (fno_pack_spectral_weights)
 *   mix operand      per mode 2 x [64 = (out, re|im)][64 = (in, re|im)] float32: the real-expanded block as tf32
 *                    hi / lo images in the tensor core's K-major operand layout (fno_pack_mix_operand)
 *   w0t              [32 in][32 out]   float32  (transpose of the Conv2d weight (out,in,1,1))
 */
#ifndef CFDBENCH_B200_H_
#define CFDBENCH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNO_ABI_VERSION 3
#define FNO_MAX_LAYERS 8

enum { FNO_ACT_F32 = 0, FNO_ACT_BF16 = 1 };
/* epilogues of fno_block_out */
enum { FNO_EPI_GELU = 0, FNO_EPI_GELU_SAVE_PRE = 1, FNO_EPI_MUL_DGELU = 2, FNO_EPI_PLAIN = 3 };

/* Device-side weights of one model, in kernel layouts (built by the fno_pack_* calls). */
typedef struct fno_weights {
  int32_t n_layers;      /* reference fno_depth (4) */
  int32_t n_case_params; /* p: 5 cavity, 8 cylinder (reference src/utils/autoregressive.py:31-37) */
  const float* fc0_w;    /* [32][5+p]  fc0.weight  (reference fno2d.py:150-156) */
  const float* fc0_b;    /* [32] */
  const void* spec_wk[FNO_MAX_LAYERS]; /* mix operand of blocks.{l}.conv0.weights1/2 (fno_pack_mix_operand) */
  const float* w0t[FNO_MAX_LAYERS];    /* blocks.{l}.w0.weight transposed */
  const float* w0_b[FNO_MAX_LAYERS];   /* blocks.{l}.w0.bias */
  const float* fc1_w;    /* [128][32] fc1.weight */
  const float* fc1_b;    /* [128] */
  const float* fc2_w;    /* [2][128]  fc2.weight */
  const float* fc2_b;    /* [2] */
  const float* gx;       /* [64] float32(np.linspace(0,1,64)): x coordinate per row h (fno2d.py:244-255) */
  const float* gy;       /* [64] y coordinate per column w */
} fno_weights;

/* Caller-provided scratch for a batch of B samples. */
typedef struct fno_workspace {
  void* act[2]; /* two activation buffers, B*32*4096 elements of act_dtype each (ping-pong) */
  void* xm;     /* B*288*32 complex64 */
  void* ym;     /* B*288*32 complex64 */
  void* z;      /* B*64*24*32 float32: rows of the half-inverted spectrum, Z[b][h][2 ky + (re|im)][o] */
  void* ym_img; /* fno_ym_image_bytes(B), or NULL.  When set and act_dtype == FNO_ACT_BF16, inference blocks take the
                 * fused output stage (fno_mode_mix_image + fno_block_fused) and `ym` / `z` are not touched. */
} fno_workspace;

int fno_version(void);
/* Frees what the library itself owns on the CURRENT device (constant operand tables built on first use, events of the
 * chunked host path); they are rebuilt on demand.  Everything else is caller-owned.  Synchronises the device. */
int fno_destroy(void);
const char* fno_last_error(void);
/* bytes of one activation buffer / one mode buffer for batch B */
size_t fno_act_bytes(int batch, int act_dtype);
size_t fno_modes_bytes(int batch);
size_t fno_z_bytes(int batch);
size_t fno_ym_image_bytes(int batch);
size_t fno_bwd_partials_bytes(void);

/* weights1, weights2: (32,32,12,12) complex64 as stored by the reference (fno2d.py:31-51).
 * conj_transpose=0 -> wk[k][i][o] = W[i][o][k] (forward); 1 -> wk[k][o][i] = conj(W[i][o][k]) (adjoint). */
int fno_pack_spectral_weights(const void* weights1, const void* weights2, void* wk, int conj_transpose, void* stream);
/* packed weights wk[288][32][32] -> the operand image fno_mode_mix consumes (fno_mix_operand_bytes() bytes).
 * Run once per weight update; for the adjoint mix feed it the conj_transpose=1 pack. */
size_t fno_mix_operand_bytes(void);
int fno_pack_mix_operand(const void* wk, void* wop, void* stream);
/* both steps in one launch: weights1/weights2 -> operand image (what the module runs after every optimizer step) */
int fno_pack_mix_operand_from_weights(const void* weights1, const void* weights2, void* wop, int conj_transpose,
                                      void* stream);
/* gradient wrt packed forward weights [288][32][32] -> gradients of weights1 / weights2 */
int fno_unpack_spectral_grads(const void* gwk, void* gw1, void* gw2, void* stream);

/* Channel assembly + fc0: torch.cat/repeat/get_coords/Conv2d(5+p,32,1) (reference fno2d.py:195-217,244-255) */
int fno_lift_fwd(const float* inputs, const float* mask, const float* case_params, const fno_weights* w,
                 void* act_out, int batch, int act_dtype, void* stream);

/* The four phases of SpectralConv2d_fast + FnoBlock (reference fno2d.py:59-82, 106-112):            */
/* (1) torch.fft.rfft2 restricted to the kept modes (fno2d.py:62,73-78); outputs scaled by s0 (ky=0), s1 (ky>0) */
int fno_spectral_dft_fwd(const void* act_in, void* xm, int batch, int act_dtype, float s0, float s1, void* stream);
/* (1') the same transform for bf16 planes as two chained tensor-core GEMMs (fno_dft_fwd_tc.cu, warp-specialised since
 *      round 2: 25 us per launch at B=256 against 35 us for the register-FFT kernel, same 2e-6 against float64).
 *      fno_spectral_dft_fwd(..., FNO_ACT_BF16, ...) routes here (environment FNO_DFT_TC=0 selects the register kernel for
 *      A/B measurements); this entry point calls it directly. */
int fno_spectral_dft_fwd_tc(const void* act_in_bf16, void* xm, int batch, float s0, float s1, void* stream);
/* (2) einsum("bixy,ioxy->boxy") on both corners (fno2d.py:54-57,73-78); wop = fno_pack_mix_operand image */
int fno_mode_mix(const void* xm, const void* wop, void* ym, int batch, void* stream);
/* (3) first half of irfft2 on the zero-padded spectrum (fno2d.py:65-72,81): inverse C2C along kx of the 24 kept
 *     rows, scaled by s0 (ky=0) / s1 (ky>0) (forward: 1/4096, 2/4096): z[b][h][2 ky + (re|im)][o], fno_z_bytes(B). */
int fno_spectral_inv_kx(const void* ym, void* z, int batch, float s0, float s1, void* stream);
/* (4) second half of irfft2 (C2R along ky, Im of the ky=0 column dropped) + Conv2d(32,32,1) + add + GELU
 *     (fno2d.py:81,104-111) as one tensor-core GEMM per 128-pixel tile.  pre_out/pre_in: see FNO_EPI_*. */
int fno_block_out(int epilogue, const void* z, const void* act_in, const float* w0t, const float* bias, void* act_out,
                  float* pre_out, const float* pre_in, int batch, int act_dtype, void* stream);
/* bf16 storage, inference: the output stage of a Fourier block in ONE kernel (block_fused_kernel, fno_block_fused.cu)
 * -- replaces fno_spectral_inv_kx + fno_block_out(FNO_EPI_GELU), i.e. irfft2 + Conv2d(32,32,1) + add + GELU of
 * reference src/models/fno/fno2d.py:81,104-111, without the Z round trip through HBM.
 *   fno_mode_mix_image: same product as fno_mode_mix, written as the per-sample tensor-core operand image the fused
 *     kernel bulk-copies (tf32 hi/lo split, fno_ym_image_bytes(B) bytes).
 *   fno_block_fused:    act_out = GELU(irfft2(pad(Y)) + W0 act_in + bias); act_in / act_out bf16 [B][32][64][64]. */
int fno_mode_mix_image(const void* xm, const void* wop, void* ym_img, int batch, void* stream);
int fno_block_fused(const void* ym_img, const void* act_in_bf16, const float* w0t, const float* bias, void* act_out_bf16,
                    int batch, void* stream);
/* all four: act_out = FnoBlock_l(act_in) */
int fno_block_fwd(const fno_weights* w, int layer, const void* act_in, void* act_out, float* pre_out,
                  const fno_workspace* ws, int batch, int act_dtype, void* stream);

/* fc1 + GELU + fc2 + "* mask" (reference fno2d.py:228-233) */
int fno_project_fwd(const void* act_in, const float* mask, const fno_weights* w, float* preds, int batch,
                    int act_dtype, void* stream);

/* Fno2d.forward without the loss (reference fno2d.py:178-233): preds[B][2][64][64] */
int fno_forward(const fno_weights* w, const float* inputs, const float* mask, const float* case_params,
                float* preds, const fno_workspace* ws, int batch, int act_dtype, void* stream);

/* Fno2d.generate_many (reference fno2d.py:269-295): preds_seq[steps][B][2][64][64], step s feeds step s+1 */
int fno_rollout(const fno_weights* w, const float* inputs, const float* mask, const float* case_params,
                float* preds_seq, int steps, const fno_workspace* ws, int batch, int act_dtype, void* stream);

/* Same, with HOST buffers (pinned or pageable): copies inputs/mask/case_params to the device scratch
 * frames `dev_io` (caller-provided: (2 + 1 + steps*2)*B*4096*4 + B*p*4 bytes), runs the rollout and copies
 * preds_seq back; returns after the D2H copy has been enqueued (synchronise the stream to read). */
int fno_rollout_host(const fno_weights* w, const float* inputs_host, const float* mask_host,
                     const float* case_params_host, float* preds_seq_host, int steps, const fno_workspace* ws,
                     void* dev_io, int batch, int act_dtype, void* stream);
/* One step (steps = 1) for host buffers, pipelined over n_chunks equal batch chunks: uploads, kernels and downloads
 * run on three caller-provided streams chained by events, so chunk c+1's upload overlaps chunk c's kernels and
 * chunk c-1's download while copies of one direction stay serialised.  ws_chunks / dev_io_chunks: one workspace and
 * one fno_rollout_host_scratch_bytes(batch / n_chunks, p, 1) buffer per chunk.  The caller orders the three streams
 * after its own stream before the call and waits for stream_out afterwards. */
int fno_rollout_host_chunked(const fno_weights* w, const float* inputs_host, const float* mask_host,
                             const float* case_params_host, float* preds_host, const fno_workspace* ws_chunks,
                             void* const* dev_io_chunks, int batch, int n_chunks, int act_dtype, void* stream_in,
                             void* stream_compute, void* stream_out);
size_t fno_rollout_host_scratch_bytes(int batch, int n_case_params, int steps);

/* ---------------------------------------------------------------------------------------------------
 * Training step (what torch.autograd derives for loss["nmse"].backward(), reference src/train_auto.py:255)
 * ------------------------------------------------------------------------------------------------- */

/* Buffers the forward pass leaves for the backward pass (caller-allocated, batch B). */
typedef struct fno_train_saved {
  void* act[FNO_MAX_LAYERS + 1]; /* a_0 (lift output) .. a_L (last block output), act_dtype */
  float* pre[FNO_MAX_LAYERS];    /* pre-activation of each block, float32 [B][32][64][64] */
  void* xm[FNO_MAX_LAYERS];      /* kept modes of a_l, complex64 [288][B][32] */
} fno_train_saved;

/* Extra weight views the backward pass needs. */
typedef struct fno_weights_bwd {
  const void* spec_wkT[FNO_MAX_LAYERS]; /* mix operand of fno_pack_spectral_weights(..., conj_transpose=1) */
  const float* w0[FNO_MAX_LAYERS];      /* blocks.{l}.w0.weight in its natural [out][in] layout */
} fno_weights_bwd;

/* Gradient outputs, reference parameter layouts (float32 / complex64). All are overwritten. */
typedef struct fno_grads {
  float* fc0_w; /* [32][5+p] */
  float* fc0_b; /* [32] */
  void* spec_w1[FNO_MAX_LAYERS]; /* (32,32,12,12) complex64 */
  void* spec_w2[FNO_MAX_LAYERS];
  float* w0_w[FNO_MAX_LAYERS]; /* [32][32] */
  float* w0_b[FNO_MAX_LAYERS]; /* [32] */
  float* fc1_w; /* [128][32] */
  float* fc1_b; /* [128] */
  float* fc2_w; /* [2][128] */
  float* fc2_b; /* [2] */
} fno_grads;

/* Scratch of the backward pass. */
typedef struct fno_bwd_scratch {
  float* d[2];   /* two float32 [B][32][64][64] gradient buffers (ping-pong) */
  float* dz1;    /* float32 [min(B,FNO_BWD_CHUNK)][128][64][64] */
  void* gm;      /* complex64 [B][288][32]: scaled modes of the block's upstream gradient */
  void* gwk;     /* complex64 [288][32][32] */
  float* partials; /* fno_bwd_partials_bytes() bytes: per-CTA shares of the small gradients (fc0/fc1/fc2/w0), summed in a
                    * fixed order by a second launch instead of float atomics -> bit-reproducible gradients */
} fno_bwd_scratch;
#define FNO_BWD_CHUNK 32

/* Fno2d.forward (reference fno2d.py:178-233) that also fills `saved`. ws->ym is used as scratch. */
int fno_forward_train(const fno_weights* w, const float* inputs, const float* mask, const float* case_params,
                      float* preds, const fno_train_saved* saved, const fno_workspace* ws, int batch,
                      int act_dtype, void* stream);

/* Backward of the above given dL/dpreds (float32 [B][2][64][64]); fills `grads`. */
int fno_backward(const fno_weights* w, const fno_weights_bwd* wb, const float* inputs, const float* mask,
                 const float* case_params, const float* dpreds, const fno_train_saved* saved,
                 const fno_grads* grads, const fno_bwd_scratch* scratch, const fno_workspace* ws, int batch,
                 int act_dtype, void* stream);
/* Same, recording CUDA events as gradient segments become final so that the caller can start their all-reduce while
 * the rest of the backward pass still runs (data-parallel training): seg_events[0] after the fc1 / fc2
 * gradients, seg_events[1 + i] after the gradients of block (n_layers - 1 - i), seg_events[n_layers + 1] after the fc0
 * gradients.  seg_events = NULL or a NULL entry: nothing recorded there.  Entries are cudaEvent_t. */
int fno_backward_ex(const fno_weights* w, const fno_weights_bwd* wb, const float* inputs, const float* mask,
                 const float* case_params, const float* dpreds, const fno_train_saved* saved,
                 const fno_grads* grads, const fno_bwd_scratch* scratch, const fno_workspace* ws, int batch,
                 int act_dtype, void* stream, void* const* seg_events);

/* Rollout evaluation on the device (reference src/test_multistep.py:73-83,153-177 get_metrics on the
 * masked u channel, three .item() syncs per step and case there).  preds_seq [S][B][2][64][64], label_u and mask
 * [S][B][64][64]; sums [S][B][3] = (sum (p-l)^2, sum l^2, sum |p-l|) with p, l multiplied by mask. */
int fno_multistep_metrics(const float* preds_seq, const float* label_u, const float* mask, float* sums, int steps,
                          int batch, void* stream);

/* one training batch gathered on the device from resident frames -- replaces DataLoader indexing +
 * collate_fn (reference src/train_auto.py:33-58: stack, channel slices, case-parameter dict loop, four .cuda() copies).
 * frames_in / frames_out: [N][3][64][64] (u, v, mask) as the dataset holds them (src/dataset/cavity.py:326-331), float32
 * (frame_dtype = FNO_ACT_F32) or bfloat16 (FNO_ACT_BF16); case_table [n_cases][p]; case_ids [N] int32; idx [n_idx] int64.
 * Outputs (float32): inputs [n][2][64][64], label [n][2][64][64], mask [n][1][64][64], case_params [n][p]. */
int fno_gather_batch(const void* frames_in, const void* frames_out, const float* case_table, const int32_t* case_ids,
                     const int64_t* idx, int n_idx, int n_case_params, int frame_dtype, float* inputs, float* label,
                     float* mask, float* case_params, void* stream);

/* ---- the rest of a training step (reference src/train_auto.py:233-260) ------------------------------------- */

/* MseLoss.forward (reference src/models/loss.py:22-37) over n = preds.numel() float32 elements, one launch:
 * out[0..4] = mse, rmse, mae, nmse (= mse / mean(labels^2)), mean(labels^2).  scratch: fno_loss_scratch_bytes()
 * bytes, zero-initialised once by the caller (the kernel leaves it zeroed).  Deterministic. */
size_t fno_loss_scratch_bytes(void);
int fno_loss_fwd(const float* preds, const float* labels, size_t n, void* scratch, float* out, void* stream);
/* dpreds[i] = d(sum_j gout[j] * out[j]) / dpreds[i], gout = upstream gradients of (mse, rmse, mae, nmse);
 * `fwd` is the out[] of the matching fno_loss_fwd call. */
int fno_loss_bwd(const float* preds, const float* labels, const float* fwd, const float* gout, float* dpreds, size_t n,
                 void* stream);

/* torch.optim.Adam.step (train_auto.py:213,256; no amsgrad, maximize=False) for up to FNO_ADAM_MAX_TENSORS parameter
 * tensors in ONE launch.  Every array holds float32 views (a complex64 tensor is 2*numel floats, as
 * torch.view_as_real presents it to Adam); n[i] = number of floats.  `step` is the 1-based step count. */
#define FNO_ADAM_MAX_TENSORS 32
typedef struct fno_adam_tensors {
  int32_t count;
  void* param[FNO_ADAM_MAX_TENSORS];
  const void* grad[FNO_ADAM_MAX_TENSORS];
  void* exp_avg[FNO_ADAM_MAX_TENSORS];
  void* exp_avg_sq[FNO_ADAM_MAX_TENSORS];
  int64_t n[FNO_ADAM_MAX_TENSORS];
} fno_adam_tensors;
int fno_adam_step(const fno_adam_tensors* t, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int64_t step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFDBENCH_B200_H_ */
