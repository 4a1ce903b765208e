/* qeb.h - C ABI of libqeb_sm100.so: the B200 (sm_100a) hot path of
 * tataganesh/Query-Efficient-Approx-to-improve-OCR.
 *
 * The reference is pure Python on PyTorch and has no FFI layer of its own: the seam is the
 * nn.Module.forward / CTCLoss.__call__ / helper-function signatures. Each entry point below names the reference
 * call site (file:line under the reference root) whose work it performs; the Python side that binds these with
 * ctypes and keeps the reference signatures is query-efficient-approx-to-improve-ocr_b200/mirror/ (see
 * INTEGRATION.md for the stub a reference maintainer adds).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated; plain pointers and sizes only, no torch types;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - int-returning functions return 0 on success, <0 on failure (QEB_ERR_*), message via qeb_last_error()
 *    (thread-local); shape/alignment violations are reported before anything is launched;
 *  - the caller owns every buffer, workspaces included; the library allocates nothing persistent;
 *  - re-entrant: may be called from the trainer thread and from the autograd engine thread.
 */
#ifndef QEB_H
#define QEB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QEB_OK 0
#define QEB_ERR_INVALID (-1)
#define QEB_ERR_CUDA (-2)
#define QEB_ERR_UNSUPPORTED (-3)

/* ---- plumbing -------------------------------------------------------------------------------------------- */
const char* qeb_last_error(void);
int qeb_abi_version(void);
long long qeb_launch_count(void);      /* kernels launched by this library since the last reset */
void qeb_reset_launch_count(void);
int qeb_check_device(void);            /* 0 iff the current device is an sm_100 part */

/* ---- CTC loss: torch.nn.CTCLoss(blank=0, reduction, zero_infinity=False) -----------------------------------
 * ctor train_nn_patch.py:143-144, train_nn_area.py:146-148, train_crnn.py:130-131;
 * calls train_nn_patch.py:178,294, train_nn_area.py:174,265, train_crnn.py:160, tracking_utils.py:68,72.
 * log_probs (T,B,V) fp32, element strides st_t/st_b, class stride 1. batch_index (optional, B ints): sample b uses
 * column batch_index[b] - the scores[:, img_indices, :] subset of tracking_utils.py:65 without the gather copy.
 * targets: concatenated int32; tgt_offsets: exclusive prefix sum of target_lengths (B). reduction: 0 none, 1 mean
 * (mean_b(nll_b / max(1,len_b))), 2 sum. log_alpha: workspace of qeb_ctc_workspace_bytes(), kept for the backward.
 * nll (B): per-sample negative log-likelihood (+inf when infeasible). loss_out: scalar, may be NULL for none.
 * Backward writes d loss / d log_probs in ATen's convention ((exp(lp) - exp(lse(alpha+beta) + nll - lp)) * grad_out:
 * the gradient at the logits); an infeasible sample yields NaN rows unless zero_infinity (the reference zeroes them
 * in CRNN.backward_hook, models/model_crnn.py:30-32). grad_out: (B) for none, (1) otherwise. With batch_index the
 * rows' gradients are ADDED into the listed columns of a grad buffer the caller zero-filled (a column may be listed several
 * times: the history depths of weighted_ctc_loss in ONE launch); the other columns are not touched. */
size_t qeb_ctc_workspace_bytes(int B, int T, int max_target_len);
int qeb_ctc_fwd(const float* log_probs, long long st_t, long long st_b, const int* batch_index, const int* targets,
                const int* tgt_offsets, const int* input_lengths, const int* target_lengths, int B, int T, int V,
                int blank, int max_target_len, int reduction, int zero_infinity, float* log_alpha, float* nll,
                float* loss_out, void* stream);
int qeb_ctc_bwd(const float* log_probs, long long st_t, long long st_b, const int* batch_index, const int* targets,
                const int* tgt_offsets, const int* input_lengths, const int* target_lengths, int B, int T, int V,
                int blank, int max_target_len, int reduction, int zero_infinity, const float* log_alpha,
                const float* nll, const float* grad_out, float* grad, long long gst_t, long long gst_b, void* stream);

/* ---- log-softmax over the class dimension: fn.log_softmax(self.linear(x), 2), models/model_crnn.py:20 ------ */
int qeb_log_softmax_fwd(const float* x, float* y, long long rows, int V, void* stream);
int qeb_log_softmax_bwd(const float* y, const float* dy, float* dx, long long rows, int V, void* stream);

/* ---- batched Levenshtein + CER: Levenshtein.distance(label, pred) / max(1, len(label)), utils.py:103-109 ----
 * (python-Levenshtein==0.12.0, requirements.txt:70). a = labels, b = predictions. Symbols are uint8 (sym_bytes 1) or
 * int32 code points / class indices (sym_bytes 4). Each side is CSR (len NULL, off has n+1 entries) or padded rows
 * (off = row starts, len = used symbols). max_len: upper bound on any string length. dist (n) int32; cer (n) fp64 or
 * NULL; scratch: one int. */
int qeb_levenshtein_batch(const void* a_syms, const int* a_off, const int* a_len, const void* b_syms, const int* b_off,
                          const int* b_len, int n, int sym_bytes, int max_len, int* dist, double* cer, int* scratch,
                          void* stream);

/* ---- greedy CTC decode: pred_to_string, utils.py:74-92 ---------------------------------------------------------
 * scores (T,B,V) fp32 (strides st_t/st_b) -> out (B,T) int32 class indices padded with -1, out_len (B);
 * raw_path (B,T) optional per-timestep arg-max. First maximal index wins (torch.argmax). */
int qeb_greedy_decode(const float* scores, long long st_t, long long st_b, int T, int B, int V, int blank, int* out,
                      int* out_len, int* raw_path, void* stream);
/* The collapse half of pred_to_string (utils.py:84-89) on a per-frame arg-max path already on the device (the fused CRNN
 * head writes one): frame t of sample b at path[t*st_t + b*st_b]; out / out_len as qeb_greedy_decode. */
int qeb_greedy_collapse(const int* path, long long st_t, long long st_b, int T, int B, int blank, int* out, int* out_len,
                        void* stream);

/* ---- minibatch-subset selection ---------------------------------------------------------------------------------
 * TopKCERSampler.query selection_utils.py:144-151: per segment the indices of the k largest fp32 CERs, descending,
 * equal values lowest-index-first. CerRangeSampler.query selection_utils.py:107-135: per segment k points
 * (max-min)*rand+min, then sequentially the first arg-min of |point - copy| with copy[idx]=100.
 * vals: all segments concatenated; seg_off (n_seg+1); seg_k (n_seg); out_off (n_seg): exclusive prefix sum of the
 * picks per segment (min(k,n) for topk, k for range); out_idx: int64, local to the segment. rands: torch.rand draws
 * made on the host, concatenated like out_idx. work: scratch, same size as vals. points_out: optional. */
int qeb_cer_topk_segmented(const float* vals, const int* seg_off, const int* seg_k, const int* out_off, int n_seg,
                           long long* out_idx, void* stream);
/* One segment of ANY size (dataset-wide top-k: pruning/methods.py:5-8; the per-rank shard reduction of a sharded top-k):
 * the same order as above - out_idx[r] = index of the r-th largest value, equal values lowest index first - through a
 * 64-bit-key sort instead of the per-warp ranking, which is quadratic in the segment size. k is clamped to n.
 * work: qeb_cer_topk_global_workspace_bytes(n) bytes, 8-byte aligned. */
size_t qeb_cer_topk_global_workspace_bytes(long long n);
int qeb_cer_topk_global(const float* vals, long long n, long long k, void* work, long long* out_idx, void* stream);
int qeb_cer_range_segmented(const float* vals, const int* seg_off, const int* seg_k, const int* out_off,
                            const float* rands, int n_seg, float* work, long long* out_idx, float* points_out,
                            void* stream);

/* ---- Gaussian jitter: AddGaussianNoice.__call__ transform_helper.py:33-45; add_noise train_nn_patch.py:187-191,
 * train_nn_area.py:184-191. out = clamp(img - coef * noise, 0, 1); noise = noise_in if given, else
 * mean + sigma[image] * N(0,1) from Philox4x32-10(seed). img/out/noise: (n_img, hw) fp32, hw % 4 == 0, 16B aligned. */
int qeb_gauss_jitter(const float* img, const float* sigma, float mean, float coef, const float* noise_in,
                     unsigned long long seed, long long n_img, int hw, float* out, float* noise_out, void* stream);
/* The same with the Philox key read from DEVICE memory (key = *seed_dev + seed_offset) at execution time: inside a captured
 * CUDA graph a by-value seed would be frozen and every replay would repeat the noise of the inner-loop copies
 * (train_nn_patch.py:278-296). */
int qeb_gauss_jitter_devseed(const float* img, const float* sigma, float mean, float coef,
                             const unsigned long long* seed_dev, unsigned long long seed_offset, long long n_img, int hw,
                             float* out, float* noise_out, void* stream);

/* ---- OCR hand-off: fp32 images in [0,1] -> the uint8 pixels ToPILImage gives the OCR engine for a float tensor
 * (pic.mul(255).byte(): fp32 multiply, truncation), ocr_helper/tess_helper.py:20-24, eocr_helper.py. x, out: n elements,
 * 16-byte aligned; the caller copies `out` to pinned host memory with its own asynchronous D2H. */
int qeb_to_uint8(const float* x, long long n, unsigned char* out, void* stream);

/* ---- crop + centre-pad with 1.0: get_text_stack / padder utils.py:118-141 -------------------------------------
 * img (H,W) fp32; boxes (n,4) int32 x_min,y_min,x_max,y_max; out (n,oh,ow). scatter is the adjoint (atomic adds into
 * a zero-initialised gimg (H,W)). */
int qeb_crop_pad_gather(const float* img, int H, int W, const int* boxes, int n, int oh, int ow, float* out,
                        void* stream);
int qeb_crop_pad_scatter(const float* gout, int H, int W, const int* boxes, int n, int oh, int ow, float* gimg,
                         void* stream);

/* ==== the two networks (module-level entry points: what nn.Module.forward / autograd backward bind to) ============
 * All activations are NHWC fp32 inside a caller-owned workspace; the module tensors (B,1,H,W) have one channel and are
 * therefore passed as they are. params / grads / buffers are arrays of DEVICE pointers living in HOST memory.
 *
 * CRNN: models/model_crnn.py:5-56 (Convolutional.forward :47-56, CRNN.forward :16-21 without the log_softmax, which is
 * qeb_log_softmax_fwd above so that CRNN.backward_hook sees the gradient at the logits, :30-32).
 * params (36, state_dict order): conv1.w,b .. conv5.w,b, batchnorm1.w,b, conv6.w,b, batchnorm2.w,b, conv7.w,b,
 *   lstm {weight_ih, weight_hh, bias_ih, bias_hh} for l0, l0_reverse, l1, l1_reverse, linear.w,b.
 * buffers (6): batchnorm1 {running_mean, running_var, num_batches_tracked(int64)}, batchnorm2 {...}.
 * x (B,1,32,W), W % 4 == 0; logits (T = W/4 - 1, B, V) dense, V <= 96. bn_train: 1 = batch statistics + running-stat
 * update (module.train()), 0 = running statistics (set_bn_eval, utils.py:113-115). ws: qeb_crnn_workspace_bytes(),
 * 256-byte aligned, must stay untouched between forward and backward.
 * backward: dlogits (T,B,V); grads[i] NULL = skip that gradient, otherwise the gradient is ACCUMULATED into it;
 * dx (B,1,32,W) or NULL. */
size_t qeb_crnn_workspace_bytes(int B, int W, int V);
int qeb_crnn_num_params(void);
int qeb_crnn_forward(const float* x, int B, int W, int V, const float* const* params, void* const* buffers, int bn_train,
                     void* ws, float* logits, void* stream);
/* The same forward with the two fusions north_star names, selected per argument. log_softmax = 1: `out` (T,B,V) receives
 * fn.log_softmax(self.linear(x), 2) (models/model_crnn.py:20) from the Linear GEMM's epilogue and argmax_path (T*B ints, frame
 * t of sample b at [t*B + b], nullable) the per-frame arg-max pred_to_string takes (utils.py:78-89; see qeb_greedy_collapse).
 * jit_sigma != NULL (B floats): x is the CLEAN batch and the network sees clamp(x - jit_coef*N(jit_mean, jit_sigma[b]), 0, 1)
 * (AddGaussianNoice, transform_helper.py:33-45; add_noise, train_nn_patch.py:187-191) generated inside conv1's input load
 * with the Philox stream of qeb_gauss_jitter (key jit_seed + *jit_seed_dev); noisy_out (B,1,32,W, required then) receives
 * that image - hand it to the OCR engine and pass it as `x` to qeb_crnn_backward - and noise_out (nullable) the noise. */
int qeb_crnn_forward_fused(const float* x, int B, int W, int V, const float* const* params, void* const* buffers,
                           int bn_train, void* ws, float* out, int log_softmax, int* argmax_path, const float* jit_sigma,
                           float jit_mean, float jit_coef, unsigned long long jit_seed,
                           const unsigned long long* jit_seed_dev, float* noisy_out, float* noise_out, void* stream);
int qeb_crnn_backward(const float* x, int B, int W, int V, const float* const* params, int bn_train, void* ws,
                      const float* dlogits, float* const* grads, float* dx, void* stream);

/* UNet: models/model_unet.py:7-109 (forward :49-76). params (64): 9 blocks (encoder1..4, bottleneck, decoder4..1) x
 * {conv1.w, norm1.w, norm1.b, conv2.w, norm2.w, norm2.b}, upconv4..1 {w, b}, conv {w, b}. buffers (54): per block
 * {norm1.running_mean, running_var, num_batches_tracked, norm2...}. x, y (B,1,H,W), H and W multiples of 16. */
size_t qeb_unet_workspace_bytes(int B, int H, int W);
int qeb_unet_num_params(void);
int qeb_unet_num_buffers(void);
int qeb_unet_forward(const float* x, int B, int H, int W, const float* const* params, void* const* buffers, int bn_train,
                     void* ws, float* y, void* stream);
int qeb_unet_backward(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws,
                      const float* y, const float* dy, float* const* grads, float* dx, void* stream);
/* The same backward for data-parallel training with an overlapped gradient exchange (NCCL all-reduces over ranges of the flat
 * gradient buffer, SURVEY.md 8(e): "bucketed in reverse-layer order and overlapped with the remaining wgrads"). Gradients
 * become final last layer first, i.e. from the END of the ABI parameter order: bucket_events = two cudaEvent_t, recorded as soon
 * as parameters [30, 64) (decoder blocks, up-convolutions, final conv: 12.2 MB) resp. [18, 30) (encoder block 4, bottleneck:
 * 17.7 MB) are final; parameters [0, 18) (1.1 MB) are final when the call's work completes. A communication stream that waits
 * for an event can reduce its range while the rest of the backward runs. */
int qeb_unet_backward_bucketed(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws,
                               const float* y, const float* dy, float* const* grads, float* dx, void* const* bucket_events,
                               void* stream);

/* ==== building blocks, exported for tests and for callers that compose their own graphs ===========================
 * Tensor-core contractions (tcgen05 kind::tf32, fp32 accumulate). Activations NHWC with a channel stride (a channel
 * slice of a wider buffer is passed as base pointer + stride); channel counts on the contracted side are multiples of 32.
 * conv_fprop: nn.Conv2d stride 1 (models/model_crnn.py:37-45, models/model_unet.py:84-104), nn.Linear and the LSTM input
 *   projections as 1x1; also every input gradient (flipped, transposed weights). wpacked [n_total][kh*kw*cin] from
 *   qeb_pack_weight (mode 0 fprop, mode 1 dgrad, mode 2 ConvTranspose fprop). v = acc*scale[c] + bias[c], optional ReLU.
 * conv_wgrad: dw[co][ci][kh][kw] += ... (torch layout), split-K with fp32 atomics.
 * convT2x2_*: nn.ConvTranspose2d(k=2, s=2) (models/model_unet.py:25-44) as a per-pixel GEMM + pixel shuffle. */
int qeb_pack_weight(const float* src, float* dst, int A, int B, int KH, int KW, int mode, void* stream);
int qeb_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int dst_cstride, void* stream);
int qeb_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int src_cstride, void* stream);
int qeb_conv_fprop_tc(const float* x, int n_img, int h_in, int w_in, int cin, int x_cstride, const float* wpacked,
                      int n_total, int kh, int kw, int ph, int pw, const float* bias, const float* scale, int relu,
                      float* out, int out_cstride, int accumulate, void* stream);

/* fp16-operand variant of qeb_conv_fprop_tc (forward pass): x16 = the input as NHWC fp16 (channel stride a multiple of
 * 8), w16 = the packed weights [n_total][kh*kw*cin] as fp16; fp32 accumulation and fp32 output as above, out16 (nullable) =
 * fp16 shadow of the output for the next layer. Same 11-bit significand as tf32, rounded to nearest. */
int qeb_conv_fprop_tc16(const void* x16, int n_img, int h_in, int w_in, int cin, int x_cstride, const void* w16,
                        int n_total, int kh, int kw, int ph, int pw, const float* bias, const float* scale, int relu,
                        float* out, int out_cstride, void* out16, void* stream);
int qeb_conv_wgrad_tc(const float* x, int cin, int x_cstride, int h_in, int w_in, const float* dy, int cout,
                      int dy_cstride, int n_img, int kh, int kw, int ph, int pw, float* dw, void* stream);
/* fp16-operand variant of qeb_conv_wgrad_tc (backward pass): x16 / dy16 = fp16 shadows of x / dy with the same element layout
 * (channel strides multiples of 8); dy16 holds dy * S for a power-of-two S and alpha_dev (device scalar, nullable) = 1 / S
 * multiplies the accumulator. kind::f16 MN-major operands, fp32 accumulation; QEB_FP16_BWD=0 falls back to the tf32 reads. */
int qeb_conv_wgrad_tc16(const float* x, const void* x16, int cin, int x_cstride, int h_in, int w_in, const float* dy,
                        const void* dy16, int cout, int dy_cstride, int n_img, int kh, int kw, int ph, int pw,
                        const float* alpha_dev, float* dw, void* stream);
int qeb_convT2x2_fprop_tc(const float* x, int n_img, int h, int w, int cin, int x_cstride, const float* wpacked,
                          const float* bias, int cout, float* out, int out_cstride, void* stream);
int qeb_convT2x2_dgrad_tc(const float* dy, int n_img, int h, int w, int cout, int dy_cstride, const float* wpacked,
                          int cin, float* dx, int dx_cstride, void* stream);
int qeb_convT2x2_wgrad_tc(const float* x, int cin, int x_cstride, int h, int w, const float* dy, int cout,
                          int dy_cstride, int n_img, float* dw, void* stream);

/* One bidirectional LSTM layer, hidden 256 (nn.LSTM(512,256,2,bidirectional=True), models/model_crnn.py:9,19).
 * gates (T,B,2,1024): x*W_ih^T + b_ih + b_hh on entry (gate order i,f,g,o), overwritten with the activated gates
 * (forward) and then with the gradients at the pre-activations (backward). cells (T,B,2,256); y, dy (T,B,512). */
int qeb_lstm_layer_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* cells, float* y, int T, int B,
                       void* stream);
int qeb_lstm_layer_bwd(float* gates, const float* cells, const float* dy, const float* w_hh_fwd, const float* w_hh_rev,
                       int T, int B, void* stream);

/* MSELoss()(img, ones) of TrainNNPrep._get_loss (train_nn_patch.py:181-184, train_nn_area.py:177-181). */
int qeb_mse_ones_fwd(const float* x, long long n, float* loss, void* stream);
int qeb_mse_ones_bwd(const float* x, long long n, const float* grad_out, float* dx, void* stream);

/* torch.optim.Adam (L2-coupled weight decay; train_nn_patch.py:146-152, train_nn_area.py:149-154) over many tensors in
 * one launch. table (device): n_tensors entries of qeb_adam_table_entry_bytes() = {float* p, const float* g, float* m,
 * float* v, int64 n, int64 chunk0}, chunk0 = exclusive prefix sum of ceil(n/1024); step = count after this update. */
int qeb_adam_table_entry_bytes(void);
int qeb_adam_multi(const void* table, int n_tensors, long long n_chunks, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream);

/* Per-launch profiling: CUDA events around every launch of the library on its stream, with the launch site's
 * algorithmic FLOPs / bytes. report: JSON {"tag": {"launches", "ms", "flops", "bytes"}}, clears the records. While it is
 * on, the module calls keep everything on the caller's stream (no side stream), so a launch's duration is its own.
 * on = 2: one record per (tag, work size). */
void qeb_prof_enable(int on);
/* Debugging aid: per-CTA clock64 stamps (16 int64 per CTA) of the tensor-core fprop / dgrad and weight-gradient kernels
 * (see csrc/conv_tc.cu, scripts/dev_timeline.py, scripts/exp/wgrad_timeline.py); NULL = off. */
void qeb_debug_set_timeline(long long* buf);
int qeb_prof_report(char* buf, int cap);

#ifdef __cplusplus
}
#endif
#endif /* QEB_H */
