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
 * in CRNN.backward_hook, models/model_crnn.py:30-32). grad_out: (B) for none, (1) otherwise. With batch_index only
 * the listed columns of grad are written. */
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

/* ---- minibatch-subset selection ---------------------------------------------------------------------------------
 * TopKCERSampler.query selection_utils.py:144-151: per segment the indices of the k largest fp32 CERs, descending,
 * equal values lowest-index-first. CerRangeSampler.query selection_utils.py:107-135: per segment k points
 * (max-min)*rand+min, then sequentially the first arg-min of |point - copy| with copy[idx]=100.
 * vals: all segments concatenated; seg_off (n_seg+1); seg_k (n_seg); out_off (n_seg): exclusive prefix sum of the
 * picks per segment (min(k,n) for topk, k for range); out_idx: int64, local to the segment. rands: torch.rand draws
 * made on the host, concatenated like out_idx. work: scratch, same size as vals. points_out: optional. */
int qeb_cer_topk_segmented(const float* vals, const int* seg_off, const int* seg_k, const int* out_off, int n_seg,
                           long long* out_idx, void* stream);
int qeb_cer_range_segmented(const float* vals, const int* seg_off, const int* seg_k, const int* out_off,
                            const float* rands, int n_seg, float* work, long long* out_idx, float* points_out,
                            void* stream);

/* ---- Gaussian jitter: AddGaussianNoice.__call__ transform_helper.py:33-45; add_noise train_nn_patch.py:187-191,
 * train_nn_area.py:184-191. out = clamp(img - coef * noise, 0, 1); noise = noise_in if given, else
 * mean + sigma[image] * N(0,1) from Philox4x32-10(seed). img/out/noise: (n_img, hw) fp32, hw % 4 == 0, 16B aligned. */
int qeb_gauss_jitter(const float* img, const float* sigma, float mean, float coef, const float* noise_in,
                     unsigned long long seed, long long n_img, int hw, float* out, float* noise_out, void* stream);

/* ---- crop + centre-pad with 1.0: get_text_stack / padder utils.py:118-141 -------------------------------------
 * img (H,W) fp32; boxes (n,4) int32 x_min,y_min,x_max,y_max; out (n,oh,ow). scatter is the adjoint (atomic adds into
 * a zero-initialised gimg (H,W)). */
int qeb_crop_pad_gather(const float* img, int H, int W, const int* boxes, int n, int oh, int ow, float* out,
                        void* stream);
int qeb_crop_pad_scatter(const float* gout, int H, int W, const int* boxes, int n, int oh, int ow, float* gimg,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QEB_H */
