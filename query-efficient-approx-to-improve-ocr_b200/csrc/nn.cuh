// Internal (C++) interface between the network engines (engine_crnn.cu, engine_unet.cu) and the kernel files.
// Activations are NHWC fp32 views with arbitrary pixel strides (channel stride 1), so that concat buffers,
// sequence-major (T,B,C) tensors and channel slices are read and written in place.
#pragma once
#include "common.cuh"
#include <functional>
#include <string>

struct Img {
  float* p;
  int n, h, w, c;
  long long sn, sh, sw;  // element strides of the image / row / pixel index
};

// contiguous NHWC image whose pixels are `cstride` floats apart (cstride >= c: a channel slice of a wider buffer)
static inline Img img_nhwc(float* p, int n, int h, int w, int c, int cstride = 0) {
  if (cstride == 0) cstride = c;
  Img i;
  i.p = p; i.n = n; i.h = h; i.w = w; i.c = c;
  i.sw = cstride; i.sh = (long long)w * cstride; i.sn = (long long)h * w * cstride;
  return i;
}
static inline Img img_slice(const Img& a, int c0, int c) {
  Img i = a;
  i.p = a.p + c0; i.c = c;
  return i;
}
static inline long long img_pixels(const Img& a) { return (long long)a.n * a.h * a.w; }
// true when pixel index (n,h,w) -> offset is a single stride (rows/images packed back to back)
static inline bool img_flat(const Img& a) { return a.sh == a.sw * a.w && a.sn == a.sh * a.h; }

// ---- scaled fp16 shadows of gradient tensors (operands of the backward contractions) ------------------------------------------
// A gradient tensor that feeds a dgrad / wgrad contraction is written twice by its producer: the fp32 tensor and an fp16 copy
// of v * S (saturating) for the tensor core - kind::f16 runs at twice the kind::tf32 rate on half the operand bytes with the
// same 11-bit significand. S is a power of two per tensor (exact scaling) taken from the running maximum the producers of the
// PREVIOUS backward call recorded for that tensor ("delayed scaling": S * max ~ 2^4, i.e. 2^12 of headroom before the shadow
// saturates and 2^18 before small values leave fp16's normal range; profiles/r1_notes.md: a scale up to 2^16 off the ideal one
// costs no accuracy). The consumer multiplies its accumulator by 1 / S (TcEpilogue::alpha, WgradShadows::alpha).
struct GradShadow {
  void* out16 = nullptr;          // fp16 copy, same element layout as the fp32 tensor; NULL: not written
  const float* scale = nullptr;   // device scalar S
  unsigned* amax = nullptr;       // device: max |v| as an fp32 bit pattern (red.max.u32), NULL: not tracked
  int only16 = 0;                 // 1 (with out16): every reader takes the shadow - the fp32 store is dropped (bn_bwd_apply_* only)
};
// One set of slots per network instance (keyed on a parameter pointer), living in the library: [amax | S | 1/S] x n.
struct GradScales {
  unsigned* amax = nullptr;
  float* scale = nullptr;
  float* inv = nullptr;
  int n = 0;
  bool valid = false;   // a previous backward call has recorded the maxima these scales come from
  GradShadow slot(int i, void* out16) const {
    GradShadow g;
    if (amax) { g.amax = amax + i; if (valid && out16) { g.out16 = out16; g.scale = scale + i; } }
    return g;
  }
};
// prepare: allocates the slots of (net_kind, key) on first use, unless `st` is being captured (then: no slots = the tf32 path).
// begin: looks them up and enqueues the kernel that turns the previous call's maxima into this call's scales and clears the
// maxima (never allocates: it may run inside a stream capture).
void grad_scales_prepare(int net_kind, const void* key, int n, cudaStream_t st);
int grad_scales_begin(int net_kind, const void* key, int n, cudaStream_t st, GradScales* out);
// host-side peek used for the call-graph key: have these slots seen a backward call yet? (-1: no slots)
int grad_scales_state(int net_kind, const void* key);
void grad_scales_commit(int net_kind, const void* key);   // after the call has been issued: the next call finds valid maxima

// ---- tensor-core (tcgen05, kind::tf32) contractions, conv_tc.cu -------------------------------------------------
struct TcEpilogue {
  const float* bias = nullptr;   // per output channel
  const float* scale = nullptr;  // per output channel, applied before bias: v = acc*scale + bias
  int relu = 0;
  const Img* mask = nullptr;     // v = (mask(n,h,w,c) > 0) ? v : 0   (ReLU backward of the layer below, fused)
  int accumulate = 0;            // out += v
  double* bn_stats = nullptr;    // train-mode BatchNorm of the layer above: per-channel sum (first n_total doubles) and sum of
                                 // squares (next n_total) of the stored values are ADDED to this zero-filled buffer - the
                                 // separate statistics pass over the conv output disappears
  // Input-gradient kernels whose output is the gradient at relu(bn(z)) of the layer below (train-mode BatchNorm): the ReLU
  // mask (z*scale + shift > 0) is applied to the stored gradient and the two reductions of the BatchNorm backward
  // (sum g, sum g*xhat per channel) are ADDED to bn_red - the separate bn_bwd_reduce pass disappears. scsh = [scale | shift |
  // mean | invstd] (4 x n_total floats). *bn_red_fused is set to 1 when the kernel did it (not with split-K / ragged rows).
  const Img* bn_z = nullptr;
  const float* bn_scsh = nullptr;
  double* bn_red = nullptr;
  int* bn_red_fused = nullptr;
  // fp16 operand shadows (forward pass): when BOTH in16 (the input tensor as fp16, same element layout / strides as the
  // fp32 Img) and w16 (the packed weights as fp16, same layout as wpacked) are given, the contraction runs with
  // kind::f16 operands - same 11-bit significand as tf32, round-to-nearest instead of truncation, half the operand bytes.
  // out16: optional fp16 shadow of the output (same element layout as `out`), written by the epilogue for the next layer.
  const void* in16 = nullptr;
  const void* w16 = nullptr;
  void* out16 = nullptr;
  // 1: the fp32 output is itself a tf32 tensor-core operand later (an activation kept for a weight gradient, a gradient that
  // feeds the next dgrad / wgrad directly): it is stored rounded to tf32 (common.cuh qeb_tf32r). Rules out split-K (partial
  // sums cannot be rounded), so the tile is narrowed instead.
  int round_out = 0;
  // Fused CRNN head (models/model_crnn.py:20): the stored values are log_softmax(acc + bias) over the n_total <= 128 output
  // channels of every pixel; argmax (optional, one int per output pixel): first index of the row's largest log-prob
  // (torch.argmax order), what pred_to_string (utils.py:78-89) takes per frame. Plain bias epilogue only.
  int log_softmax = 0;
  int* argmax = nullptr;
  // backward pass with scaled fp16 operands: alpha (device scalar, 1 / S of the in16 operand) multiplies the accumulator before
  // anything else; gs: the OUTPUT is itself an operand of a later contraction - its scaled fp16 shadow / running maximum.
  // *gs_done = 1 when the kernel honoured gs (vector epilogues only: aligned rows, multiples of 32 channels, no split-K)
  const float* alpha = nullptr;
  GradShadow gs;
  int* gs_done = nullptr;
  int gs_cols = 0;   // > 0 (a multiple of 32): only output channels [0, gs_cols) are an operand later - shadow / maximum of those
};
// stride-1 convolution / GEMM. wpacked: [n_total][kh*kw*x.c] (K-major). out.c >= n_total channels are written.
int tc_conv_fprop(const Img& x, const float* wpacked, int n_total, int kh, int kw, int ph, int pw, const Img& out,
                  const TcEpilogue& ep, cudaStream_t st);
// ConvTranspose2d 2x2 stride 2: x (n,h,w,cin) -> out (n,2h,2w,cout). wpacked: [(dh*2+dw)*cout + co][cin]; bias (cout).
int tc_convT_fprop(const Img& x, const float* wpacked, const float* bias, const Img& out, cudaStream_t st,
                   const TcEpilogue* shadows = nullptr);   // shadows: only in16 / w16 / out16 are read
// its input gradient: dy (n,2h,2w,cout) -> dx (n,h,w,cin). wpacked: [cin][(dh*2+dw)*cout + co].
int tc_convT_dgrad(const Img& dy, const float* wpacked, const Img& dx, const TcEpilogue& ep, cudaStream_t st);
// weight gradient, accumulated with atomics: dw[co*s_co + ci*s_ci + ky*s_kh + kx*s_kw] += sum_pix x(pix+tap, ci) dy(pix, co)
// sh (optional): fp16 shadows of both operands (same element layout as x / dy; dy16 = dy * S for a power-of-two S, alpha -> 1 / S
// on the device or NULL) - the contraction then runs with kind::f16 MN-major operands (half the TMA rows and bytes), same
// 11-bit significand as the rounded tf32 reads. QEB_FP16_BWD=0 ignores the shadows.
struct WgradShadows {
  const void* x16 = nullptr;
  const void* dy16 = nullptr;
  const float* alpha = nullptr;
};
int tc_conv_wgrad(const Img& x, const Img& dy, int kh, int kw, int ph, int pw, float* dw, long long s_co, long long s_ci,
                  long long s_kh, long long s_kw, cudaStream_t st, const WgradShadows* sh = nullptr);
// ConvTranspose2d 2x2 s2 weight gradient: dw[ci][co][dh][dw] (torch layout) += sum x(n,h,w,ci) dy(n,2h+dh,2w+dw,co)
int tc_convT_wgrad(const Img& x, const Img& dy, float* dw, cudaStream_t st, const WgradShadows* sh = nullptr);

// ---- weight packing, layout.cu ------------------------------------------------------------------------------------
// generic strided gather: dst[i0*d0 + i1*d1 + i2] = src[i0*s0 + i1*s1 + i2*s2] for i0<n0, i1<n1, i2<n2 (dst innermost
// contiguous); entries of a padded destination are expected to be zero-filled by the caller.
int pack_3d(const float* src, float* dst, int n0, int n1, int n2, long long s0, long long s1, long long s2, long long d0,
            long long d1, cudaStream_t st);
// the same gather for up to kMax tensors in ONE launch (all weight packs of a network pass)
struct PackJob {
  const float* src;
  float* dst;
  int n0, n1, n2;
  long long s0, s1, s2, d0, d1;
  long long start;  // first element of this job in the batch-wide numbering
  int half_out;     // 1: dst is an fp16 buffer (forward operand copies of weights)
};
// conv weight re-layouts go through a tiled transpose (32 x 32 channels x taps through shared memory, coalesced both ways)
struct ConvPackJob {
  const float* src;
  float* dst;
  int A, B, taps;   // torch weight (A = Cout, B = Cin, taps); both multiples of 32 for this path
  int mode;         // 0: dst[a][tap][b] = src[a][b][tap]   fprop B operand
                    // 1: dst[b][taps-1-tap][a] = src[a][b][tap]   dgrad B operand (flipped taps, channels transposed)
                    // 2: dst[a][b][tap] (+)= src[a][tap][b]   packed weight gradient -> torch layout
  int tile_start;   // first 32 x 32 tile of this job in the batch-wide numbering
  int half_out;     // modes 0, 1: dst is an fp16 buffer
};
struct PackBatch {
  enum { kMax = 28 };
  PackJob jobs[kMax];
  ConvPackJob cjobs[kMax];
  int n = 0, nc = 0, ctiles = 0;
  long long total = 0;
  int accumulate = 0;  // 1: dst += src for every job of the batch
  void add(const float* src, float* dst, int n0, int n1, int n2, long long s0, long long s1, long long s2, long long d0,
           long long d1) {
    PackJob& j = jobs[n++];
    j.src = src; j.dst = dst; j.n0 = n0; j.n1 = n1; j.n2 = n2; j.s0 = s0; j.s1 = s1; j.s2 = s2; j.d0 = d0; j.d1 = d1;
    j.start = total;
    j.half_out = 0;
    total += (long long)n0 * n1 * n2;
  }
  void last_to_half() { if (n > 0) jobs[n - 1].half_out = 1; }   // the job just added writes fp16 (dst cast from a half buffer)
  bool add_conv(const float* src, float* dst, int A, int B, int taps, int mode) {
    if (A % 32 || B % 32 || taps > 9) return false;
    ConvPackJob& j = cjobs[nc++];
    j.src = src; j.dst = dst; j.A = A; j.B = B; j.taps = taps; j.mode = mode; j.tile_start = ctiles; j.half_out = 0;
    ctiles += (A / 32) * (B / 32);
    return true;
  }
  // torch Conv2d weight (Cout,Cin,taps) -> fprop B operand [Cout][tap][Cin]
  void add_fprop(const float* w, float* dst, int cout, int cin, int taps) {
    if (!add_conv(w, dst, cout, cin, taps, 0)) add(w, dst, cout, taps, cin, (long long)cin * taps, 1, taps, (long long)taps * cin, cin);
  }
  // the same B operand as fp16 (dst16: a buffer of cout*taps*cin halves)
  void add_fprop16(const float* w, void* dst16, int cout, int cin, int taps) {
    float* d = static_cast<float*>(dst16);
    if (add_conv(w, d, cout, cin, taps, 0)) cjobs[nc - 1].half_out = 1;
    else { add(w, d, cout, taps, cin, (long long)cin * taps, 1, taps, (long long)taps * cin, cin); last_to_half(); }
  }
  // -> dgrad B operand [Cin][flipped tap][Cout] (source taps walked backwards with a negative stride)
  void add_dgrad(const float* w, float* dst, int cout, int cin, int taps) {
    if (!add_conv(w, dst, cout, cin, taps, 1))
      add(w + (taps - 1), dst, cin, taps, cout, taps, -1, (long long)cin * taps, (long long)taps * cout, cout);
  }
  // the dgrad B operand as fp16 (dst16: a buffer of cin*taps*cout halves)
  void add_dgrad16(const float* w, void* dst16, int cout, int cin, int taps) {
    float* d = static_cast<float*>(dst16);
    if (add_conv(w, d, cout, cin, taps, 1)) cjobs[nc - 1].half_out = 1;
    else { add(w + (taps - 1), d, cin, taps, cout, taps, -1, (long long)cin * taps, (long long)taps * cout, cout); last_to_half(); }
  }
  void add_copy(const float* src, float* dst, long long n) { add(src, dst, 1, 1, (int)n, 0, 0, 1, 0, 0); }
  // packed weight gradient [Cout][tap][Cin] (what the wgrad kernel accumulates into with coalesced atomics) -> torch
  // layout (Cout,Cin,taps)
  void add_unpack_grad(const float* packed, float* dw, int cout, int cin, int taps) {
    if (!add_conv(packed, dw, cout, cin, taps, 2)) add(packed, dw, cout, cin, taps, (long long)taps * cin, 1, cin, (long long)cin * taps, taps);
  }
};
int pack_flush(PackBatch& b, cudaStream_t st);

// ---- direct (SIMT) kernels, nn_ops.cu ---------------------------------------------------------------------------
// 3x3 pad-1 convolution with ONE input channel: x (n,h,w,1) -> out (n,h,w,cout); w torch layout (cout,1,3,3).
int c1_conv_fwd(const Img& x, const float* w, const float* bias, int relu, const Img& out, cudaStream_t st);
// Gaussian jitter fused into the input load (nn_ops.cu c1_jitter_fwd_kernel): x (n,h,w) contiguous is the CLEAN image; the
// convolution sees clamp(x - coef * N(mean, sigma[n]), 0, 1), which is also written to noisy_out (and the noise to noise_out)
struct JitterArgs {
  const float* sigma = nullptr;                 // one std per image (device)
  float mean = 0.f, coef = 1.f;
  unsigned long long seed = 0;
  const unsigned long long* seed_dev = nullptr;  // optional device-resident key added to `seed` when the kernel runs
  float* noisy_out = nullptr;                   // (n,h,w): the jittered image (OCR hand-off, conv1 weight gradient)
  float* noise_out = nullptr;                   // (n,h,w) or NULL
};
int c1_conv_fwd_jitter(const float* x, int n_img, int h, int w, const JitterArgs& j, const float* wgt, const float* bias, int relu,
                       const Img& out, cudaStream_t st);
int c1_conv_wgrad(const Img& x, const Img& dy, float* dw, float* dbias, cudaStream_t st);  // accumulates
int c1_conv_dgrad(const Img& dy, const float* w, const Img& dx, cudaStream_t st);          // overwrites dx
// 1x1 convolution to ONE output channel + sigmoid: y = sigmoid(sum_c x*w[c] + b)
int o1_conv_sigmoid_fwd(const Img& x, const float* w, const float* b, float* y, cudaStream_t st);
// dz = dy*y*(1-y); dx = dz*w[c]; dw[c] += sum dz*x; db += sum dz
int o1_conv_sigmoid_bwd(const Img& x, const float* w, const float* y, const float* dy, const Img& dx, float* dw, float* db,
                        cudaStream_t st);
// max pooling (ph x pw window = stride), NHWC
int maxpool_fwd(const Img& x, int ph, int pw, const Img& out, cudaStream_t st, void* out16 = nullptr);  // out16: fp16 shadow
// dx = routed dy (first maximal element wins) * chan_scale[c] (if given), zeroed where x <= 0 when relu_mask, plus `add`
// (same shape as x) if given
// bn_z / bn_scsh / bn_red (all or none): dx is the gradient at the output of a train-mode conv + BN + ReLU unit with
// pre-activation bn_z; that unit's BatchNorm-backward reductions are accumulated into bn_red[2C] in the same pass
int maxpool_bwd(const Img& x, const Img& dy, int ph, int pw, int relu_mask, const float* chan_scale, const Img* add,
                const Img& dx, cudaStream_t st, const Img* bn_z = nullptr, const float* bn_scsh = nullptr, double* bn_red = nullptr,
                const GradShadow* gs = nullptr);   // gs: dx is a dgrad / wgrad operand
// batch norm over all pixels of z (train: batch statistics; eval: running statistics)
struct BnParams {
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float eps, momentum;
};
// stats: [0,c) sum, [c,2c) sumsq (double accumulators); scsh: [0,c) scale, [c,2c) shift, then train: [2c,3c) mean,
// [3c,4c) invstd; eval: [2c,3c) beta, [3c,4c) 1/gamma (xhat is recovered from the layer output)
int bn_train_stats(const Img& z, double* stats, cudaStream_t st);
int bn_train_finalize(const double* stats, long long count, int c, const BnParams& bn, float* scsh, cudaStream_t st);
// finalize + apply fused (train mode): out = relu?(bn(z)) from the batch sums; writes scsh, updates the running statistics
// only16 (with out16): every later reader takes the fp16 shadow - the fp32 store is dropped
int bn_train_finalize_apply(const Img& z, const double* stats, const BnParams& bn, float* scsh, int relu, const Img& out,
                            cudaStream_t st, void* out16 = nullptr, int only16 = 0);
// the same (ReLU on) fused with the 2 x 2 max pooling of the unit's output: writes out (+ out16) and pool (+ pool16)
int bn_train_finalize_apply_pool(const Img& z, const double* stats, const BnParams& bn, float* scsh, const Img& out, void* out16,
                                 const Img& pool, void* pool16, cudaStream_t st);   // out16: fp16 shadow of out (same element layout)
int bn_eval_scsh(int c, const BnParams& bn, const float* conv_bias, float* scsh, cudaStream_t st);
// out = relu?(z*scale + shift)
int bn_apply(const Img& z, const float* scsh, int relu, const Img& out, cudaStream_t st, void* out16 = nullptr);
// backward through relu(bn(z)): red: [0,c) sum g, [c,2c) sum g*xhat (double), g = dy * (z*scale+shift > 0 if relu)
int bn_bwd_reduce(const Img& z, const Img& dy, const float* scsh, int relu, double* red, cudaStream_t st);
// train: dz = gamma*invstd*(g - sum_g/M - xhat*sum_gx/M); dgamma += sum_gx, dbeta += sum_g (when given)
int bn_bwd_apply_train(const Img& z, const Img& dy, const float* scsh, int relu, const double* red, const float* gamma,
                       const Img& dz, float* dgamma, float* dbeta, cudaStream_t st, const GradShadow* gs = nullptr);
// eval (frozen statistics): dz = g*scale; relu = 2: `z` is the layer's ReLU OUTPUT (the pre-activation is not kept)
// red (from bn_bwd_reduce on the same tensors) is only needed for dgamma += sum g*xhat, dbeta += sum g.
int bn_bwd_apply_eval(const Img& z, const Img& dy, const float* scsh, int relu, const double* red, const Img& dz,
                      float* dgamma, float* dbeta, cudaStream_t st, const GradShadow* gs = nullptr);
// out[c] += sum over pixels of x(pix, c)   (bias gradients)
int colsum_acc(const Img& x, float* out, cudaStream_t st, float* out2 = nullptr);   // out2: a second accumulator receiving the same sums
// dx = dy * (a > 0)
int relu_bwd(const Img& a, const Img& dy, const Img& dx, cudaStream_t st);
int fill_zero(void* p, size_t bytes, cudaStream_t st);
// out = a + b (b may be null)
int vec_add(const float* a, const float* b, float* out, int n, cudaStream_t st);

long long* qeb_debug_timeline();  // conv_tc.cu: buffer set by qeb_debug_set_timeline, or NULL

// ---- LSTM recurrence, lstm.cu ------------------------------------------------------------------------------------
// One bidirectional layer, hidden size 256. gates: (T,B,2,4*256) holding x*W_ih^T + b_ih + b_hh on entry (gate order
// i,f,g,o as torch), overwritten with the ACTIVATED gates; w_hh[dir]: torch layout (1024,256); cells: (T,B,2,256) c_t;
// y: (T,B,512) [forward | reverse].
int lstm_layer_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* cells, float* y, int T, int B,
                   cudaStream_t st, void* y16 = nullptr, int round_io = 0);   // round_io: y stored rounded to tf32 (see lstm.cu)
// dy: (T,B,512). gates (activated) are overwritten with the gradients at the pre-activations (T,B,2,1024).
int lstm_layer_bwd(float* gates, const float* cells, const float* dy, const float* w_hh_fwd, const float* w_hh_rev, int T,
                   int B, cudaStream_t st, int round_io = 0);

// ---- per-call CUDA graphs, abi.cu --------------------------------------------------------------------------------------
// An engine call (one network forward or backward) is 30-90 kernel launches issued from one C call; at ~3 us of host time per
// cudaLaunchKernelEx an eagerly launched training step is bound by the HOST (4.3 ms issued for 3.5 ms of GPU work). The
// launch sequence of a call is a pure function of its arguments - shapes, flags and POINTERS, and the caching allocator hands
// the same workspace / gradient / input blocks back step after step - so the call is captured once per distinct argument
// set (on an internal stream: the caller's stream may be the legacy default stream, which cannot be captured) and later
// calls with the same arguments replay the instantiated graph on the caller's stream: one cudaGraphLaunch. Bypassed while
// the caller's stream is itself being captured (GraphedStep), under qeb_prof_enable / QEB_DBG_SKIP, and with
// QEB_CALL_GRAPHS=0. `key` must hold EVERYTHING the body's launches depend on.
struct CallKey {
  std::string bytes;
  void raw(const void* p, size_t n) { bytes.append(static_cast<const char*>(p), n); }
  template <typename T>
  CallKey& add(const T& v) { raw(&v, sizeof(T)); return *this; }
  CallKey& ptrs(const void* const* a, int n) {   // the VALUES of an array of pointers (NULL array: nothing)
    if (a) raw(a, sizeof(void*) * (size_t)n);
    else add(n);
    return *this;
  }
};
int qeb_run_cached(const CallKey& key, cudaStream_t st, const std::function<int(cudaStream_t)>& body);

// ---- side stream, abi.cu ----------------------------------------------------------------------------------------
// Weight and bias gradients are off the backward's critical path (only the input gradients feed the next layer), so the
// engines enqueue them on a second stream and let them fill the tail waves of the dgrad / BatchNorm kernels.
// fork(): the side stream waits for everything enqueued on the main stream so far; join(): the reverse. The stream and
// its two events are created lazily per host thread and device and live until the process ends; QEB_SIDE_STREAM=0
// keeps everything on the caller's stream.
struct SideStream {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr, mark_ev = nullptr, mark2_ev = nullptr;
  bool enabled = false, dirty = false, marked = false, marked2 = false;
  int init(cudaStream_t main_stream);
  int fork();
  int join();
  // mark(): remember the side stream's current position; wait_mark(): the main stream waits for that position only (not
  // for side work queued after it). Used for the weight re-layouts, which run beside the first kernels of a pass.
  int mark();
  int wait_mark();
  // a second, independent mark: the re-layouts of the weights a pass needs late (the deep layers - most of the bytes) are queued
  // behind those it needs at once and waited for where they are first read
  int mark2();
  int wait_mark2();
  cudaStream_t s() const { return enabled ? side : main; }
};
