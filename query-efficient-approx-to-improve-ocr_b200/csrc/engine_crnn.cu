// CRNN surrogate, forward and backward, as one host-side launch sequence over the sm_100a kernels.
// Reference: models/model_crnn.py:5-56 (Convolutional.forward :47-56, CRNN.forward :16-21, map_to_sequence :23-28).
//
//   x (B,1,32,W) --conv1+ReLU--> pool 2x2 --conv2+ReLU--> pool 2x2 --conv3+ReLU--> conv4+ReLU --> pool (2,1)
//     --conv5+BN1+ReLU--> conv6+BN2+ReLU --> pool (2,1) --conv7 (2x2, no pad)--> (T = W/4-1, B, 512)
//     --BiLSTM x2 (hidden 256)--> Linear(512, V) --> logits (T,B,V)          [log_softmax is its own op, ctc.cu]
//
// Activations are NHWC fp32 in a caller-owned workspace (layout: CrnnPlan below); the module input (B,1,32,W) is
// already NHWC because it has one channel. conv2..conv7, the LSTM input projections and the Linear layer run on
// tcgen05 (conv_tc.cu); conv1, pooling, batch norm and the bias gradients are direct kernels (nn_ops.cu); the
// recurrence is the cluster kernel of lstm.cu. BatchNorm follows the module mode: batch statistics + running-stat
// update (phase A, train_nn_patch.py:226) or frozen running statistics folded into the conv epilogue (phase B after
// set_bn_eval, utils.py:113-115).
#include "nn.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

enum {  // parameter order of the C ABI (= state_dict order of the reference module)
  P_C1W, P_C1B, P_C2W, P_C2B, P_C3W, P_C3B, P_C4W, P_C4B, P_C5W, P_C5B, P_BN1W, P_BN1B, P_C6W, P_C6B, P_BN2W, P_BN2B,
  P_C7W, P_C7B, P_LSTM0,  // + (layer*2 + dir)*4 + {w_ih, w_hh, b_ih, b_hh}
  P_LINW = P_LSTM0 + 16, P_LINB, P_COUNT
};
enum { B_BN1_MEAN, B_BN1_VAR, B_BN1_NBT, B_BN2_MEAN, B_BN2_VAR, B_BN2_NBT, B_COUNT };

constexpr int kHid = 256;

struct Arena {
  char* base;
  size_t off = 0;
  explicit Arena(void* b) : base(static_cast<char*>(b)) {}
  float* take(size_t n_floats) {
    float* p = reinterpret_cast<float*>(base + off);
    off += (n_floats * sizeof(float) + 255) & ~size_t(255);
    return p;
  }
};

struct CrnnPlan {
  int B, W, V, T, W2, W4;
  // saved by the forward for the backward
  float *a1f, *a1, *a2f, *a2, *a3, *a4f, *a4, *z5, *a5, *z6, *a6f, *a6, *x0, *g0, *c0, *y0, *g1, *c1, *y1;
  float *scsh5, *scsh6;  // BN scale/shift/mean/invstd (4*512 each)
  double* bnstats;       // 2 layers x 2*512
  // packed weights
  float *wp2, *wp3, *wp4, *wp5, *wp6, *wp7, *wih0, *wih1, *bias0, *bias1;
  // backward scratch
  float *dlp, *wlinT, *wihT0, *wihT1, *wpd2, *wpd3, *wpd4, *wpd5, *wpd6, *wpd7, *dy1, *dy0, *dx0, *d6, *d6f, *d5, *d4, *d4f, *d3, *d2, *d2f, *d1, *d1f;
  double* bnred;
  // fp16 operand shadows of the forward pass (kind::f16 tensor-core path): activations that feed a contraction, and the
  // forward B operands. Typed void*: only the kernels look inside.
  void *a1h, *a2h, *a3h, *a4h, *a5h, *a6h, *x0h, *y0h, *y1h;
  void *wp2h, *wp3h, *wp4h, *wp5h, *wp6h, *wp7h, *wih0h, *wih1h, *wlinh;
  float *dwp2, *dwp3, *dwp4, *dwp5, *dwp6, *dwp7;  // packed [Cout][tap][Cin] weight-gradient accumulators, contiguous
  size_t dwp_bytes;
  // backward pass with scaled fp16 operands (nn.cuh GradShadow): shadows of the conv stack's gradient tensors, fp16 dgrad B operands
  void *dx0h, *d6fh, *d5h, *d4fh, *d3h, *d2fh;
  void *wpd2h, *wpd3h, *wpd4h, *wpd5h, *wpd6h, *wpd7h;
  size_t bytes;
};

CrnnPlan make_plan(int B, int W, int V, void* base) {
  CrnnPlan p;
  p.B = B; p.W = W; p.V = V; p.W2 = W / 2; p.W4 = W / 4; p.T = W / 4 - 1;
  Arena a(base);
  const size_t px32 = (size_t)B * 32 * W, px16 = (size_t)B * 16 * p.W2, px8 = (size_t)B * 8 * p.W4, px4 = (size_t)B * 4 * p.W4,
               px2 = (size_t)B * 2 * p.W4, tb = (size_t)p.T * B;
  p.a1f = a.take(px32 * 64); p.a1 = a.take(px16 * 64);
  p.a2f = a.take(px16 * 128); p.a2 = a.take(px8 * 128);
  p.a3 = a.take(px8 * 256);
  p.a4f = a.take(px8 * 256); p.a4 = a.take(px4 * 256);
  p.z5 = a.take(px4 * 512); p.a5 = a.take(px4 * 512);
  p.z6 = a.take(px4 * 512); p.a6f = a.take(px4 * 512); p.a6 = a.take(px2 * 512);
  p.x0 = a.take(tb * 512);
  p.g0 = a.take(tb * 2048); p.c0 = a.take(tb * 512); p.y0 = a.take(tb * 512);
  p.g1 = a.take(tb * 2048); p.c1 = a.take(tb * 512); p.y1 = a.take(tb * 512);
  p.scsh5 = a.take(4 * 512); p.scsh6 = a.take(4 * 512);
  p.bnstats = reinterpret_cast<double*>(a.take(2 * 2 * 512 * 2));
  p.wp2 = a.take((size_t)128 * 9 * 64); p.wp3 = a.take((size_t)256 * 9 * 128); p.wp4 = a.take((size_t)256 * 9 * 256);
  p.wp5 = a.take((size_t)512 * 9 * 256); p.wp6 = a.take((size_t)512 * 9 * 512); p.wp7 = a.take((size_t)512 * 4 * 512);
  p.wih0 = a.take((size_t)2048 * 512); p.wih1 = a.take((size_t)2048 * 512);
  p.bias0 = a.take(2048); p.bias1 = a.take(2048);
  p.dlp = a.take(tb * 96); p.wlinT = a.take((size_t)512 * 96);
  p.wihT0 = a.take((size_t)512 * 2048); p.wihT1 = a.take((size_t)512 * 2048);
  p.wpd2 = a.take((size_t)128 * 9 * 64); p.wpd3 = a.take((size_t)256 * 9 * 128); p.wpd4 = a.take((size_t)256 * 9 * 256);
  p.wpd5 = a.take((size_t)512 * 9 * 256); p.wpd6 = a.take((size_t)512 * 9 * 512); p.wpd7 = a.take((size_t)512 * 4 * 512);
  p.dy1 = a.take(tb * 512); p.dy0 = a.take(tb * 512); p.dx0 = a.take(tb * 512);
  p.d6 = a.take(px2 * 512); p.d6f = a.take(px4 * 512); p.d5 = a.take(px4 * 512);
  p.d4 = a.take(px4 * 256); p.d4f = a.take(px8 * 256); p.d3 = a.take(px8 * 256);
  p.d2 = a.take(px8 * 128); p.d2f = a.take(px16 * 128);
  p.d1 = a.take(px16 * 64); p.d1f = a.take(px32 * 64);
  p.bnred = reinterpret_cast<double*>(a.take(2 * 2 * 512 * 2));  // 2 layers x (sum g, sum g*xhat) x 512 doubles
  {
    const size_t o0 = a.off;
    p.dwp2 = a.take((size_t)128 * 9 * 64); p.dwp3 = a.take((size_t)256 * 9 * 128); p.dwp4 = a.take((size_t)256 * 9 * 256);
    p.dwp5 = a.take((size_t)512 * 9 * 256); p.dwp6 = a.take((size_t)512 * 9 * 512); p.dwp7 = a.take((size_t)512 * 4 * 512);
    p.dwp_bytes = a.off - o0;
  }
  {
    auto half = [&](size_t n) { return static_cast<void*>(a.take((n + 1) / 2)); };
    p.a1h = half(px16 * 64); p.a2h = half(px8 * 128); p.a3h = half(px8 * 256); p.a4h = half(px4 * 256);
    p.a5h = half(px4 * 512); p.a6h = half(px2 * 512); p.x0h = half(tb * 512); p.y0h = half(tb * 512); p.y1h = half(tb * 512);
    p.wp2h = half((size_t)128 * 9 * 64); p.wp3h = half((size_t)256 * 9 * 128); p.wp4h = half((size_t)256 * 9 * 256);
    p.wp5h = half((size_t)512 * 9 * 256); p.wp6h = half((size_t)512 * 9 * 512); p.wp7h = half((size_t)512 * 4 * 512);
    p.wih0h = half((size_t)2048 * 512); p.wih1h = half((size_t)2048 * 512); p.wlinh = half((size_t)96 * 512);
    p.dx0h = half(tb * 512); p.d6fh = half(px4 * 512); p.d5h = half(px4 * 512); p.d4fh = half(px8 * 256); p.d3h = half(px8 * 256);
    p.d2fh = half(px16 * 128);
    p.wpd2h = half((size_t)128 * 9 * 64); p.wpd3h = half((size_t)256 * 9 * 128); p.wpd4h = half((size_t)256 * 9 * 256);
    p.wpd5h = half((size_t)512 * 9 * 256); p.wpd6h = half((size_t)512 * 9 * 512); p.wpd7h = half((size_t)512 * 4 * 512);
  }
  p.bytes = a.off;
  return p;
}

// forward contractions with fp16 operands (default) or tf32 operands read from the fp32 tensors (QEB_FP16_FWD=0)
bool fp16_fwd() {
  static const bool on = !(getenv("QEB_FP16_FWD") && atoi(getenv("QEB_FP16_FWD")) == 0);
  return on;
}

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != QEB_OK) return _rc; \
  } while (0)

BnParams bn_of(const float* const* params, void* const* buffers, int pw, int pb, int bm) {
  BnParams b;
  b.gamma = params[pw]; b.beta = params[pb];
  b.running_mean = static_cast<float*>(buffers[bm]);
  b.running_var = static_cast<float*>(buffers[bm + 1]);
  b.num_batches_tracked = static_cast<long long*>(buffers[bm + 2]);
  b.eps = 1e-5f; b.momentum = 0.1f;
  return b;
}

}  // namespace

QEB_API size_t qeb_crnn_workspace_bytes(int B, int W, int V) {
  if (B <= 0 || W < 8 || W % 4 != 0 || V <= 0 || V > 96) return 0;
  return make_plan(B, W, V, nullptr).bytes;
}

QEB_API int qeb_crnn_num_params(void) { return P_COUNT; }

// x: (B,1,32,W) fp32. params: P_COUNT device pointers in state_dict order; buffers: BN running_mean/var (fp32) and
// num_batches_tracked (int64) for batchnorm1, batchnorm2. bn_train: 1 = batch statistics (+ running-stat update),
// 0 = running statistics. logits: (T,B,V) dense. ws: qeb_crnn_workspace_bytes(), kept untouched until the backward.
namespace {
struct CrnnFwdOpts {
  const JitterArgs* jitter = nullptr;   // non-NULL: x is the CLEAN image, the Gaussian jitter rides the conv1 input load
  int log_softmax = 0;                  // 1: `logits` receives log_softmax(logits) straight from the Linear epilogue
  int* argmax = nullptr;                // (T,B) per-frame arg-max of the log-probs (with log_softmax), nullable
};
int crnn_forward_body(const float* x, int B, int W, int V, const float* const* params, void* const* buffers, int bn_train,
                      void* ws, float* logits, const CrnnFwdOpts& opt, void* stream);
// one captured graph per distinct argument set (nn.cuh qeb_run_cached); a host-seeded jitter changes its key every call: plain launches
int crnn_forward_impl(const float* x, int B, int W, int V, const float* const* params, void* const* buffers, int bn_train,
                      void* ws, float* logits, const CrnnFwdOpts& opt, void* stream) {
  QEB_REQUIRE(x && params && buffers && ws && logits, "crnn_forward: null pointer");
  if (opt.jitter && !opt.jitter->seed_dev)
    return crnn_forward_body(x, B, W, V, params, buffers, bn_train, ws, logits, opt, stream);
  CallKey key;
  key.add(1).add(x).add(B).add(W).add(V).add(bn_train).add(ws).add(logits).add(opt.log_softmax).add(opt.argmax);
  key.ptrs(reinterpret_cast<const void* const*>(params), P_COUNT).ptrs(reinterpret_cast<const void* const*>(buffers), B_COUNT);
  if (opt.jitter) {
    const JitterArgs& j = *opt.jitter;
    key.add(j.sigma).add(j.mean).add(j.coef).add(j.seed).add(j.seed_dev).add(j.noisy_out).add(j.noise_out);
  } else {
    key.add(0);
  }
  return qeb_run_cached(key, (cudaStream_t)stream, [&](cudaStream_t st) {
    return crnn_forward_body(x, B, W, V, params, buffers, bn_train, ws, logits, opt, (void*)st);
  });
}
}  // namespace

QEB_API int qeb_crnn_forward(const float* x, int B, int W, int V, const float* const* params, void* const* buffers,
                             int bn_train, void* ws, float* logits, void* stream) {
  return crnn_forward_impl(x, B, W, V, params, buffers, bn_train, ws, logits, CrnnFwdOpts(), stream);
}

// The same forward with the two north_star fusions switched on per argument:
//  - log_softmax = 1: `out` receives fn.log_softmax(self.linear(x), 2) (models/model_crnn.py:20) computed in the Linear
//    GEMM's epilogue (no logits round trip, no separate launch); argmax_path (T*B ints, nullable) receives the per-frame
//    arg-max class that pred_to_string takes (utils.py:78-89) - qeb_greedy_collapse turns it into strings' class indices;
//  - jit_sigma != NULL: x is the CLEAN (B,1,32,W) batch and the network sees clamp(x - jit_coef * N(jit_mean, jit_sigma[b]),
//    0, 1) (AddGaussianNoice, transform_helper.py:33-45), generated inside conv1's input load with the Philox stream of
//    qeb_gauss_jitter (key = jit_seed + *jit_seed_dev). noisy_out (B,1,32,W, required) receives that image - the OCR engine
//    and conv1's weight gradient need it: pass IT as `x` to qeb_crnn_backward - noise_out (nullable) the noise.
QEB_API int qeb_crnn_forward_fused(const float* x, int B, int W, int V, const float* const* params, void* const* buffers,
                                   int bn_train, void* ws, float* out, int log_softmax, int* argmax_path,
                                   const float* jit_sigma, float jit_mean, float jit_coef, unsigned long long jit_seed,
                                   const unsigned long long* jit_seed_dev, float* noisy_out, float* noise_out, void* stream) {
  CrnnFwdOpts opt;
  JitterArgs j;
  if (jit_sigma) {
    QEB_REQUIRE(noisy_out, "crnn_forward_fused: noisy_out is required with jitter (the backward pass reads it)");
    j.sigma = jit_sigma; j.mean = jit_mean; j.coef = jit_coef; j.seed = jit_seed; j.seed_dev = jit_seed_dev;
    j.noisy_out = noisy_out; j.noise_out = noise_out;
    opt.jitter = &j;
  }
  opt.log_softmax = log_softmax;
  opt.argmax = argmax_path;
  QEB_REQUIRE(log_softmax || !argmax_path, "crnn_forward_fused: argmax_path needs log_softmax = 1");
  return crnn_forward_impl(x, B, W, V, params, buffers, bn_train, ws, out, opt, stream);
}

namespace {
int crnn_forward_body(const float* x, int B, int W, int V, const float* const* params, void* const* buffers, int bn_train,
                      void* ws, float* logits, const CrnnFwdOpts& opt, void* stream) {
  QEB_REQUIRE(x && params && buffers && ws && logits, "crnn_forward: null pointer");
  QEB_REQUIRE(B > 0 && W >= 8 && W % 4 == 0 && V > 0 && V <= 96, "crnn_forward: B=%d W=%d V=%d unsupported", B, W, V);
  QEB_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)x & 15) == 0, "crnn_forward: workspace/input alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const CrnnPlan p = make_plan(B, W, V, ws);
  const int T = p.T;
  Img X = img_nhwc(const_cast<float*>(x), B, 32, W, 1);
  Img A1f = img_nhwc(p.a1f, B, 32, W, 64), A1 = img_nhwc(p.a1, B, 16, p.W2, 64);
  Img A2f = img_nhwc(p.a2f, B, 16, p.W2, 128), A2 = img_nhwc(p.a2, B, 8, p.W4, 128);
  Img A3 = img_nhwc(p.a3, B, 8, p.W4, 256);
  Img A4f = img_nhwc(p.a4f, B, 8, p.W4, 256), A4 = img_nhwc(p.a4, B, 4, p.W4, 256);
  Img Z5 = img_nhwc(p.z5, B, 4, p.W4, 512), A5 = img_nhwc(p.a5, B, 4, p.W4, 512);
  Img Z6 = img_nhwc(p.z6, B, 4, p.W4, 512), A6f = img_nhwc(p.a6f, B, 4, p.W4, 512), A6 = img_nhwc(p.a6, B, 2, p.W4, 512);

  const bool h = fp16_fwd();
  SideStream ss;
  TRY(ss.init(st));
  // the weight re-layouts of this pass in two launches: conv2-conv4 (10 % of the bytes), needed at once, and the rest - first read
  // by conv5, ~100 us into the pass - behind them (QEB_PACK_SPLIT=0: one launch, one wait)
  static const bool split = !(getenv("QEB_PACK_SPLIT") && atoi(getenv("QEB_PACK_SPLIT")) == 0);
  PackBatch pk2;
  {
    PackBatch pk;
    if (h) {   // fp16 B operands; the fp32 packs are not needed by the forward pass then
      PackBatch& late = split ? pk2 : pk;
      pk.add_fprop16(params[P_C2W], p.wp2h, 128, 64, 9);
      pk.add_fprop16(params[P_C3W], p.wp3h, 256, 128, 9);
      pk.add_fprop16(params[P_C4W], p.wp4h, 256, 256, 9);
      late.add_fprop16(params[P_C5W], p.wp5h, 512, 256, 9);
      late.add_fprop16(params[P_C6W], p.wp6h, 512, 512, 9);
      late.add_fprop16(params[P_C7W], p.wp7h, 512, 512, 4);
      for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d) {
          __half* dst = static_cast<__half*>(l ? p.wih1h : p.wih0h) + (size_t)d * 1024 * 512;
          late.add_copy(params[P_LSTM0 + l * 8 + d * 4], reinterpret_cast<float*>(dst), (long long)1024 * 512);
          late.last_to_half();
        }
      late.add_copy(params[P_LINW], static_cast<float*>(p.wlinh), (long long)V * 512);
      late.last_to_half();
    } else {
      pk.add_fprop(params[P_C2W], p.wp2, 128, 64, 9);
      pk.add_fprop(params[P_C3W], p.wp3, 256, 128, 9);
      pk.add_fprop(params[P_C4W], p.wp4, 256, 256, 9);
      pk.add_fprop(params[P_C5W], p.wp5, 512, 256, 9);
      pk.add_fprop(params[P_C6W], p.wp6, 512, 512, 9);
      pk.add_fprop(params[P_C7W], p.wp7, 512, 512, 4);
      for (int l = 0; l < 2; ++l)
        for (int d = 0; d < 2; ++d)
          pk.add_copy(params[P_LSTM0 + l * 8 + d * 4], (l ? p.wih1 : p.wih0) + (size_t)d * 1024 * 512, (long long)1024 * 512);
    }
    TRY(ss.fork());   // beside conv1 (direct kernel, no packed weights) and its pooling
    TRY(pack_flush(pk, ss.s()));
    if (pk2.n + pk2.nc > 0) TRY(ss.mark());
  }
  const bool two_marks = pk2.n + pk2.nc > 0;
  if (two_marks) TRY(pack_flush(pk2, ss.s()));
  // the other parameter-only preparations ride along: summed LSTM biases, folded frozen-BatchNorm scale / shift
  const BnParams bn1 = bn_of(params, buffers, P_BN1W, P_BN1B, B_BN1_MEAN), bn2 = bn_of(params, buffers, P_BN2W, P_BN2B, B_BN2_MEAN);
  for (int l = 0; l < 2; ++l)
    for (int d = 0; d < 2; ++d)
      TRY(vec_add(params[P_LSTM0 + l * 8 + d * 4 + 2], params[P_LSTM0 + l * 8 + d * 4 + 3], (l ? p.bias1 : p.bias0) + d * 1024, 1024, ss.s()));
  if (!bn_train) {
    TRY(bn_eval_scsh(512, bn1, params[P_C5B], p.scsh5, ss.s()));
    TRY(bn_eval_scsh(512, bn2, params[P_C6B], p.scsh6, ss.s()));
  }
  if (two_marks) TRY(ss.mark2());
  else TRY(ss.mark());
  // shadows(in, w, out): the fp16 copies a contraction reads / writes in fp16 mode (nulls otherwise: tf32 path)
  auto shadows = [&](TcEpilogue& e, const void* in16, const void* w16, void* out16) {
    e.in16 = h ? in16 : nullptr; e.w16 = h ? w16 : nullptr; e.out16 = h ? out16 : nullptr;
  };
  if (opt.jitter) TRY(c1_conv_fwd_jitter(x, B, 32, W, *opt.jitter, params[P_C1W], params[P_C1B], 1, A1f, st));
  else TRY(c1_conv_fwd(X, params[P_C1W], params[P_C1B], 1, A1f, st));
  TRY(maxpool_fwd(A1f, 2, 2, A1, st, h ? p.a1h : nullptr));
  TRY(ss.wait_mark());
  TcEpilogue ep;
  ep.relu = 1;
  ep.bias = params[P_C2B];
  shadows(ep, p.a1h, p.wp2h, nullptr);
  TRY(tc_conv_fprop(A1, p.wp2, 128, 3, 3, 1, 1, A2f, ep, st));
  TRY(maxpool_fwd(A2f, 2, 2, A2, st, h ? p.a2h : nullptr));
  ep.bias = params[P_C3B];
  shadows(ep, p.a2h, p.wp3h, p.a3h);
  ep.round_out = 1;   // A3 is the A operand of conv4's weight gradient
  TRY(tc_conv_fprop(A2, p.wp3, 256, 3, 3, 1, 1, A3, ep, st));
  ep.round_out = 0;
  ep.bias = params[P_C4B];
  shadows(ep, p.a3h, p.wp4h, nullptr);
  TRY(tc_conv_fprop(A3, p.wp4, 256, 3, 3, 1, 1, A4f, ep, st));
  TRY(maxpool_fwd(A4f, 2, 1, A4, st, h ? p.a4h : nullptr));
  TRY(ss.wait_mark2());   // conv5 ... Linear operands, LSTM bias sums, folded BatchNorm constants

  if (bn_train) {
    TRY(fill_zero(p.bnstats, 2 * 2 * 512 * sizeof(double), st));
    TcEpilogue raw;
    raw.bias = params[P_C5B];
    raw.bn_stats = p.bnstats;             // BatchNorm statistics out of the conv epilogue
    shadows(raw, p.a4h, p.wp5h, nullptr);
    TRY(tc_conv_fprop(A4, p.wp5, 512, 3, 3, 1, 1, Z5, raw, st));
    TRY(bn_train_finalize_apply(Z5, p.bnstats, bn1, p.scsh5, 1, A5, st, h ? p.a5h : nullptr));
    raw.bias = params[P_C6B];
    raw.bn_stats = p.bnstats + 1024;
    shadows(raw, p.a5h, p.wp6h, nullptr);
    TRY(tc_conv_fprop(A5, p.wp6, 512, 3, 3, 1, 1, Z6, raw, st));
    TRY(bn_train_finalize_apply(Z6, p.bnstats + 1024, bn2, p.scsh6, 1, A6f, st));
  } else {
    // frozen statistics: y = relu(conv*scale + shift), shift folds the conv bias (scsh5 / scsh6 prepared above)
    TcEpilogue f;
    f.relu = 1;
    f.scale = p.scsh5; f.bias = p.scsh5 + 512;
    shadows(f, p.a4h, p.wp5h, p.a5h);
    f.round_out = 1;   // A5 is the A operand of conv6's weight gradient
    TRY(tc_conv_fprop(A4, p.wp5, 512, 3, 3, 1, 1, A5, f, st));
    f.round_out = 0;
    f.scale = p.scsh6; f.bias = p.scsh6 + 512;
    shadows(f, p.a5h, p.wp6h, nullptr);
    TRY(tc_conv_fprop(A5, p.wp6, 512, 3, 3, 1, 1, A6f, f, st));
  }
  TRY(maxpool_fwd(A6f, 2, 1, A6, st, h ? p.a6h : nullptr));

  // conv7 writes the sequence-major (T,B,512) tensor directly (map_to_sequence fused)
  Img X0seq;
  X0seq.p = p.x0; X0seq.n = B; X0seq.h = 1; X0seq.w = T; X0seq.c = 512; X0seq.sn = 512; X0seq.sh = 0; X0seq.sw = (long long)B * 512;
  TcEpilogue e7;
  e7.bias = params[P_C7B];
  e7.round_out = 1;   // x0 is the A operand of the layer-0 W_ih weight gradient
  shadows(e7, p.a6h, p.wp7h, p.x0h);
  TRY(tc_conv_fprop(A6, p.wp7, 512, 2, 2, 0, 0, X0seq, e7, st));

  // two bidirectional LSTM layers: input projection GEMM (both directions at once, N = 2048) + recurrence
  const int TB = T * B;
  float* xin = p.x0;
  const void* xin16 = p.x0h;
  for (int l = 0; l < 2; ++l) {
    const float* const* lp = params + P_LSTM0 + l * 8;
    float* wih = l ? p.wih1 : p.wih0;
    float* bias = l ? p.bias1 : p.bias0;
    float* g = l ? p.g1 : p.g0;
    float* c = l ? p.c1 : p.c0;
    float* y = l ? p.y1 : p.y0;
    void* y16 = l ? p.y1h : p.y0h;
    TcEpilogue eg;
    eg.bias = bias;
    shadows(eg, xin16, l ? p.wih1h : p.wih0h, nullptr);
    TRY(tc_conv_fprop(img_nhwc(xin, 1, 1, TB, 512), wih, 2048, 1, 1, 0, 0, img_nhwc(g, 1, 1, TB, 2048), eg, st));
    TRY(lstm_layer_fwd(g, lp[1], lp[5], c, y, T, B, st, h ? y16 : nullptr, 1));
    xin = y;
    xin16 = y16;
  }
  TcEpilogue el;
  el.bias = params[P_LINB];
  el.log_softmax = opt.log_softmax;
  el.argmax = opt.argmax;
  shadows(el, p.y1h, p.wlinh, nullptr);
  TRY(tc_conv_fprop(img_nhwc(p.y1, 1, 1, TB, 512), params[P_LINW], V, 1, 1, 0, 0, img_nhwc(logits, 1, 1, TB, V), el, st));
  return QEB_OK;
}
}  // namespace

// dlogits: (T,B,V) dense. grads: P_COUNT pointers (entries may be NULL, all NULL-able: that gradient is skipped);
// every non-NULL gradient is ACCUMULATED into (zero it for a plain gradient). dx: (B,1,32,W) or NULL.
// The workspace must be the one the forward filled, with the same B, W, V, params and bn_train.
namespace {
// gradient tensors of the conv stack that are operands of a dgrad / wgrad contraction (slots of nn.cuh GradScales)
enum { GS_DZ7 = 0, GS_D6F, GS_D5, GS_D4F, GS_D3, GS_D2F, kGradSlots };
constexpr int kNetKindCrnn = 2;
int crnn_backward_body(const float* x, int B, int W, int V, const float* const* params, int bn_train, void* ws,
                       const float* dlogits, float* const* grads, float* dx, void* stream);
}
QEB_API int qeb_crnn_backward(const float* x, int B, int W, int V, const float* const* params, int bn_train, void* ws,
                              const float* dlogits, float* const* grads, float* dx, void* stream) {
  QEB_REQUIRE(x && params && ws && dlogits && grads, "crnn_backward: null pointer");
  grad_scales_prepare(kNetKindCrnn, params[0], kGradSlots, (cudaStream_t)stream);
  CallKey key;
  key.add(2).add(x).add(B).add(W).add(V).add(bn_train).add(ws).add(dlogits).add(dx).add(grad_scales_state(kNetKindCrnn, params[0]));
  key.ptrs(reinterpret_cast<const void* const*>(params), P_COUNT).ptrs(reinterpret_cast<const void* const*>(grads), P_COUNT);
  return qeb_run_cached(key, (cudaStream_t)stream, [&](cudaStream_t st) {
    return crnn_backward_body(x, B, W, V, params, bn_train, ws, dlogits, grads, dx, (void*)st);
  });
}
namespace {
int crnn_backward_body(const float* x, int B, int W, int V, const float* const* params, int bn_train, void* ws,
                       const float* dlogits, float* const* grads, float* dx, void* stream) {
  QEB_REQUIRE(x && params && ws && dlogits && grads, "crnn_backward: null pointer");
  QEB_REQUIRE(B > 0 && W >= 8 && W % 4 == 0 && V > 0 && V <= 96, "crnn_backward: B=%d W=%d V=%d unsupported", B, W, V);
  cudaStream_t st = (cudaStream_t)stream;
  const CrnnPlan p = make_plan(B, W, V, ws);
  const int T = p.T, TB = T * B;
  Img X = img_nhwc(const_cast<float*>(x), B, 32, W, 1);
  Img A1f = img_nhwc(p.a1f, B, 32, W, 64), A1 = img_nhwc(p.a1, B, 16, p.W2, 64);
  Img A2f = img_nhwc(p.a2f, B, 16, p.W2, 128), A2 = img_nhwc(p.a2, B, 8, p.W4, 128);
  Img A3 = img_nhwc(p.a3, B, 8, p.W4, 256);
  Img A4f = img_nhwc(p.a4f, B, 8, p.W4, 256), A4 = img_nhwc(p.a4, B, 4, p.W4, 256);
  Img Z5 = img_nhwc(p.z5, B, 4, p.W4, 512), A5 = img_nhwc(p.a5, B, 4, p.W4, 512);
  Img Z6 = img_nhwc(p.z6, B, 4, p.W4, 512), A6f = img_nhwc(p.a6f, B, 4, p.W4, 512), A6 = img_nhwc(p.a6, B, 2, p.W4, 512);
  Img D1f = img_nhwc(p.d1f, B, 32, W, 64), D1 = img_nhwc(p.d1, B, 16, p.W2, 64);
  Img D2f = img_nhwc(p.d2f, B, 16, p.W2, 128), D2 = img_nhwc(p.d2, B, 8, p.W4, 128);
  Img D3 = img_nhwc(p.d3, B, 8, p.W4, 256);
  Img D4f = img_nhwc(p.d4f, B, 8, p.W4, 256), D4 = img_nhwc(p.d4, B, 4, p.W4, 256);
  Img D5 = img_nhwc(p.d5, B, 4, p.W4, 512), D6f = img_nhwc(p.d6f, B, 4, p.W4, 512), D6 = img_nhwc(p.d6, B, 2, p.W4, 512);
  const TcEpilogue plain;
  TcEpilogue chained;   // the output is the operand of the next contraction(s) without an elementwise kernel in between
  chained.round_out = 1;
  SideStream ss;  // weight / bias gradients run beside the input-gradient chain
  TRY(ss.init(st));
  // scaled fp16 operands for the conv stack's backward contractions: needs the fp16 activation shadows of the forward pass
  GradScales gsc;
  if (fp16_fwd()) TRY(grad_scales_begin(kNetKindCrnn, params[0], kGradSlots, st, &gsc));
  const bool b16 = gsc.valid;
  // operands(slot, x16, dy16): shadows of a weight gradient's operands; dgrad16(e, slot, dy16, w16): the same for an input gradient
  WgradShadows wsh;
  auto operands = [&](int slot, const void* x16, const void* dy16) -> const WgradShadows* {
    if (!b16) return nullptr;
    wsh.x16 = x16; wsh.dy16 = dy16; wsh.alpha = gsc.inv + slot;
    return &wsh;
  };
  auto dgrad16 = [&](TcEpilogue& e, int slot, const void* dy16, const void* w16) {
    if (b16) { e.in16 = dy16; e.w16 = w16; e.alpha = gsc.inv + slot; }
  };

  // ---- Linear
  TRY(fill_zero(p.dlp, (size_t)TB * 96 * sizeof(float), st));
  TRY(fill_zero(p.wlinT, (size_t)512 * 96 * sizeof(float), st));
  {  // every re-layout of this pass in one launch
    PackBatch pk;
    pk.add(dlogits, p.dlp, 1, TB, V, 0, V, 1, 0, 96);
    pk.add(params[P_LINW], p.wlinT, 1, 512, V, 0, 1, 512, 0, 96);  // wlinT[c][v] = W[v][c]
    // the LSTM layers' d(input) operands are first read after the layer-1 recurrence: side stream, ahead of the conv stack's
    PackBatch pkl;
    static const bool split = !(getenv("QEB_PACK_SPLIT") && atoi(getenv("QEB_PACK_SPLIT")) == 0);
    for (int l = 0; l < 2; ++l)  // d(input) B operand [512][2048]: wihT[c][d*1024 + r] = W_ih_d[r][c]
      for (int d = 0; d < 2; ++d)
        (split ? pkl : pk).add(params[P_LSTM0 + l * 8 + d * 4], (l ? p.wihT1 : p.wihT0) + d * 1024, 1, 512, 1024, 0, 1, 512, 0, 2048);
    TRY(pack_flush(pk, st));
    if (split) {
      TRY(ss.fork());
      TRY(pack_flush(pkl, ss.s()));
      TRY(ss.mark2());
    }
  }
  {  // the conv stack's input-gradient operands are not needed before the LSTM layers are done: side stream
    PackBatch pk;
    if (b16) {   // fp16 B operands; the fp32 packs are not read then
      pk.add_dgrad16(params[P_C7W], p.wpd7h, 512, 512, 4);
      pk.add_dgrad16(params[P_C6W], p.wpd6h, 512, 512, 9);
      pk.add_dgrad16(params[P_C5W], p.wpd5h, 512, 256, 9);
      pk.add_dgrad16(params[P_C4W], p.wpd4h, 256, 256, 9);
      pk.add_dgrad16(params[P_C3W], p.wpd3h, 256, 128, 9);
      pk.add_dgrad16(params[P_C2W], p.wpd2h, 128, 64, 9);
    } else {
      pk.add_dgrad(params[P_C7W], p.wpd7, 512, 512, 4);
      pk.add_dgrad(params[P_C6W], p.wpd6, 512, 512, 9);
      pk.add_dgrad(params[P_C5W], p.wpd5, 512, 256, 9);
      pk.add_dgrad(params[P_C4W], p.wpd4, 256, 256, 9);
      pk.add_dgrad(params[P_C3W], p.wpd3, 256, 128, 9);
      pk.add_dgrad(params[P_C2W], p.wpd2, 128, 64, 9);
    }
    TRY(ss.fork());
    TRY(fill_zero(p.dwp2, p.dwp_bytes, ss.s()));  // packed conv weight-gradient accumulators: only side-stream kernels add into them
    TRY(pack_flush(pk, ss.s()));
    TRY(ss.mark());
  }
  Img DLP = img_nhwc(p.dlp, 1, 1, TB, 96);
  Img DLPv = img_nhwc(p.dlp, 1, 1, TB, V, 96);
  Img Y1 = img_nhwc(p.y1, 1, 1, TB, 512), Y0 = img_nhwc(p.y0, 1, 1, TB, 512), X0 = img_nhwc(p.x0, 1, 1, TB, 512);
  TRY(ss.fork());
  if (grads[P_LINW]) TRY(tc_conv_wgrad(Y1, DLPv, 1, 1, 0, 0, grads[P_LINW], 512, 1, 0, 0, ss.s()));
  if (grads[P_LINB]) TRY(colsum_acc(DLPv, grads[P_LINB], ss.s()));
  TRY(tc_conv_fprop(DLP, p.wlinT, 512, 1, 1, 0, 0, img_nhwc(p.dy1, 1, 1, TB, 512), plain, st));

  // ---- LSTM layers, top down
  for (int l = 1; l >= 0; --l) {
    const float* const* lp = params + P_LSTM0 + l * 8;
    float* const* lg = grads + P_LSTM0 + l * 8;
    float* g = l ? p.g1 : p.g0;
    const float* c = l ? p.c1 : p.c0;
    float* y = l ? p.y1 : p.y0;
    const float* dy = l ? p.dy1 : p.dy0;
    const Img Xin = l ? Y0 : X0;
    TRY(lstm_layer_bwd(g, c, dy, lp[1], lp[5], T, B, st, 1));  // g now holds d(pre-activations), (T,B,2,1024)
    TRY(ss.wait_mark2());   // the d(input) operands (first pass of the loop only)
    TRY(ss.fork());
    for (int d = 0; d < 2; ++d) {
      Img DG = img_nhwc(g + d * 1024, 1, 1, TB, 1024, 2048);
      if (lg[d * 4 + 0]) TRY(tc_conv_wgrad(Xin, DG, 1, 1, 0, 0, lg[d * 4 + 0], 512, 1, 0, 0, ss.s()));
      if (lg[d * 4 + 1]) {
        // dW_hh = sum_t dG_t^T h_{t-1} (forward) / h_{t+1} (reverse): the h sequence shifted by one step, zero outside
        Img Hs;
        Hs.p = y + d * kHid; Hs.n = 1; Hs.h = T; Hs.w = B; Hs.c = kHid; Hs.sn = (long long)T * B * 512; Hs.sh = (long long)B * 512; Hs.sw = 512;
        Img DGs;
        DGs.p = g + d * 1024; DGs.n = 1; DGs.h = T; DGs.w = B; DGs.c = 1024; DGs.sn = (long long)T * B * 2048; DGs.sh = (long long)B * 2048; DGs.sw = 2048;
        TRY(tc_conv_wgrad(Hs, DGs, 1, 1, d ? -1 : 1, 0, lg[d * 4 + 1], kHid, 1, 0, 0, ss.s()));
      }
      // b_ih and b_hh receive the same gradient: one pass over dG
      if (lg[d * 4 + 2]) TRY(colsum_acc(DG, lg[d * 4 + 2], ss.s(), lg[d * 4 + 3]));
      else if (lg[d * 4 + 3]) TRY(colsum_acc(DG, lg[d * 4 + 3], ss.s()));
    }
    // d(input) = dG * [W_ih_fwd ; W_ih_rev]
    // layer 1 -> dy0 feeds the layer-0 recurrence (which rounds its own operands); layer 0 -> dx0 = dz7 feeds conv7's
    // weight- and input-gradient contractions directly
    TcEpilogue eo = l ? plain : chained;
    int dz7_done = 0;
    if (l == 0) { eo.gs = gsc.slot(GS_DZ7, p.dx0h); eo.gs_done = &dz7_done; }
    TRY(tc_conv_fprop(img_nhwc(g, 1, 1, TB, 2048), l ? p.wihT1 : p.wihT0, 512, 1, 1, 0, 0, img_nhwc(l ? p.dy0 : p.dx0, 1, 1, TB, 512),
                      eo, st));
    QEB_REQUIRE(l || !gsc.amax || dz7_done, "crnn_backward: the gradient shadow of conv7's output was not written");
  }

  // ---- conv7 (dz7 = dx0, sequence-major view of a (B,1,T,512) image)
  Img DZ7;
  DZ7.p = p.dx0; DZ7.n = B; DZ7.h = 1; DZ7.w = T; DZ7.c = 512; DZ7.sn = 512; DZ7.sh = 0; DZ7.sw = (long long)B * 512;
  TRY(ss.fork());
  if (grads[P_C7W]) TRY(tc_conv_wgrad(A6, DZ7, 2, 2, 0, 0, p.dwp7, 4 * 512, 1, 2 * 512, 512, ss.s(), operands(GS_DZ7, p.a6h, p.dx0h)));
  if (grads[P_C7B]) TRY(colsum_acc(img_nhwc(p.dx0, 1, 1, TB, 512), grads[P_C7B], ss.s()));
  TRY(ss.wait_mark());
  {
    TcEpilogue e;
    dgrad16(e, GS_DZ7, p.dx0h, p.wpd7h);
    TRY(tc_conv_fprop(DZ7, p.wpd7, 512, 2, 2, 1, 1, D6, e, st));
  }

  // ---- conv6 + BN2 + ReLU + pool(2,1), conv5 + BN1 + ReLU
  const bool bn_grads = grads[P_BN1W] || grads[P_BN2W];
  if (bn_train || bn_grads) TRY(fill_zero(p.bnred, 2 * 512 * 2 * sizeof(double), st));
  const GradShadow gs6 = gsc.slot(GS_D6F, p.d6fh), gs5 = gsc.slot(GS_D5, p.d5h), gs4 = gsc.slot(GS_D4F, p.d4fh),
                   gs3 = gsc.slot(GS_D3, p.d3h), gs2 = gsc.slot(GS_D2F, p.d2fh);
  if (bn_train) {
    TRY(maxpool_bwd(A6f, D6, 2, 1, 0, nullptr, nullptr, D6f, st));  // grad at relu(bn(z6)); the ReLU mask comes from z6
    TRY(bn_bwd_reduce(Z6, D6f, p.scsh6, 1, p.bnred, st));
    TRY(bn_bwd_apply_train(Z6, D6f, p.scsh6, 1, p.bnred, params[P_BN2W], D6f, grads[P_BN2W], grads[P_BN2B], st, &gs6));
  } else if (bn_grads) {
    TRY(maxpool_bwd(A6f, D6, 2, 1, 1, nullptr, nullptr, D6f, st));  // g = routed grad * (a6f > 0)
    TRY(bn_bwd_reduce(A6f, D6f, p.scsh6, 2, p.bnred, st));
    TRY(bn_bwd_apply_eval(A6f, D6f, p.scsh6, 2, p.bnred, D6f, grads[P_BN2W], grads[P_BN2B], st, &gs6));
  } else {
    TRY(maxpool_bwd(A6f, D6, 2, 1, 1, p.scsh6, nullptr, D6f, st, nullptr, nullptr, nullptr, &gs6));  // dz6 = routed grad * (a6f > 0) * scale, one pass
  }
  TRY(ss.fork());
  if (grads[P_C6W]) TRY(tc_conv_wgrad(A5, D6f, 3, 3, 1, 1, p.dwp6, 9 * 512, 1, 3 * 512, 512, ss.s(), operands(GS_D6F, p.a5h, p.d6fh)));
  if (grads[P_C6B]) TRY(colsum_acc(D6f, grads[P_C6B], ss.s()));
  if (bn_train) {
    TcEpilogue e;
    dgrad16(e, GS_D6F, p.d6fh, p.wpd6h);
    TRY(tc_conv_fprop(D6f, p.wpd6, 512, 3, 3, 1, 1, D5, e, st));
    TRY(bn_bwd_reduce(Z5, D5, p.scsh5, 1, p.bnred + 1024, st));
    TRY(bn_bwd_apply_train(Z5, D5, p.scsh5, 1, p.bnred + 1024, params[P_BN1W], D5, grads[P_BN1W], grads[P_BN1B], st, &gs5));
  } else if (bn_grads) {
    TcEpilogue e;
    dgrad16(e, GS_D6F, p.d6fh, p.wpd6h);
    TRY(tc_conv_fprop(D6f, p.wpd6, 512, 3, 3, 1, 1, D5, e, st));
    TRY(bn_bwd_reduce(A5, D5, p.scsh5, 2, p.bnred + 1024, st));
    TRY(bn_bwd_apply_eval(A5, D5, p.scsh5, 2, p.bnred + 1024, D5, grads[P_BN1W], grads[P_BN1B], st, &gs5));
  } else {
    TcEpilogue e;
    e.scale = p.scsh5; e.mask = &A5;  // dz5 = d(a5) * scale, zero where a5 == 0, fused into the dgrad epilogue
    e.round_out = 1;
    int done = 0;
    e.gs = gs5; e.gs_done = &done;
    dgrad16(e, GS_D6F, p.d6fh, p.wpd6h);
    TRY(tc_conv_fprop(D6f, p.wpd6, 512, 3, 3, 1, 1, D5, e, st));
    QEB_REQUIRE(!gsc.amax || done, "crnn_backward: the gradient shadow of conv5's output was not written");
  }
  TRY(ss.fork());
  if (grads[P_C5W]) TRY(tc_conv_wgrad(A4, D5, 3, 3, 1, 1, p.dwp5, 9 * 256, 1, 3 * 256, 256, ss.s(), operands(GS_D5, p.a4h, p.d5h)));
  if (grads[P_C5B]) TRY(colsum_acc(D5, grads[P_C5B], ss.s()));
  {
    TcEpilogue e;
    dgrad16(e, GS_D5, p.d5h, p.wpd5h);
    TRY(tc_conv_fprop(D5, p.wpd5, 256, 3, 3, 1, 1, D4, e, st));
  }

  // ---- conv4 + ReLU + pool(2,1), conv3 + ReLU
  TRY(maxpool_bwd(A4f, D4, 2, 1, 1, nullptr, nullptr, D4f, st, nullptr, nullptr, nullptr, &gs4));
  TRY(ss.fork());
  if (grads[P_C4W]) TRY(tc_conv_wgrad(A3, D4f, 3, 3, 1, 1, p.dwp4, 9 * 256, 1, 3 * 256, 256, ss.s(), operands(GS_D4F, p.a3h, p.d4fh)));
  if (grads[P_C4B]) TRY(colsum_acc(D4f, grads[P_C4B], ss.s()));
  {
    TcEpilogue e;
    e.mask = &A3;  // ReLU of conv3 fused into the dgrad epilogue
    e.round_out = 1;
    int done = 0;
    e.gs = gs3; e.gs_done = &done;
    dgrad16(e, GS_D4F, p.d4fh, p.wpd4h);
    TRY(tc_conv_fprop(D4f, p.wpd4, 256, 3, 3, 1, 1, D3, e, st));
    QEB_REQUIRE(!gsc.amax || done, "crnn_backward: the gradient shadow of conv3's output was not written");
  }
  TRY(ss.fork());
  if (grads[P_C3W]) TRY(tc_conv_wgrad(A2, D3, 3, 3, 1, 1, p.dwp3, 9 * 128, 1, 3 * 128, 128, ss.s(), operands(GS_D3, p.a2h, p.d3h)));
  if (grads[P_C3B]) TRY(colsum_acc(D3, grads[P_C3B], ss.s()));
  {
    TcEpilogue e;
    dgrad16(e, GS_D3, p.d3h, p.wpd3h);
    TRY(tc_conv_fprop(D3, p.wpd3, 128, 3, 3, 1, 1, D2, e, st));
  }

  // ---- conv2 + ReLU + pool, conv1 + ReLU + pool
  TRY(maxpool_bwd(A2f, D2, 2, 2, 1, nullptr, nullptr, D2f, st, nullptr, nullptr, nullptr, &gs2));
  TRY(ss.fork());
  if (grads[P_C2W]) TRY(tc_conv_wgrad(A1, D2f, 3, 3, 1, 1, p.dwp2, 9 * 64, 1, 3 * 64, 64, ss.s(), operands(GS_D2F, p.a1h, p.d2fh)));
  if (grads[P_C2B]) TRY(colsum_acc(D2f, grads[P_C2B], ss.s()));
  const bool need_d1 = grads[P_C1W] || grads[P_C1B] || dx;
  if (need_d1) {
    {
      TcEpilogue e;
      dgrad16(e, GS_D2F, p.d2fh, p.wpd2h);
      TRY(tc_conv_fprop(D2f, p.wpd2, 64, 3, 3, 1, 1, D1, e, st));
    }
    TRY(maxpool_bwd(A1f, D1, 2, 2, 1, nullptr, nullptr, D1f, st));
    TRY(ss.fork());
    if (grads[P_C1W]) TRY(c1_conv_wgrad(X, D1f, grads[P_C1W], grads[P_C1B], ss.s()));
    if (dx) TRY(c1_conv_dgrad(D1f, params[P_C1W], img_nhwc(dx, B, 32, W, 1), st));
  }
  TRY(ss.join());
  {  // packed conv weight gradients -> torch layout, added into the caller's gradient tensors, one launch
    PackBatch pk;
    pk.accumulate = 1;
    if (grads[P_C2W]) pk.add_unpack_grad(p.dwp2, grads[P_C2W], 128, 64, 9);
    if (grads[P_C3W]) pk.add_unpack_grad(p.dwp3, grads[P_C3W], 256, 128, 9);
    if (grads[P_C4W]) pk.add_unpack_grad(p.dwp4, grads[P_C4W], 256, 256, 9);
    if (grads[P_C5W]) pk.add_unpack_grad(p.dwp5, grads[P_C5W], 512, 256, 9);
    if (grads[P_C6W]) pk.add_unpack_grad(p.dwp6, grads[P_C6W], 512, 512, 9);
    if (grads[P_C7W]) pk.add_unpack_grad(p.dwp7, grads[P_C7W], 512, 512, 4);
    TRY(pack_flush(pk, st));
  }
  grad_scales_commit(kNetKindCrnn, params[0]);
  return QEB_OK;
}
}  // namespace
