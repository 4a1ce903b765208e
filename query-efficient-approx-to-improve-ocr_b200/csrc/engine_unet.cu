// UNet preprocessor, forward and backward, as one host-side launch sequence over the sm_100a kernels.
// Reference: models/model_unet.py:9-109 (forward :49-76, _block :78-109): features 32..512, 18 x (conv3x3 no bias +
// BatchNorm + ReLU), 4 max-pools, 4 ConvTranspose2d(2, stride 2) with skip concats, 1x1 conv + sigmoid.
//
// Layout: NHWC fp32 in a caller-owned workspace. torch.cat((upconv, encoder), 1) costs nothing: the encoder writes
// channels [C,2C) and the up-convolution channels [0,C) of one shared buffer per level, and the decoder's first
// conv reads the 2C-channel buffer in place. All 3x3 convs except the first (one input channel, direct kernel) and
// the up-convolutions run on tcgen05 (conv_tc.cu). BatchNorm uses batch statistics in train mode (per-channel
// sums by a reduction kernel, normalise+ReLU by an elementwise kernel) and is folded into the conv epilogue in
// eval mode.
#include "nn.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int kLevels = 5;   // enc1..enc4 + bottleneck
constexpr int kUnits = 18;   // conv+BN+ReLU units
// parameter order of the C ABI: 9 blocks (enc1..4, bottleneck, dec4..1) x {conv1.w, norm1.w, norm1.b, conv2.w, norm2.w,
// norm2.b}, then upconv4..1 {w, b}, then conv {w, b}
enum { P_BLOCKS = 0, P_UP = 54, P_CONVW = 62, P_CONVB = 63, P_COUNT = 64 };
// buffers: per block {norm1.mean, norm1.var, norm1.nbt, norm2.mean, norm2.var, norm2.nbt}
constexpr int B_COUNT = 54;

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

struct UnetPlan {
  int B, H, W;
  int C[kLevels], h[kLevels], w[kLevels];
  size_t M[kLevels];
  // encoder / bottleneck blocks (index = level) and decoder blocks (index = level 0..3)
  float *ez1[kLevels], *ea1[kLevels], *ez2[kLevels], *pool[kLevels - 1], *bott;
  float *cat[kLevels - 1];
  float *dz1[kLevels - 1], *da1[kLevels - 1], *dz2[kLevels - 1], *dout[kLevels - 1];
  float* scsh;      // kUnits x 4*512
  double* bnstats;  // kUnits x 1024
  float *wp[kUnits], *wup[4];    // fprop B operands per conv unit / up-convolution
  float *wpd[kUnits], *wupd[4];  // dgrad B operands
  float* dwp[kUnits];            // packed [Cout][tap][Cin] weight-gradient accumulators, contiguous
  size_t dwp_bytes;
  float *sA[kLevels], *sB[kLevels], *sC[kLevels];
  double* bnred;    // kUnits x 1024
  // fp16 operand shadows of the forward pass (kind::f16 tensor-core path), same element layouts as their fp32 twins
  __half *ea1h[kLevels], *pool_h[kLevels - 1], *cat_h[kLevels - 1], *bott_h;
  __half *da1h[kLevels - 1], *dout_h[kLevels - 1];
  __half *wph[kUnits], *wuph[4];
  // backward pass with scaled fp16 operands (nn.cuh GradShadow): shadows of the gradient buffers and fp16 dgrad B operands
  __half *sA_h[kLevels], *sB_h[kLevels], *sC_h[kLevels];
  __half *wpdh[kUnits], *wupdh[4];
  size_t bytes;
};

UnetPlan make_plan(int B, int H, int W, void* base) {
  UnetPlan p;
  p.B = B; p.H = H; p.W = W;
  Arena a(base);
  for (int i = 0; i < kLevels; ++i) {
    p.C[i] = 32 << i; p.h[i] = H >> i; p.w[i] = W >> i;
    p.M[i] = (size_t)B * p.h[i] * p.w[i];
  }
  for (int i = 0; i < kLevels; ++i) {
    p.ez1[i] = a.take(p.M[i] * p.C[i]); p.ea1[i] = a.take(p.M[i] * p.C[i]); p.ez2[i] = a.take(p.M[i] * p.C[i]);
  }
  p.bott = a.take(p.M[4] * p.C[4]);
  for (int i = 0; i < kLevels - 1; ++i) {
    p.pool[i] = a.take(p.M[i + 1] * p.C[i]);
    p.cat[i] = a.take(p.M[i] * 2 * p.C[i]);
    p.dz1[i] = a.take(p.M[i] * p.C[i]); p.da1[i] = a.take(p.M[i] * p.C[i]); p.dz2[i] = a.take(p.M[i] * p.C[i]);
    p.dout[i] = a.take(p.M[i] * p.C[i]);
  }
  p.scsh = a.take((size_t)kUnits * 4 * 512);
  p.bnstats = reinterpret_cast<double*>(a.take((size_t)kUnits * 1024 * 2));
  for (int blk = 0; blk < 9; ++blk) {
    // block -> (cin of conv1, cout): encoders 1,32,64,128,256 -> C; decoders (block 5 + j, level 3 - j): 2C -> C
    const int lvl = blk < 5 ? blk : 3 - (blk - 5);
    const int cout = p.C[lvl];
    const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
    p.wp[blk * 2] = a.take((size_t)cout * 9 * cin1); p.wp[blk * 2 + 1] = a.take((size_t)cout * 9 * cout);
    p.wpd[blk * 2] = a.take((size_t)cout * 9 * cin1); p.wpd[blk * 2 + 1] = a.take((size_t)cout * 9 * cout);
  }
  for (int up = 0; up < 4; ++up) {  // up-conv `up` (upconv4..1) maps 2C -> C at level 3 - up
    const int C = p.C[3 - up];
    p.wup[up] = a.take((size_t)4 * C * 2 * C); p.wupd[up] = a.take((size_t)4 * C * 2 * C);
  }
  {
    const size_t o0 = a.off;
    for (int blk = 0; blk < 9; ++blk) {
      const int lvl = blk < 5 ? blk : 3 - (blk - 5);
      const int cout = p.C[lvl];
      const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
      p.dwp[blk * 2] = a.take((size_t)cout * 9 * cin1); p.dwp[blk * 2 + 1] = a.take((size_t)cout * 9 * cout);
    }
    p.dwp_bytes = a.off - o0;
  }
  for (int i = 0; i < kLevels; ++i) {
    p.sA[i] = a.take(p.M[i] * 2 * p.C[i]); p.sB[i] = a.take(p.M[i] * p.C[i]); p.sC[i] = a.take(p.M[i] * p.C[i]);
  }
  p.bnred = reinterpret_cast<double*>(a.take((size_t)kUnits * 1024 * 2));
  {
    auto half = [&](size_t n) { return reinterpret_cast<__half*>(a.take((n + 1) / 2)); };
    for (int i = 0; i < kLevels; ++i) p.ea1h[i] = half(p.M[i] * p.C[i]);
    p.bott_h = half(p.M[4] * p.C[4]);
    for (int i = 0; i < kLevels - 1; ++i) {
      p.pool_h[i] = half(p.M[i + 1] * p.C[i]);
      p.cat_h[i] = half(p.M[i] * 2 * p.C[i]);
      p.da1h[i] = half(p.M[i] * p.C[i]);
      p.dout_h[i] = half(p.M[i] * p.C[i]);
    }
    for (int blk = 0; blk < 9; ++blk) {
      const int lvl = blk < 5 ? blk : 3 - (blk - 5);
      const int cout = p.C[lvl];
      const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
      p.wph[blk * 2] = half((size_t)cout * 9 * cin1); p.wph[blk * 2 + 1] = half((size_t)cout * 9 * cout);
    }
    for (int up = 0; up < 4; ++up) p.wuph[up] = half((size_t)4 * p.C[3 - up] * 2 * p.C[3 - up]);
    for (int i = 0; i < kLevels; ++i) {
      p.sA_h[i] = half(p.M[i] * 2 * p.C[i]); p.sB_h[i] = half(p.M[i] * p.C[i]); p.sC_h[i] = half(p.M[i] * p.C[i]);
    }
    for (int blk = 0; blk < 9; ++blk) {
      const int lvl = blk < 5 ? blk : 3 - (blk - 5);
      const int cout = p.C[lvl];
      const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
      p.wpdh[blk * 2] = half((size_t)cout * 9 * cin1); p.wpdh[blk * 2 + 1] = half((size_t)cout * 9 * cout);
    }
    for (int up = 0; up < 4; ++up) p.wupdh[up] = half((size_t)4 * p.C[3 - up] * 2 * p.C[3 - up]);
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

struct Ctx {
  const float* const* params;
  void* const* buffers;
  float* const* grads;
  const UnetPlan* p;
  int bn_train;
  cudaStream_t st;
  SideStream* ss;  // backward only: weight / bias gradients run beside the input-gradient chain
  int* red_done;   // backward only, [kUnits]: 1 = the BatchNorm-backward reductions of that unit were produced by the kernel
                   // that wrote its output gradient (fused epilogue), the separate reduction pass is skipped
  bool scsh_ready = false;   // eval-mode forward: scale / shift of units 1.. were folded on the side stream beforehand
  bool act_only16 = false;   // train-mode forward of a network whose backward runs on fp16 operands: activations that only feed
                             // contractions (the first unit's output of every block, decoder outputs) skip their fp32 copy
  const GradScales* gs = nullptr;   // backward only: slots [0, 18) = dz of the conv units, [18, 22) = the up-convolutions' output gradients
};
constexpr int kGradSlots = kUnits + 4;
constexpr int kNetKindUnet = 1;

// epilogue of an input-gradient kernel whose output is the gradient at the OUTPUT of conv unit `unit` (train-mode BatchNorm):
// mask with relu(bn(z)) > 0 and accumulate that unit's BatchNorm-backward reductions
bool pool_fuse() {
  static const bool fuse = !(getenv("QEB_BN_POOL_FUSE") && atoi(getenv("QEB_BN_POOL_FUSE")) == 0);
  return fuse;
}

// BatchNorm-backward reductions inside the dgrad epilogue: it needs the transposing (legacy) epilogue, and since the plain outputs
// leave through the TMA-store epilogue the separate reduction pass is the cheaper of the two (same-box A/B with fp16 backward
// operands: 3.107 ms fused from 128 channels up, 3.082 from 256, 3.054 not at all). Off by default; QEB_BN_RED_FUSE=1 fuses again.
bool red_fuse() {
  static const bool fuse = getenv("QEB_BN_RED_FUSE") && atoi(getenv("QEB_BN_RED_FUSE")) != 0;
  return fuse;
}

TcEpilogue grad_into_unit(const Ctx& c, int unit, const Img* z) {
  TcEpilogue e;
  // fused only from `min_c` channels up: on the 32 / 64-channel layers the epilogue is the kernel's bottleneck already and the
  // separate reduction pass (3-4 TB/s) is cheaper than the longer epilogue (same-box A/B: 3.71 ms fused everywhere, 3.68 ms
  // from 128 channels or not at all)
  static const int min_c = getenv("QEB_BN_RED_FUSE_MINC") ? atoi(getenv("QEB_BN_RED_FUSE_MINC")) : 128;
  if (red_fuse() && c.bn_train && c.red_done && z && z->c >= min_c) {
    e.bn_z = z;
    e.bn_scsh = c.p->scsh + (size_t)unit * 4 * 512;
    e.bn_red = c.p->bnred + (size_t)unit * 1024;
    e.bn_red_fused = &c.red_done[unit];
  }
  return e;
}

BnParams bn_of(const Ctx& c, int block, int which) {
  BnParams b;
  b.gamma = c.params[block * 6 + which * 3 + 1];
  b.beta = c.params[block * 6 + which * 3 + 2];
  b.running_mean = static_cast<float*>(c.buffers[block * 6 + which * 3 + 0]);
  b.running_var = static_cast<float*>(c.buffers[block * 6 + which * 3 + 1]);
  b.num_batches_tracked = static_cast<long long*>(c.buffers[block * 6 + which * 3 + 2]);
  b.eps = 1e-5f; b.momentum = 0.1f;
  return b;
}

// one conv3x3 (no bias) + BN + ReLU unit. z: raw conv output (train mode only), out: activation. in16 / out16: fp16 shadows
// of `in` / `out` (NULL: tf32 operands from the fp32 tensors / no shadow wanted).
int unit_fwd(const Ctx& c, int block, int which, const Img& in, const Img& z, const Img& out, const __half* in16 = nullptr,
             __half* out16 = nullptr, const Img* pool = nullptr, __half* pool16 = nullptr, bool out_only16 = false) {
  const int unit = block * 2 + which;
  const float* w = c.params[block * 6 + which * 3];
  float* scsh = c.p->scsh + (size_t)unit * 4 * 512;
  const BnParams bn = bn_of(c, block, which);
  const int cout = out.c;
  if (in.c == 1) {
    TRY(c1_conv_fwd(in, w, nullptr, 0, z, c.st));
    if (c.bn_train) {
      TRY(bn_train_stats(z, c.p->bnstats + (size_t)unit * 1024, c.st));
      return bn_train_finalize_apply(z, c.p->bnstats + (size_t)unit * 1024, bn, scsh, 1, out, c.st, out16);
    }
    TRY(bn_eval_scsh(cout, bn, nullptr, scsh, c.st));
    return bn_apply(z, scsh, 1, out, c.st, out16);
  }
  const float* wp = c.p->wp[unit];
  if (c.bn_train) {
    TcEpilogue raw;
    raw.bn_stats = c.p->bnstats + (size_t)unit * 1024;   // per-channel sums come out of the conv epilogue
    if (in16) { raw.in16 = in16; raw.w16 = c.p->wph[unit]; }
    TRY(tc_conv_fprop(in, wp, cout, 3, 3, 1, 1, z, raw, c.st));
    if (pool) return bn_train_finalize_apply_pool(z, c.p->bnstats + (size_t)unit * 1024, bn, scsh, out, out16, *pool, pool16, c.st);
    return bn_train_finalize_apply(z, c.p->bnstats + (size_t)unit * 1024, bn, scsh, 1, out, c.st, out16,
                                   (out_only16 && c.act_only16 && in16) ? 1 : 0);
  }
  if (!c.scsh_ready) TRY(bn_eval_scsh(cout, bn, nullptr, scsh, c.st));
  TcEpilogue f;
  f.relu = 1; f.scale = scsh; f.bias = scsh + cout;
  f.round_out = 1;   // the activation is the A operand of the next unit's weight gradient
  if (in16) { f.in16 = in16; f.w16 = c.p->wph[unit]; }
  f.out16 = out16;
  return tc_conv_fprop(in, wp, cout, 3, 3, 1, 1, out, f, c.st);
}

// backward of one unit. g: gradient at the unit's output (overwritten with the gradient at the conv output);
// din: where the gradient at the unit's input goes (nullptr: not needed).
// z_lower: z of the conv unit that produced `in` (NULL: `in` is not the output of a conv unit, e.g. a pooled or concatenated
// tensor) - its BatchNorm-backward reductions are then fused into this unit's input-gradient kernel.
// in16: fp16 shadow of `in` (forward pass), g16: where the scaled fp16 shadow of dz goes, din16 / *din16_ok: the same for the input
// gradient of a decoder block's first unit (slot 18 + level) - all NULL-able: that contraction then reads tf32.
int unit_bwd(const Ctx& c, int block, int which, const Img& in, const Img& z, const Img& out, const Img& g, const Img* din,
             const Img* z_lower = nullptr, const __half* in16 = nullptr, __half* g16 = nullptr, __half* din16 = nullptr,
             int din_slot = -1, int* din16_ok = nullptr) {
  const int unit = block * 2 + which;
  const float* w = c.params[block * 6 + which * 3];
  float* const* gr = c.grads + block * 6 + which * 3;
  const float* scsh = c.p->scsh + (size_t)unit * 4 * 512;
  static const GradScales no_slots;
  const GradScales& gsc = c.gs ? *c.gs : no_slots;
  GradShadow gsh = gsc.slot(unit, in.c == 1 ? nullptr : g16);   // the one-channel layer's kernels read the fp32 gradient
  // dz is read by this unit's weight- and input-gradient contractions only (the convolutions have no bias): with both on the
  // fp16 shadow the fp32 copy is never read, so it is not written (QEB_DZ_ONLY16=0 keeps it)
  static const bool only16 = !(getenv("QEB_DZ_ONLY16") && atoi(getenv("QEB_DZ_ONLY16")) == 0);
  if (gsh.out16 && in16 && only16) gsh.only16 = 1;
  if (c.bn_train) {
    double* red = c.p->bnred + (size_t)unit * 1024;
    if (!(c.red_done && c.red_done[unit])) TRY(bn_bwd_reduce(z, g, scsh, 1, red, c.st));
    TRY(bn_bwd_apply_train(z, g, scsh, 1, red, nullptr, g, gr[1], gr[2], c.st, &gsh));
  } else {  // frozen statistics: the mask and xhat come from the layer output
    double* red = gr[1] ? c.p->bnred + (size_t)unit * 1024 : nullptr;
    if (red) TRY(bn_bwd_reduce(out, g, scsh, 2, red, c.st));
    TRY(bn_bwd_apply_eval(out, g, scsh, 2, red, g, gr[1], gr[2], c.st, &gsh));
  }
  TRY(c.ss->fork());
  if (in.c == 1) {
    if (gr[0]) TRY(c1_conv_wgrad(in, g, gr[0], nullptr, c.ss->s()));
    if (din) TRY(c1_conv_dgrad(g, w, *din, c.st));
    return QEB_OK;
  }
  const bool h16 = gsh.out16 != nullptr;   // dz has a scaled fp16 shadow: kind::f16 contractions, accumulators x 1 / S
  if (gr[0]) {
    WgradShadows sh;
    sh.x16 = in16; sh.dy16 = g16; sh.alpha = h16 ? gsc.inv + unit : nullptr;
    TRY(tc_conv_wgrad(in, g, 3, 3, 1, 1, c.p->dwp[unit], (long long)in.c * 9, 1, (long long)in.c * 3, in.c, c.ss->s(),
                      (h16 && in16) ? &sh : nullptr));
  }
  if (din) {
    TcEpilogue e = (which == 1 && z_lower) ? grad_into_unit(c, unit - 1, z_lower) : TcEpilogue();
    // first unit of a decoder block: half of its input gradient (the up-convolution's output gradient) is the operand of the
    // ConvTranspose weight- and input-gradient contractions with no elementwise kernel in between
    if (which == 0 && block >= 5) {
      e.round_out = 1;
      if (din_slot >= 0) { e.gs = gsc.slot(din_slot, din16); e.gs_done = din16_ok; e.gs_cols = din->c / 2; }   // the up-convolved half
    }
    if (h16) { e.in16 = g16; e.w16 = c.p->wpdh[unit]; e.alpha = gsc.inv + unit; }
    TRY(tc_conv_fprop(g, c.p->wpd[unit], in.c, 3, 3, 1, 1, *din, e, c.st));
  }
  return QEB_OK;
}

}  // namespace

QEB_API size_t qeb_unet_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H < 16 || W < 16 || H % 16 || W % 16) return 0;
  return make_plan(B, H, W, nullptr).bytes;
}

QEB_API int qeb_unet_num_params(void) { return P_COUNT; }
QEB_API int qeb_unet_num_buffers(void) { return B_COUNT; }

// x: (B,1,H,W) fp32, H and W multiples of 16. y: (B,1,H,W) in (0,1). bn_train as in qeb_crnn_forward.
namespace {
int unet_forward_body(const float* x, int B, int H, int W, const float* const* params, void* const* buffers, int bn_train, void* ws,
                      float* y, void* stream);
}
QEB_API int qeb_unet_forward(const float* x, int B, int H, int W, const float* const* params, void* const* buffers,
                             int bn_train, void* ws, float* y, void* stream) {
  QEB_REQUIRE(x && params && buffers && ws && y, "unet_forward: null pointer");
  CallKey key;   // one captured graph per distinct argument set (nn.cuh qeb_run_cached)
  key.add(3).add(x).add(B).add(H).add(W).add(bn_train).add(ws).add(y).add(grad_scales_state(kNetKindUnet, params[0]));
  key.ptrs(reinterpret_cast<const void* const*>(params), P_COUNT).ptrs(reinterpret_cast<const void* const*>(buffers), 3 * kUnits);
  return qeb_run_cached(key, (cudaStream_t)stream, [&](cudaStream_t st) {
    return unet_forward_body(x, B, H, W, params, buffers, bn_train, ws, y, (void*)st);
  });
}
namespace {
int unet_forward_body(const float* x, int B, int H, int W, const float* const* params, void* const* buffers, int bn_train, void* ws,
                      float* y, void* stream) {
  QEB_REQUIRE(x && params && buffers && ws && y, "unet_forward: null pointer");
  QEB_REQUIRE(B > 0 && H >= 16 && W >= 16 && H % 16 == 0 && W % 16 == 0, "unet_forward: B=%d H=%d W=%d unsupported", B, H, W);
  QEB_REQUIRE(((uintptr_t)ws & 255) == 0 && ((uintptr_t)x & 15) == 0, "unet_forward: workspace/input alignment");
  const UnetPlan p = make_plan(B, H, W, ws);
  Ctx c;
  c.params = params; c.buffers = buffers; c.grads = nullptr; c.p = &p; c.bn_train = bn_train; c.st = (cudaStream_t)stream;
  c.ss = nullptr;
  c.red_done = nullptr;
  SideStream ss;
  TRY(ss.init(c.st));
  if (bn_train) TRY(fill_zero(p.bnstats, (size_t)kUnits * 1024 * sizeof(double), c.st));
  const bool h16 = fp16_fwd();
  // the network has been through a backward call: the next one reads fp16 operands only (QEB_ACT_ONLY16=0 keeps the fp32 copies)
  static const bool act16 = !(getenv("QEB_ACT_ONLY16") && atoi(getenv("QEB_ACT_ONLY16")) == 0);
  c.act_only16 = act16 && bn_train && h16 && grad_scales_state(kNetKindUnet, params[0]) == 1;
  {  // the weight re-layouts of this pass in two launches: encoder blocks 1-3 (3 % of the bytes), needed at once, and the rest
     // - first read by encoder block 4, ~150 us into the pass - behind them (QEB_PACK_SPLIT=0: one launch, one wait)
    static const bool split = !(getenv("QEB_PACK_SPLIT") && atoi(getenv("QEB_PACK_SPLIT")) == 0);
    PackBatch pk, pk2;
    for (int blk = 0; blk < 9; ++blk) {
      const int lvl = blk < 5 ? blk : 3 - (blk - 5);
      const int cout = p.C[lvl];
      const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
      PackBatch& b = (split && blk >= 3) ? pk2 : pk;
      if (h16) {   // fp16 B operands; the fp32 packs are not needed by the forward pass then
        if (cin1 > 1) b.add_fprop16(params[blk * 6], p.wph[blk * 2], cout, cin1, 9);
        b.add_fprop16(params[blk * 6 + 3], p.wph[blk * 2 + 1], cout, cout, 9);
      } else {
        if (cin1 > 1) b.add_fprop(params[blk * 6], p.wp[blk * 2], cout, cin1, 9);
        b.add_fprop(params[blk * 6 + 3], p.wp[blk * 2 + 1], cout, cout, 9);
      }
    }
    for (int up = 0; up < 4; ++up) {  // ConvTranspose weight (2C, C, 2, 2) -> B operand [(dh*2+dw)*C + co][2C]
      const int C = p.C[3 - up];
      PackBatch& b = split ? pk2 : pk;
      b.add(params[P_UP + up * 2], h16 ? reinterpret_cast<float*>(p.wuph[up]) : p.wup[up], 4, C, 2 * C, 1, 4, (long long)C * 4,
            (long long)C * 2 * C, 2 * C);
      if (h16) b.last_to_half();
    }
    // beside the first (one-channel, direct) convolution unit, which reads no packed weights
    TRY(ss.fork());
    TRY(pack_flush(pk, ss.s()));
    if (split) {
      TRY(ss.mark());
      TRY(pack_flush(pk2, ss.s()));
    }
    if (!bn_train && ss.enabled) {   // frozen statistics: fold every unit's scale / shift here too (unit 0 is needed at once: main)
      for (int unit = 1; unit < kUnits; ++unit) {
        const int blk = unit / 2, lvl = blk < 5 ? blk : 3 - (blk - 5);
        TRY(bn_eval_scsh(p.C[lvl], bn_of(c, blk, unit & 1), nullptr, p.scsh + (size_t)unit * 4 * 512, ss.s()));
      }
      c.scsh_ready = true;
    }
    if (split) TRY(ss.mark2());
    else TRY(ss.mark());
  }

  Img in = img_nhwc(const_cast<float*>(x), B, H, W, 1);
  const __half* in16 = nullptr;   // fp16 shadow of `in` (none for the one-channel network input)
  for (int i = 0; i < kLevels; ++i) {  // encoder blocks + bottleneck
    const int C = p.C[i];
    Img z1 = img_nhwc(p.ez1[i], B, p.h[i], p.w[i], C), a1 = img_nhwc(p.ea1[i], B, p.h[i], p.w[i], C);
    Img z2 = img_nhwc(p.ez2[i], B, p.h[i], p.w[i], C);
    Img out = i < 4 ? img_nhwc(p.cat[i] + C, B, p.h[i], p.w[i], C, 2 * C) : img_nhwc(p.bott, B, p.h[i], p.w[i], C);
    __half* out16 = h16 ? (i < 4 ? p.cat_h[i] + C : p.bott_h) : nullptr;   // same channel slice of the fp16 concat buffer
    if (i == 3 || (i == 0 && !bn_train)) TRY(ss.wait_mark2());   // the deep layers' weights (eval mode: every unit's folded scale / shift)
    TRY(unit_fwd(c, i, 0, in, z1, a1, in16, h16 ? p.ea1h[i] : nullptr, nullptr, nullptr, true));
    TRY(ss.wait_mark());   // the packed weights (first pass of the loop only)
    const bool pool_fused = i < 4 && bn_train && pool_fuse();   // train mode: BatchNorm + ReLU + 2x2 pooling in one pass over z2
    Img pl_f = img_nhwc(p.pool[i < 4 ? i : 0], B, p.h[i < 4 ? i + 1 : 1], p.w[i < 4 ? i + 1 : 1], C);
    TRY(unit_fwd(c, i, 1, a1, z2, out, h16 ? p.ea1h[i] : nullptr, out16, pool_fused ? &pl_f : nullptr,
                 pool_fused && h16 ? p.pool_h[i] : nullptr));
    if (i < 4) {
      Img pl = img_nhwc(p.pool[i], B, p.h[i + 1], p.w[i + 1], C);
      if (!pool_fused) TRY(maxpool_fwd(out, 2, 2, pl, c.st, h16 ? p.pool_h[i] : nullptr));
      in = pl;
      in16 = h16 ? p.pool_h[i] : nullptr;
    }
  }
  Img below = img_nhwc(p.bott, B, p.h[4], p.w[4], p.C[4]);
  const __half* below16 = h16 ? p.bott_h : nullptr;
  for (int i = 3; i >= 0; --i) {  // decoder blocks: block index 5 + (3 - i), up-conv index (3 - i)
    const int C = p.C[i], blk = 5 + (3 - i), up = 3 - i;
    Img upo = img_nhwc(p.cat[i], B, p.h[i], p.w[i], C, 2 * C);
    TcEpilogue sh;
    if (h16) { sh.in16 = below16; sh.w16 = p.wuph[up]; sh.out16 = p.cat_h[i]; }
    sh.round_out = 1;   // the up-convolved half of the concat buffer is an A operand of dec conv1's weight gradient
    TRY(tc_convT_fprop(below, p.wup[up], params[P_UP + up * 2 + 1], upo, c.st, &sh));
    Img cat = img_nhwc(p.cat[i], B, p.h[i], p.w[i], 2 * C);
    Img z1 = img_nhwc(p.dz1[i], B, p.h[i], p.w[i], C), a1 = img_nhwc(p.da1[i], B, p.h[i], p.w[i], C);
    Img z2 = img_nhwc(p.dz2[i], B, p.h[i], p.w[i], C), out = img_nhwc(p.dout[i], B, p.h[i], p.w[i], C);
    TRY(unit_fwd(c, blk, 0, cat, z1, a1, h16 ? p.cat_h[i] : nullptr, h16 ? p.da1h[i] : nullptr, nullptr, nullptr, true));
    TRY(unit_fwd(c, blk, 1, a1, z2, out, h16 ? p.da1h[i] : nullptr, (h16 && i > 0) ? p.dout_h[i] : nullptr, nullptr, nullptr, i > 0));
    below = out;
    below16 = h16 ? p.dout_h[i] : nullptr;
  }
  return o1_conv_sigmoid_fwd(below, params[P_CONVW], params[P_CONVB], y, c.st);
}
}  // namespace

// dy: (B,1,H,W). grads: P_COUNT pointers, NULL = skip, non-NULL gradients are ACCUMULATED into. dx: (B,1,H,W) or NULL.
// y must be the forward's output; ws the forward's workspace.
namespace {
// packed [Cout][tap][Cin] weight-gradient accumulators of conv blocks [blk0, blk1) -> torch layout, ADDED into the caller's
// gradient tensors, one launch
int unpack_weight_grads(const UnetPlan& p, float* const* grads, int blk0, int blk1, cudaStream_t st) {
  PackBatch pk;
  pk.accumulate = 1;
  for (int blk = blk0; blk < blk1; ++blk) {
    const int lvl = blk < 5 ? blk : 3 - (blk - 5);
    const int cout = p.C[lvl];
    const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
    if (cin1 > 1 && grads[blk * 6]) pk.add_unpack_grad(p.dwp[blk * 2], grads[blk * 6], cout, cin1, 9);
    if (grads[blk * 6 + 3]) pk.add_unpack_grad(p.dwp[blk * 2 + 1], grads[blk * 6 + 3], cout, cout, 9);
  }
  return pack_flush(pk, st);
}
int unet_backward_impl(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws, const float* y,
                       const float* dy, float* const* grads, float* dx, void* const* bucket_events, void* stream);
}
QEB_API int qeb_unet_backward(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws,
                              const float* y, const float* dy, float* const* grads, float* dx, void* stream) {
  return unet_backward_impl(x, B, H, W, params, bn_train, ws, y, dy, grads, dx, nullptr, stream);
}

// The same backward for data-parallel training with an overlapped gradient exchange. Gradients become final in the order the
// backward pass (and, behind it, the weight-gradient stream) walks the network - last layer first - which is the flat gradient
// buffer from its END: `bucket_events` = two cudaEvent_t, recorded on the library's side stream (which the call joins before it
// returns) as soon as a range of the ABI parameter order is FINAL:
//   bucket_events[0]: parameters [30, 64) - the four decoder blocks, the up-convolutions, the final conv (12.2 MB)
//   bucket_events[1]: parameters [18, 30) - encoder block 4 and the bottleneck (17.7 MB)
// and parameters [0, 18) - encoder blocks 1-3, 1.1 MB - when the call's work completes, as always. A communication stream that
// waits for an event can all-reduce its range while the rest of the backward runs.
QEB_API int qeb_unet_backward_bucketed(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws,
                                       const float* y, const float* dy, float* const* grads, float* dx, void* const* bucket_events,
                                       void* stream) {
  QEB_REQUIRE(bucket_events && bucket_events[0] && bucket_events[1], "unet_backward_bucketed: two events are required");
  return unet_backward_impl(x, B, H, W, params, bn_train, ws, y, dy, grads, dx, bucket_events, stream);
}

namespace {
int unet_backward_body(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws, const float* y,
                       const float* dy, float* const* grads, float* dx, void* const* bucket_events, void* stream);
int unet_backward_impl(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws, const float* y,
                       const float* dy, float* const* grads, float* dx, void* const* bucket_events, void* stream) {
  QEB_REQUIRE(x && params && ws && y && dy && grads, "unet_backward: null pointer");
  // an event recorded inside a private capture could not be waited for by the caller's communication stream: plain launches
  if (bucket_events) {
    grad_scales_prepare(kNetKindUnet, params[0], kGradSlots, (cudaStream_t)stream);
    return unet_backward_body(x, B, H, W, params, bn_train, ws, y, dy, grads, dx, bucket_events, stream);
  }
  grad_scales_prepare(kNetKindUnet, params[0], kGradSlots, (cudaStream_t)stream);
  CallKey key;
  key.add(4).add(x).add(B).add(H).add(W).add(bn_train).add(ws).add(y).add(dy).add(dx).add(grad_scales_state(kNetKindUnet, params[0]));
  key.ptrs(reinterpret_cast<const void* const*>(params), P_COUNT).ptrs(reinterpret_cast<const void* const*>(grads), P_COUNT);
  return qeb_run_cached(key, (cudaStream_t)stream, [&](cudaStream_t st) {
    return unet_backward_body(x, B, H, W, params, bn_train, ws, y, dy, grads, dx, nullptr, (void*)st);
  });
}
int unet_backward_body(const float* x, int B, int H, int W, const float* const* params, int bn_train, void* ws, const float* y,
                       const float* dy, float* const* grads, float* dx, void* const* bucket_events, void* stream) {
  QEB_REQUIRE(x && params && ws && y && dy && grads, "unet_backward: null pointer");
  QEB_REQUIRE(B > 0 && H >= 16 && W >= 16 && H % 16 == 0 && W % 16 == 0, "unet_backward: B=%d H=%d W=%d unsupported", B, H, W);
  const UnetPlan p = make_plan(B, H, W, ws);
  Ctx c;
  c.params = params; c.buffers = nullptr; c.grads = grads; c.p = &p; c.bn_train = bn_train; c.st = (cudaStream_t)stream;
  SideStream ss;
  TRY(ss.init(c.st));
  c.ss = &ss;
  int red_done[kUnits] = {0};
  c.red_done = red_done;
  // scaled fp16 operands for the backward contractions (nn.cuh GradShadow): needs the fp16 activation shadows of the forward pass
  GradScales gsc;
  if (fp16_fwd()) TRY(grad_scales_begin(kNetKindUnet, params[0], kGradSlots, c.st, &gsc));
  c.gs = &gsc;
  const bool b16 = gsc.valid;
  TRY(fill_zero(p.bnred, (size_t)kUnits * 1024 * sizeof(double), c.st));
  {
    // two launches: the decoder blocks 1 and 2 with their up-convolutions and encoder blocks 1-3 (small), needed first, then the
    // deep layers (decoder blocks 3 and 4, bottleneck, encoder block 4: most of the bytes), first read by decoder block 3
    static const bool split = !(getenv("QEB_PACK_SPLIT") && atoi(getenv("QEB_PACK_SPLIT")) == 0);
    PackBatch pk, pk2;
    for (int blk = 0; blk < 9; ++blk) {
      const int lvl = blk < 5 ? blk : 3 - (blk - 5);
      const int cout = p.C[lvl];
      const int cin1 = blk == 0 ? 1 : (blk < 5 ? p.C[lvl - 1] : 2 * cout);
      PackBatch& b = (split && lvl >= 2) ? pk2 : pk;
      if (b16) {   // fp16 B operands; the fp32 packs are not read then
        if (cin1 > 1) b.add_dgrad16(params[blk * 6], p.wpdh[blk * 2], cout, cin1, 9);
        b.add_dgrad16(params[blk * 6 + 3], p.wpdh[blk * 2 + 1], cout, cout, 9);
      } else {
        if (cin1 > 1) b.add_dgrad(params[blk * 6], p.wpd[blk * 2], cout, cin1, 9);
        b.add_dgrad(params[blk * 6 + 3], p.wpd[blk * 2 + 1], cout, cout, 9);
      }
    }
    for (int up = 0; up < 4; ++up) {  // dgrad B operand [2C][(dh*2+dw)*C + co] from the torch weight (2C, C, 2, 2)
      const int C = p.C[3 - up];
      PackBatch& b = (split && up < 2) ? pk2 : pk;
      b.add(params[P_UP + up * 2], p.wupd[up], 2 * C, 4, C, (long long)C * 4, 1, 4, (long long)4 * C, C);
      if (b16) {
        b.add(params[P_UP + up * 2], reinterpret_cast<float*>(p.wupdh[up]), 2 * C, 4, C, (long long)C * 4, 1, 4, (long long)4 * C, C);
        b.last_to_half();
      }
    }
    TRY(ss.fork());   // beside the final 1x1 conv's backward kernel
    TRY(fill_zero(p.dwp[0], p.dwp_bytes, ss.s()));   // packed weight-gradient accumulators: only side-stream kernels add into them
    TRY(pack_flush(pk, ss.s()));
    TRY(ss.mark());
    if (split) {
      TRY(pack_flush(pk2, ss.s()));
      TRY(ss.mark2());
    }
  }

  // final 1x1 conv + sigmoid
  Img d0 = img_nhwc(p.dout[0], B, H, W, 32), g0 = img_nhwc(p.sC[0], B, H, W, 32);
  QEB_REQUIRE(grads[P_CONVW] && grads[P_CONVB], "unet_backward: the final conv's gradients are required");
  TRY(o1_conv_sigmoid_bwd(d0, params[P_CONVW], y, dy, g0, grads[P_CONVW], grads[P_CONVB], c.st));
  TRY(ss.wait_mark());

  for (int i = 0; i <= 3; ++i) {  // decoder blocks, top (full resolution) first
    const int C = p.C[i], blk = 5 + (3 - i), up = 3 - i;
    Img cat = img_nhwc(p.cat[i], B, p.h[i], p.w[i], 2 * C);
    Img z1 = img_nhwc(p.dz1[i], B, p.h[i], p.w[i], C), a1 = img_nhwc(p.da1[i], B, p.h[i], p.w[i], C);
    Img z2 = img_nhwc(p.dz2[i], B, p.h[i], p.w[i], C), out = img_nhwc(p.dout[i], B, p.h[i], p.w[i], C);
    Img g = img_nhwc(p.sC[i], B, p.h[i], p.w[i], C), ga1 = img_nhwc(p.sB[i], B, p.h[i], p.w[i], C);
    Img gcat = img_nhwc(p.sA[i], B, p.h[i], p.w[i], 2 * C);
    if (i == 2) TRY(ss.wait_mark2());   // the deep layers' operands (levels 2 and 3, bottleneck, encoder block 4)
    int du16 = 0;   // 1: the up-convolution's output gradient got its scaled fp16 shadow (slot 18 + i)
    TRY(unit_bwd(c, blk, 1, a1, z2, out, g, &ga1, &z1, p.da1h[i], p.sC_h[i]));
    TRY(unit_bwd(c, blk, 0, cat, z1, a1, ga1, &gcat, nullptr, p.cat_h[i], p.sB_h[i], p.sA_h[i], kUnits + i, &du16));
    du16 = du16 && b16;
    // up-convolution: dU = gcat[:, :C]
    Img dU = img_slice(gcat, 0, C);
    Img below = i < 3 ? img_nhwc(p.dout[i + 1], B, p.h[i + 1], p.w[i + 1], 2 * C) : img_nhwc(p.bott, B, p.h[4], p.w[4], 2 * C);
    const __half* below16 = i < 3 ? p.dout_h[i + 1] : p.bott_h;
    TRY(ss.fork());
    if (grads[P_UP + up * 2 + 1]) TRY(colsum_acc(dU, grads[P_UP + up * 2 + 1], ss.s()));
    if (grads[P_UP + up * 2]) {
      WgradShadows sh;
      sh.x16 = below16; sh.dy16 = p.sA_h[i]; sh.alpha = gsc.inv + kUnits + i;
      TRY(tc_convT_wgrad(below, dU, grads[P_UP + up * 2], ss.s(), du16 ? &sh : nullptr));
    }
    Img gbelow = img_nhwc(p.sC[i + 1], B, p.h[i + 1], p.w[i + 1], 2 * C);
    // `below` is the output of the second conv unit of the block one level down: fuse its BatchNorm-backward reductions
    const int blk_below = i < 3 ? 5 + (3 - (i + 1)) : 4;
    const Img z_below = img_nhwc(i < 3 ? p.dz2[i + 1] : p.ez2[4], B, p.h[i + 1], p.w[i + 1], 2 * C);
    TcEpilogue e = grad_into_unit(c, blk_below * 2 + 1, &z_below);
    if (du16) { e.in16 = p.sA_h[i]; e.w16 = p.wupdh[up]; e.alpha = gsc.inv + kUnits + i; }
    TRY(tc_convT_dgrad(dU, p.wupd[up], gbelow, e, c.st));
  }
  int unpacked_from = 9;   // conv blocks [unpacked_from, 9) already have their weight gradients in the caller's tensors
  if (bucket_events) {
    // the decoder is done on the main chain: the side stream (which holds its weight gradients, in order) waits for this point
    // of the main stream - the BatchNorm gradients are produced there -, finishes the decoder's conv weight gradients and
    // signals the caller
    TRY(ss.fork());
    TRY(unpack_weight_grads(p, grads, 5, 9, ss.s()));
    unpacked_from = 5;
    QEB_CUDA(cudaEventRecord((cudaEvent_t)bucket_events[0], ss.s()));
  }
  TRY(ss.join());  // the encoder phase re-uses the decoder phase's gradient buffers
  for (int i = 4; i >= 0; --i) {  // bottleneck, then encoder blocks
    const int C = p.C[i];
    Img z1 = img_nhwc(p.ez1[i], B, p.h[i], p.w[i], C), a1 = img_nhwc(p.ea1[i], B, p.h[i], p.w[i], C);
    Img z2 = img_nhwc(p.ez2[i], B, p.h[i], p.w[i], C);
    Img out = i < 4 ? img_nhwc(p.cat[i] + C, B, p.h[i], p.w[i], C, 2 * C) : img_nhwc(p.bott, B, p.h[i], p.w[i], C);
    Img g = img_nhwc(p.sC[i], B, p.h[i], p.w[i], C), ga1 = img_nhwc(p.sB[i], B, p.h[i], p.w[i], C);
    if (i < 4) {
      // gradient at the encoder output = skip-connection part of the concat gradient + routed pooling gradient
      Img gpool = img_nhwc(p.sA[i + 1], B, p.h[i + 1], p.w[i + 1], C);
      Img skip = img_nhwc(p.sA[i] + C, B, p.h[i], p.w[i], C, 2 * C);
      static const bool pool_red = !(getenv("QEB_BN_RED_FUSE_POOL") && atoi(getenv("QEB_BN_RED_FUSE_POOL")) == 0);
      if (bn_train && pool_red) {   // + the BatchNorm-backward reductions of this block's second unit (its output is `out`)
        const int unit = i * 2 + 1;
        TRY(maxpool_bwd(out, gpool, 2, 2, 0, nullptr, &skip, g, c.st, &z2, p.scsh + (size_t)unit * 4 * 512, p.bnred + (size_t)unit * 1024));
        red_done[unit] = 1;
      } else {
        TRY(maxpool_bwd(out, gpool, 2, 2, 0, nullptr, &skip, g, c.st));
      }
    }
    TRY(unit_bwd(c, i, 1, a1, z2, out, g, &ga1, &z1, p.ea1h[i], p.sC_h[i]));
    if (i > 0) {
      Img in = img_nhwc(p.pool[i - 1], B, p.h[i], p.w[i], p.C[i - 1]);
      Img gin = img_nhwc(p.sA[i], B, p.h[i], p.w[i], p.C[i - 1]);
      TRY(unit_bwd(c, i, 0, in, z1, a1, ga1, &gin, nullptr, p.pool_h[i - 1], p.sB_h[i]));
      if (i == 3 && bucket_events) {   // the bottleneck and encoder block 4: the same on the side stream, the main chain goes on
        TRY(ss.fork());
        TRY(unpack_weight_grads(p, grads, 3, 5, ss.s()));
        unpacked_from = 3;
        QEB_CUDA(cudaEventRecord((cudaEvent_t)bucket_events[1], ss.s()));
      }
    } else {
      Img in = img_nhwc(const_cast<float*>(x), B, H, W, 1);
      if (dx) {
        Img gx = img_nhwc(dx, B, H, W, 1);
        TRY(unit_bwd(c, i, 0, in, z1, a1, ga1, &gx));
      } else {
        TRY(unit_bwd(c, i, 0, in, z1, a1, ga1, nullptr));
      }
    }
  }
  TRY(ss.join());
  TRY(unpack_weight_grads(p, grads, 0, unpacked_from, c.st));
  grad_scales_commit(kNetKindUnet, params[0]);
  return QEB_OK;
}
}  // namespace
