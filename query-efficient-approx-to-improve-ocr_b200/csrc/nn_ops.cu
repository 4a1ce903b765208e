// Direct (SIMT) kernels of the UNet / CRNN step: the HBM-bound layers that are not dense contractions.
//   - 3x3 convolution with one input channel (CRNN conv1 models/model_crnn.py:37,48; UNet enc1conv1
//     models/model_unet.py:78-92): K = 9, bandwidth-bound, fprop / wgrad / dgrad
//   - final 1x1 convolution to one channel + sigmoid (models/model_unet.py:45-47,76) and its backward
//   - max pooling 2x2 / (2,1) (models/model_crnn.py:48-54, models/model_unet.py:14-20) forward and backward,
//     the backward fused with the ReLU mask and the skip-connection gradient add
//   - BatchNorm2d in train mode (batch statistics, running-stat update) and eval mode, fused with ReLU, forward
//     and backward (models/model_crnn.py:42-44,52-53; models/model_unet.py:93-106)
//   - bias gradients (column sums), ReLU backward.
// Layout: NHWC fp32 (nn.cuh Img). All kernels are grid-stride with float4 accesses along the channel dimension.
#include <mutex>
#include <unordered_map>
#include "nn.cuh"
#include "philox.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

struct Geo {  // pixel decomposition of an Img
  int h, w;
  long long sn, sh, sw;
};
__host__ __device__ inline Geo geo(const Img& a) {
  Geo g;
  g.h = a.h; g.w = a.w; g.sn = a.sn; g.sh = a.sh; g.sw = a.sw;
  return g;
}
// fp16 shadow of four consecutive channels (the tensor-core operand copy of an activation): 8-byte store
// (values beyond fp16's range saturate at +-65504 instead of becoming inf: the fp32 tensor keeps the exact value)
__device__ __forceinline__ float sat16(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
__device__ __forceinline__ void st4h(__half* p, const float4 v) {
  const __half2 a = __floats2half2_rn(sat16(v.x), sat16(v.y)), b = __floats2half2_rn(sat16(v.z), sat16(v.w));
  uint2 pk;
  pk.x = *reinterpret_cast<const uint32_t*>(&a);
  pk.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = pk;
}

// scaled fp16 shadow of a gradient (nn.cuh GradShadow) and the running maximum of |v|
__device__ __forceinline__ void st4h_scaled(__half* p, const float4 v, float s) { st4h(p, make_float4(v.x * s, v.y * s, v.z * s, v.w * s)); }
__device__ __forceinline__ float amax4(float m, const float4 v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// whole block (every thread calls it once, at the end of the kernel): one atomic per block, and only if it can raise the slot -
// thousands of same-address atomics per launch serialise in L2 (measured: +3-4 us per launch with one atomic per warp)
__device__ __forceinline__ void amax_flush(unsigned* slot, float m) {
  __shared__ float wmax[32];
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float b = threadIdx.x < (blockDim.x >> 5) ? wmax[threadIdx.x] : 0.f;
    b = warp_max(b);
    if (threadIdx.x == 0 && b > 0.f && __float_as_uint(b) > *reinterpret_cast<volatile unsigned*>(slot)) atomicMax(slot, __float_as_uint(b));
  }
}

__device__ __forceinline__ long long pix_off(const Geo& g, long long pix) {
  const int w = (int)(pix % g.w);
  const long long t = pix / g.w;
  const int h = (int)(t % g.h);
  const long long n = t / g.h;
  return n * g.sn + h * g.sh + w * g.sw;
}

bool vec4_ok(const Img& a) { return ((uintptr_t)a.p & 15) == 0 && a.sn % 4 == 0 && a.sh % 4 == 0 && a.sw % 4 == 0 && a.c % 4 == 0; }

// ------------------------------------------------------------------------------------------------ conv, Cin = 1
// thread = (4 consecutive pixels of a row, 4 output channels). The 3 x 6 input window of the pixel quad is loaded once
// (18 scalar loads for 4 pixels instead of 36), index arithmetic is 32-bit and amortised over the quad. w % 4 == 0.
// A quad index q = ((n * h + hv) * (w / 4) + wq) decomposes the image; CQ = COUT / 4 threads share a quad.
struct QuadPos {
  int n, hv, w0;
};
__device__ __forceinline__ QuadPos quad_pos(int q, int h, int wq_n) {
  QuadPos p;
  const int row = q / wq_n;
  p.w0 = (q - row * wq_n) * 4;
  p.n = row / h;
  p.hv = row - p.n * h;
  return p;
}
// xv[r][c] = x(hv - 1 + r, w0 - 1 + c), zero outside the image
__device__ __forceinline__ void load_window(const float* __restrict__ xb, const Geo& gx, int hv, int w0, float (&xv)[3][6]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hh = hv + r - 1;
    const bool rok = hh >= 0 && hh < gx.h;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const int ww = w0 + c - 1;
      xv[r][c] = (rok && ww >= 0 && ww < gx.w) ? __ldg(xb + (long long)hh * gx.sh + (long long)ww * gx.sw) : 0.f;
    }
  }
}

template <int COUT>
__global__ void __launch_bounds__(kThreads) c1_fwd_kernel(const float* __restrict__ x, Geo gx, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int relu, float* __restrict__ out,
                                                          Geo go, int n_quads) {
  qeb_pdl_sync();
  constexpr int CQ = COUT / 4;
  const int cq = threadIdx.x % CQ;
  float wr[4][9], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = __ldg(w + (cq * 4 + j) * 9 + t);
    br[j] = bias ? __ldg(bias + cq * 4 + j) : 0.f;
  }
  const int wq_n = gx.w / 4;
  const int stride = gridDim.x * (kThreads / CQ);
  for (int q = blockIdx.x * (kThreads / CQ) + threadIdx.x / CQ; q < n_quads; q += stride) {
    const QuadPos qp = quad_pos(q, gx.h, wq_n);
    float xv[3][6];
    load_window(x + (long long)qp.n * gx.sn, gx, qp.hv, qp.w0, xv);
    float* ob = out + (long long)qp.n * go.sn + (long long)qp.hv * go.sh + (long long)qp.w0 * go.sw + cq * 4;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = br[j];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) a = fmaf(xv[ky][px + kx], wr[j][ky * 3 + kx], a);
        acc[j] = relu ? fmaxf(a, 0.f) : a;
      }
      st4(ob + (long long)px * go.sw, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
  }
}

// Gaussian jitter FUSED into the conv1 input load (north_star; transform_helper.py:33-45 -> models/model_crnn.py:47-48).
// One block = one image x one band of kJitRows rows. Phase 1: the band plus one halo row above and below is jittered ONCE
// per pixel into shared memory (same Philox counters as jitter_kernel, so the image is bit-identical to what the standalone
// kernel writes); the band's own rows are also stored to `noisy_out` - the reference hands the noisy image to the OCR engine
// (train_nn_patch.py:290-291) and the weight gradient of conv1 reads it - and the noise to `noise_out` when asked for
// (AddGaussianNoice(return_noise=True)). Phase 2: the 3x3 convolution + bias + ReLU of c1_fwd_kernel out of shared memory.
// The separate jitter launch and its 16 KB-in / 16 KB-out pass per patch disappear; the base image is read once.
constexpr int kJitRows = 8;
template <int COUT>
__global__ void __launch_bounds__(kThreads) c1_jitter_fwd_kernel(const float* __restrict__ x, int h, int w,
                                                                 const float* __restrict__ sigma, float mean, float coef,
                                                                 unsigned long long seed, const unsigned long long* __restrict__ seed_dev,
                                                                 float* __restrict__ noisy_out, float* __restrict__ noise_out,
                                                                 const float* __restrict__ wgt, const float* __restrict__ bias, int relu,
                                                                 float* __restrict__ out, Geo go) {
  qeb_pdl_sync();
  extern __shared__ float tile[];   // (kJitRows + 2) rows x (w + 8) floats; pixel (r, c) at tile[(r - h0 + 1) * ts + c + 4]
  if (seed_dev) seed += *seed_dev;
  const int bands = (h + kJitRows - 1) / kJitRows;
  const int n = blockIdx.x / bands, h0 = (blockIdx.x - n * bands) * kJitRows;
  const int ts = w + 8, wq = w >> 2;
  const float* xb = x + (long long)n * h * w;
  const float sg = __ldg(sigma + n);
  for (int i = threadIdx.x; i < (kJitRows + 2) * wq; i += kThreads) {
    const int rr = i / wq, g = i - rr * wq, row = h0 - 1 + rr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= 0 && row < h) {
      const unsigned int grp = (unsigned int)(row * wq + g);
      const float4 z = qebrng::noise4(grp, n, seed, mean, sg);
      const float4 p = __ldg(reinterpret_cast<const float4*>(xb + (long long)row * w) + g);
      v = make_float4(qebrng::jitter1(p.x, z.x, coef), qebrng::jitter1(p.y, z.y, coef), qebrng::jitter1(p.z, z.z, coef),
                      qebrng::jitter1(p.w, z.w, coef));
      if (rr >= 1 && rr <= kJitRows) {
        const long long o = ((long long)n * h + row) * w + 4 * g;
        if (noisy_out) st4(noisy_out + o, v);
        if (noise_out) st4(noise_out + o, z);
      }
    }
    st4(tile + rr * ts + 4 + 4 * g, v);
  }
  for (int rr = threadIdx.x; rr < kJitRows + 2; rr += kThreads) {   // zero padding left and right of every row
    tile[rr * ts + 3] = 0.f;
    tile[rr * ts + 4 + w] = 0.f;
  }
  __syncthreads();
  constexpr int CQ = COUT / 4;
  const int cq = threadIdx.x % CQ;
  float wr[4][9], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = __ldg(wgt + (cq * 4 + j) * 9 + t);
    br[j] = bias ? __ldg(bias + cq * 4 + j) : 0.f;
  }
  const int rows = min(kJitRows, h - h0);
  for (int q = threadIdx.x / CQ; q < rows * wq; q += kThreads / CQ) {
    const int rr = q / wq, w0 = (q - rr * wq) * 4;
    float xv[3][6];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) xv[r][c] = tile[(rr + r) * ts + 3 + w0 + c];
    float* ob = out + (long long)n * go.sn + (long long)(h0 + rr) * go.sh + (long long)w0 * go.sw + cq * 4;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      float acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = br[j];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) a = fmaf(xv[ky][px + kx], wr[j][ky * 3 + kx], a);
        acc[j] = relu ? fmaxf(a, 0.f) : a;
      }
      st4(ob + (long long)px * go.sw, make_float4(acc[0], acc[1], acc[2], acc[3]));
    }
  }
}

// dw[co][tap] += sum_pix dy[pix][co] * x[pix+tap]; dbias[co] += sum_pix dy[pix][co]
template <int COUT>
__global__ void __launch_bounds__(kThreads) c1_wgrad_kernel(const float* __restrict__ x, Geo gx, const float* __restrict__ dy,
                                                            Geo gd, float* __restrict__ dw, float* __restrict__ dbias,
                                                            int n_quads) {
  qeb_pdl_sync();
  constexpr int CQ = COUT / 4;
  constexpr int PPB = kThreads / CQ;  // pixel quads per block iteration
  const int cq = threadIdx.x % CQ, pl = threadIdx.x / CQ;
  float acc[4][10];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[j][t] = 0.f;
  const int wq_n = gx.w / 4;
  const int stride = gridDim.x * PPB;
  for (int q = blockIdx.x * PPB + pl; q < n_quads; q += stride) {
    const QuadPos qp = quad_pos(q, gx.h, wq_n);
    float xv[3][6];
    load_window(x + (long long)qp.n * gx.sn, gx, qp.hv, qp.w0, xv);
    const float* db = dy + (long long)qp.n * gd.sn + (long long)qp.hv * gd.sh + (long long)qp.w0 * gd.sw + cq * 4;
    float4 g4[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) g4[px] = ld4(db + (long long)px * gd.sw);
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const float gv[4] = {g4[px].x, g4[px].y, g4[px].z, g4[px].w};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float xe = xv[ky][px + kx];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j][ky * 3 + kx] = fmaf(gv[j], xe, acc[j][ky * 3 + kx]);
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j][9] += gv[j];
    }
  }
  // reduce over the PPB quad lanes of the block through shared memory, then one atomic per (channel, tap)
  __shared__ float red[PPB][CQ * 4 + 1];
  for (int t = 0; t < 10; ++t) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[pl][cq * 4 + j] = acc[j][t];
    __syncthreads();
    if (threadIdx.x < COUT) {
      float s = 0.f;
      for (int i = 0; i < PPB; ++i) s += red[i][threadIdx.x];
      if (t < 9) atomicAdd(dw + threadIdx.x * 9 + t, s);
      else if (dbias) atomicAdd(dbias + threadIdx.x, s);
    }
  }
}

// dx[h][w] = sum_{ky,kx,co} dy[h-ky+1][w-kx+1][co] * w[co][ky][kx]; CQ lanes per pixel quad (4 channels each), the
// 3 x 6 window of dy vectors is loaded once for the four pixels, partial sums are shuffle-reduced over the CQ lanes
template <int COUT>
__global__ void __launch_bounds__(kThreads) c1_dgrad_kernel(const float* __restrict__ dy, Geo gd, const float* __restrict__ w,
                                                            float* __restrict__ dx, Geo gx, int n_quads) {
  qeb_pdl_sync();
  constexpr int CQ = COUT / 4;
  const int cq = threadIdx.x % CQ;
  float wr[4][9];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[j][t] = __ldg(w + (cq * 4 + j) * 9 + t);
  const int wq_n = gx.w / 4;
  const int stride = gridDim.x * (kThreads / CQ);
  const int n_iter = (n_quads + stride - 1) / stride;
  int q = blockIdx.x * (kThreads / CQ) + threadIdx.x / CQ;
  for (int it = 0; it < n_iter; ++it, q += stride) {  // uniform trip count: shuffles below need full warps
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const bool live = q < n_quads;
    QuadPos qp = {0, 0, 0};
    if (live) {
      qp = quad_pos(q, gx.h, wq_n);
      const float* db = dy + (long long)qp.n * gd.sn + cq * 4;
#pragma unroll
      for (int r = 0; r < 3; ++r) {   // dy row hv - 1 + r contributes with ky = 2 - r
        const int hh = qp.hv + r - 1;
        if (hh < 0 || hh >= gd.h) continue;
#pragma unroll
        for (int c = 0; c < 6; ++c) {   // dy column w0 - 1 + c contributes to pixel px with kx = px + 1 - (c - 1) ... kx = px - c + 2
          const int ww = qp.w0 + c - 1;
          if (ww < 0 || ww >= gd.w) continue;
          const float4 g = ld4(db + (long long)hh * gd.sh + (long long)ww * gd.sw);
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const int kx = px - c + 2;
            if (kx >= 0 && kx < 3) {
              const int tap = (2 - r) * 3 + kx;
              a[px] = fmaf(g.x, wr[0][tap], a[px]);
              a[px] = fmaf(g.y, wr[1][tap], a[px]);
              a[px] = fmaf(g.z, wr[2][tap], a[px]);
              a[px] = fmaf(g.w, wr[3][tap], a[px]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = CQ / 2; o > 0; o >>= 1) {
#pragma unroll
      for (int px = 0; px < 4; ++px) a[px] += __shfl_xor_sync(FULL_MASK, a[px], o);
    }
    if (live && cq < 4) {   // lane cq writes pixel cq of the quad
      const float v = cq == 0 ? a[0] : (cq == 1 ? a[1] : (cq == 2 ? a[2] : a[3]));
      dx[(long long)qp.n * gx.sn + (long long)qp.hv * gx.sh + (long long)(qp.w0 + cq) * gx.sw] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ 1x1 conv -> 1 ch + sigmoid
template <int CIN>
__global__ void __launch_bounds__(kThreads) o1_fwd_kernel(const float* __restrict__ x, Geo gx, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ y, long long n_pix) {
  qeb_pdl_sync();
  constexpr int CQ = CIN / 4;
  const int cq = threadIdx.x % CQ;
  const float4 wv = ld4(w + cq * 4);
  const float bv = __ldg(b);
  const long long stride = (long long)gridDim.x * (kThreads / CQ);
  const long long n_iter = (n_pix + stride - 1) / stride;
  long long pix = (long long)blockIdx.x * (kThreads / CQ) + threadIdx.x / CQ;
  for (long long it = 0; it < n_iter; ++it, pix += stride) {
    float a = 0.f;
    const bool live = pix < n_pix;
    if (live) {
      const float4 v = ld4(x + pix_off(gx, pix) + cq * 4);
      a = v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
    }
#pragma unroll
    for (int o = CQ / 2; o > 0; o >>= 1) a += __shfl_xor_sync(FULL_MASK, a, o);
    if (live && cq == 0) y[pix] = 1.f / (1.f + expf(-(a + bv)));
  }
}

template <int CIN>
__global__ void __launch_bounds__(kThreads) o1_bwd_kernel(const float* __restrict__ x, Geo gx, const float* __restrict__ w,
                                                          const float* __restrict__ y, const float* __restrict__ dy,
                                                          float* __restrict__ dx, Geo gdx, float* __restrict__ dw,
                                                          float* __restrict__ db, long long n_pix) {
  qeb_pdl_sync();
  constexpr int CQ = CIN / 4;
  constexpr int PPB = kThreads / CQ;
  const int cq = threadIdx.x % CQ, pl = threadIdx.x / CQ;
  const float4 wv = ld4(w + cq * 4);
  float4 aw = make_float4(0.f, 0.f, 0.f, 0.f);
  float ab = 0.f;
  const long long stride = (long long)gridDim.x * PPB;
  for (long long pix = (long long)blockIdx.x * PPB + pl; pix < n_pix; pix += stride) {
    const float yv = y[pix];
    const float dz = dy[pix] * yv * (1.f - yv);
    const float4 xv = ld4(x + pix_off(gx, pix) + cq * 4);
    aw.x = fmaf(dz, xv.x, aw.x); aw.y = fmaf(dz, xv.y, aw.y); aw.z = fmaf(dz, xv.z, aw.z); aw.w = fmaf(dz, xv.w, aw.w);
    ab += dz;
    st4(dx + pix_off(gdx, pix) + cq * 4, make_float4(dz * wv.x, dz * wv.y, dz * wv.z, dz * wv.w));
  }
  __shared__ float red[PPB][CIN + 1];
  __shared__ float redb[PPB];
  red[pl][cq * 4] = aw.x; red[pl][cq * 4 + 1] = aw.y; red[pl][cq * 4 + 2] = aw.z; red[pl][cq * 4 + 3] = aw.w;
  if (cq == 0) redb[pl] = ab;
  __syncthreads();
  if (threadIdx.x < CIN) {
    float s = 0.f;
    for (int i = 0; i < PPB; ++i) s += red[i][threadIdx.x];
    atomicAdd(dw + threadIdx.x, s);
  } else if (threadIdx.x == CIN) {
    float s = 0.f;
    for (int i = 0; i < PPB; ++i) s += redb[i];
    atomicAdd(db, s);
  }
}

// relu: 0 = none, 1 = mask recomputed from the pre-activation v (z*scale+shift > 0), 2 = v IS the ReLU output (v > 0)
__device__ __forceinline__ float4 bn_masked_grad(const float4 v, const float4 g, const float4 sc, const float4 sh, int relu) {
  if (!relu) return g;
  if (relu == 2) return make_float4(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f, v.z > 0.f ? g.z : 0.f, v.w > 0.f ? g.w : 0.f);
  return make_float4(fmaf(v.x, sc.x, sh.x) > 0.f ? g.x : 0.f, fmaf(v.y, sc.y, sh.y) > 0.f ? g.y : 0.f,
                     fmaf(v.z, sc.z, sh.z) > 0.f ? g.z : 0.f, fmaf(v.w, sc.w, sh.w) > 0.f ? g.w : 0.f);
}

// ------------------------------------------------------------------------------------------------ max pooling
template <int PH, int PW>
__global__ void __launch_bounds__(kThreads) maxpool_fwd_kernel(const float* __restrict__ x, Geo gx, float* __restrict__ out,
                                                               Geo go, int cq_n, long long total, __half* __restrict__ out16) {
  qeb_pdl_sync();
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < (int)total; i += gridDim.x * kThreads) {   // 32-bit index math (host check)
    const int cq = i % cq_n;
    const int pix = i / cq_n;
    const int wv = pix % go.w;
    const int t = pix / go.w;
    const int hv = t % go.h;
    const long long n = t / go.h;
    const float* xb = x + n * gx.sn + (long long)(hv * PH) * gx.sh + (long long)(wv * PW) * gx.sw + cq * 4;
    float4 m = ld4(xb);
#pragma unroll
    for (int a = 0; a < PH; ++a)
#pragma unroll
      for (int b = 0; b < PW; ++b) {
        if (a == 0 && b == 0) continue;
        const float4 v = ld4(xb + a * gx.sh + b * gx.sw);
        // NaN-propagating max like ATen (v > m || isnan(v))
        m.x = (v.x > m.x || v.x != v.x) ? v.x : m.x;
        m.y = (v.y > m.y || v.y != v.y) ? v.y : m.y;
        m.z = (v.z > m.z || v.z != v.z) ? v.z : m.z;
        m.w = (v.w > m.w || v.w != v.w) ? v.w : m.w;
      }
    st4(out + n * go.sn + hv * go.sh + wv * go.sw + cq * 4, qeb_tf32r4(m));   // weight-gradient operand: rounded to tf32 here
    if (out16) st4h(out16 + n * go.sn + hv * go.sh + wv * go.sw + cq * 4, m);
  }
}

__device__ __forceinline__ float comp4(float4 q, int j) { return j == 0 ? q.x : (j == 1 ? q.y : (j == 2 ? q.z : q.w)); }

template <int PH, int PW, bool RED>
__global__ void __launch_bounds__(kThreads) maxpool_bwd_kernel(const float* __restrict__ x, Geo gx, const float* __restrict__ dy,
                                                               Geo gd, int relu_mask, const float* __restrict__ chan_scale,
                                                               const float* __restrict__ add, Geo ga, float* __restrict__ dx,
                                                               Geo gdx, int cq_n, long long total, const float* __restrict__ bn_z,
                                                               Geo gz, const float* __restrict__ bn_scsh, double* __restrict__ bn_red,
                                                               __half* __restrict__ dx16, const float* __restrict__ gscale,
                                                               unsigned* __restrict__ gamax) {
  qeb_pdl_sync();
  const float gs_s = dx16 ? __ldg(gscale) : 1.f;
  float gs_m = 0.f;
  // bn_red != NULL: dx is the gradient at the output of a train-mode conv + BN + ReLU unit whose pre-activation is bn_z; its
  // BatchNorm-backward reductions (sum of masked g, sum of masked g * xhat) are accumulated here instead of in a separate
  // pass over (z, dx). kThreads is a multiple of cq_n (host check): a thread keeps its channel quad.
  float rs[4] = {0.f, 0.f, 0.f, 0.f}, rq[4] = {0.f, 0.f, 0.f, 0.f};
  float4 bsc, bsh, bmu, bis;
  if constexpr (RED) {
    const int C = cq_n * 4, cq = threadIdx.x % cq_n;
    bsc = ld4(bn_scsh + cq * 4); bsh = ld4(bn_scsh + C + cq * 4); bmu = ld4(bn_scsh + 2 * C + cq * 4); bis = ld4(bn_scsh + 3 * C + cq * 4);
  }
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < (int)total; i += gridDim.x * kThreads) {   // 32-bit index math (host check)
    const int cq = i % cq_n;
    const int pix = i / cq_n;
    const int wv = pix % gd.w;
    const int t = pix / gd.w;
    const int hv = t % gd.h;
    const long long n = t / gd.h;
    const float* xb = x + n * gx.sn + (long long)(hv * PH) * gx.sh + (long long)(wv * PW) * gx.sw + cq * 4;
    float4 v[PH * PW];   // accessed through comp4() with compile-time indices only: stays in registers
#pragma unroll
    for (int a = 0; a < PH; ++a)
#pragma unroll
      for (int b = 0; b < PW; ++b) v[a * PW + b] = ld4(xb + a * gx.sh + b * gx.sw);
    float4 g4 = ld4(dy + n * gd.sn + hv * gd.sh + wv * gd.sw + cq * 4);
    if (chan_scale) {
      const float4 s4 = ld4(chan_scale + cq * 4);
      g4.x *= s4.x; g4.y *= s4.y; g4.z *= s4.z; g4.w *= s4.w;
    }
    int arg[4];
    bool pass[4];   // ReLU mask of the winning element, decided here so that the scatter loop below never indexes v[] by arg[]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int best = 0;
      float m = comp4(v[0], j);
#pragma unroll
      for (int k = 1; k < PH * PW; ++k) {
        const float c = comp4(v[k], j);
        if (c > m || c != c) { m = c; best = k; }
      }
      arg[j] = best;
      pass[j] = !relu_mask || m > 0.f;
    }
#pragma unroll
    for (int a = 0; a < PH; ++a)
#pragma unroll
      for (int b = 0; b < PW; ++b) {
        const int k = a * PW + b;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (arg[j] == k && pass[j]) ? comp4(g4, j) : 0.f;
        const long long off = n * gdx.sn + (long long)(hv * PH + a) * gdx.sh + (long long)(wv * PW + b) * gdx.sw + cq * 4;
        if (add) {
          const float4 q = ld4(add + n * ga.sn + (long long)(hv * PH + a) * ga.sh + (long long)(wv * PW + b) * ga.sw + cq * 4);
          o[0] += q.x; o[1] += q.y; o[2] += q.z; o[3] += q.w;
        }
        st4(dx + off, qeb_tf32r4(make_float4(o[0], o[1], o[2], o[3])));   // dgrad / wgrad operand: rounded to tf32 here
        if (dx16) st4h_scaled(dx16 + off, make_float4(o[0], o[1], o[2], o[3]), gs_s);
        if (gamax) gs_m = amax4(gs_m, make_float4(o[0], o[1], o[2], o[3]));
        if constexpr (RED) {
          const float4 zv = ld4(bn_z + n * gz.sn + (long long)(hv * PH + a) * gz.sh + (long long)(wv * PW + b) * gz.sw + cq * 4);
          const float4 gm = bn_masked_grad(zv, make_float4(o[0], o[1], o[2], o[3]), bsc, bsh, 1);
          rs[0] += gm.x; rs[1] += gm.y; rs[2] += gm.z; rs[3] += gm.w;
          rq[0] = fmaf(gm.x, (zv.x - bmu.x) * bis.x, rq[0]); rq[1] = fmaf(gm.y, (zv.y - bmu.y) * bis.y, rq[1]);
          rq[2] = fmaf(gm.z, (zv.z - bmu.z) * bis.z, rq[2]); rq[3] = fmaf(gm.w, (zv.w - bmu.w) * bis.w, rq[3]);
        }
      }
  }
  if (gamax) amax_flush(gamax, gs_m);
  if constexpr (RED) {   // block reduction over the row lanes of each channel, then one double atomic per channel and block
    extern __shared__ float sm[];
    const int C = cq_n * 4, rpb = kThreads / cq_n, cq = threadIdx.x % cq_n, rl = threadIdx.x / cq_n;
    float* ss = sm;
    float* sq = sm + rpb * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) { ss[rl * C + cq * 4 + j] = rs[j]; sq[rl * C + cq * 4 + j] = rq[j]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      double a = 0.0, b = 0.0;
      for (int r = 0; r < rpb; ++r) { a += ss[r * C + c]; b += sq[r * C + c]; }
      atomicAdd(bn_red + c, a);
      atomicAdd(bn_red + C + c, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------ batch norm
// z: [M][C] rows `zs` floats apart. thread = (row lane, channel quad).
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const float* __restrict__ z, long long zs, long long M, int C,
                                                            double* __restrict__ stats) {
  qeb_pdl_sync();
  const int cq_n = C / 4, rpb = kThreads / cq_n;
  const int cq = threadIdx.x % cq_n, rl = threadIdx.x / cq_n;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < rpb) {
    const long long step = (long long)gridDim.x * rpb;
    long long r = (long long)blockIdx.x * rpb + rl;
    for (; r + 3 * step < M; r += 4 * step) {  // four independent loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld4(z + (r + u * step) * zs + cq * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
        q[0] = fmaf(v[u].x, v[u].x, q[0]); q[1] = fmaf(v[u].y, v[u].y, q[1]);
        q[2] = fmaf(v[u].z, v[u].z, q[2]); q[3] = fmaf(v[u].w, v[u].w, q[3]);
      }
    }
    for (; r < M; r += step) {
      const float4 v = ld4(z + r * zs + cq * 4);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
  }
  extern __shared__ float sm[];  // [2][rpb][C]
  float* ss = sm;
  float* sq = sm + rpb * C;
  if (rl < rpb) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { ss[rl * C + cq * 4 + j] = s[j]; sq[rl * C + cq * 4 + j] = q[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kThreads) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < rpb; ++i) { a += ss[i * C + c]; b += sq[i * C + c]; }
    atomicAdd(stats + c, a);
    atomicAdd(stats + C + c, b);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long count, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   long long* nbt, float eps, float momentum, float* __restrict__ scsh) {
  qeb_pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const double mean = stats[c] / (double)count;
  double var = stats[C + c] / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  scsh[c] = sc;
  scsh[C + c] = beta[c] - (float)mean * sc;
  scsh[2 * C + c] = (float)mean;
  scsh[3 * C + c] = invstd;
  if (running_mean) {
    const double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_scsh_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                    const float* __restrict__ conv_bias, float* __restrict__ scsh) {
  qeb_pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(rv[c] + eps);
  const float sc = gamma[c] * invstd;
  scsh[c] = sc;
  // conv_bias given: shift folds the convolution bias (the conv epilogue then applies scale/shift directly)
  scsh[C + c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - rm[c]) * sc;
  // frozen statistics: the backward only keeps the layer OUTPUT, from which xhat = (out - beta)/gamma where out > 0
  scsh[2 * C + c] = beta[c];
  scsh[3 * C + c] = gamma[c] != 0.f ? 1.f / gamma[c] : 0.f;
}

__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const float* __restrict__ z, long long zs, long long M, int C,
                                                            const float* __restrict__ scsh, int relu, float* __restrict__ out,
                                                            long long os, __half* __restrict__ out16) {
  qeb_pdl_sync();
  const int cq_n = C / 4;
  const int cq = threadIdx.x % cq_n;   // kThreads is a multiple of C/4 (host check): constant channel quad, no division in the loop
  const float4 sc = ld4(scsh + cq * 4), sh = ld4(scsh + C + cq * 4);
  const long long rstep = (long long)gridDim.x * (kThreads / cq_n);
  for (long long r = (long long)blockIdx.x * (kThreads / cq_n) + threadIdx.x / cq_n; r < M; r += rstep) {
    const float4 v = ld4(z + r * zs + cq * 4);
    float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    st4(out + r * os + cq * 4, qeb_tf32r4(o));   // weight-gradient operand: rounded to tf32 here
    if (out16) st4h(out16 + r * os + cq * 4, o);
  }
}

// train mode, finalize + apply in one launch: every thread derives scale/shift of its four channels from the batch sums
// (kThreads is a multiple of C/4, so a thread keeps its channel quad over the grid-stride loop); block 0 also stores
// scale/shift/mean/invstd for the backward and updates the running statistics.
__device__ __forceinline__ void bn_train_scale_shift(long long M, int C, int cq, const double* __restrict__ stats,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* running_mean, float* running_var, long long* nbt, float eps,
                                                     float momentum, float* __restrict__ scsh, float (&sc)[4], float (&sh)[4]) {
  const int cq_n = C / 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cq * 4 + j;
    const double mean = stats[c] / (double)M;
    double var = stats[C + c] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    sc[j] = gamma[c] * invstd;
    sh[j] = beta[c] - (float)mean * sc[j];
    if (blockIdx.x == 0 && threadIdx.x < cq_n) {
      scsh[c] = sc[j];
      scsh[C + c] = sh[j];
      scsh[2 * C + c] = (float)mean;
      scsh[3 * C + c] = invstd;
      if (running_mean) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
}

__global__ void __launch_bounds__(kThreads) bn_apply_train_kernel(const float* __restrict__ z, long long zs, long long M, int C,
                                                                  const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* running_mean,
                                                                  float* running_var, long long* nbt, float eps, float momentum,
                                                                  float* __restrict__ scsh, int relu, float* __restrict__ out,
                                                                  long long os, __half* __restrict__ out16, int only16) {
  qeb_pdl_sync();
  const int cq_n = C / 4;
  const int cq = threadIdx.x % cq_n;
  float sc[4], sh[4];
  bn_train_scale_shift(M, C, cq, stats, gamma, beta, running_mean, running_var, nbt, eps, momentum, scsh, sc, sh);
  const long long rstep = (long long)gridDim.x * (kThreads / cq_n);
  for (long long r = (long long)blockIdx.x * (kThreads / cq_n) + threadIdx.x / cq_n; r < M; r += rstep) {
    const float4 v = ld4(z + r * zs + cq * 4);
    float4 o = make_float4(fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3]));
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (!only16) st4(out + r * os + cq * 4, qeb_tf32r4(o));   // weight-gradient operand: rounded to tf32 here
    if (out16) st4h(out16 + r * os + cq * 4, o);
  }
}

// the same (with ReLU) fused with the 2 x 2 max pooling that follows the unit in the UNet encoder blocks: one pass over z
// writes the activation (+ fp16 shadow) and the pooled activation (+ shadow). thread = (pooled pixel, channel quad)
__global__ void __launch_bounds__(kThreads) bn_apply_train_pool_kernel(const float* __restrict__ z, Geo gz, long long M, int C,
                                                                       const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, float* running_mean,
                                                                       float* running_var, long long* nbt, float eps, float momentum,
                                                                       float* __restrict__ scsh, float* __restrict__ out, Geo go,
                                                                       __half* __restrict__ out16, float* __restrict__ pool, Geo gp,
                                                                       __half* __restrict__ pool16, int n_pooled) {
  qeb_pdl_sync();
  const int cq_n = C / 4;
  const int cq = threadIdx.x % cq_n;
  float sc[4], sh[4];
  bn_train_scale_shift(M, C, cq, stats, gamma, beta, running_mean, running_var, nbt, eps, momentum, scsh, sc, sh);
  const int rstep = gridDim.x * (kThreads / cq_n);
  for (int r = blockIdx.x * (kThreads / cq_n) + threadIdx.x / cq_n; r < n_pooled; r += rstep) {
    const int wv = r % gp.w, t = r / gp.w, hv = t % gp.h;
    const long long n = t / gp.h;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const float4 v = ld4(z + n * gz.sn + (long long)(2 * hv + a) * gz.sh + (long long)(2 * wv + b) * gz.sw + cq * 4);
        const float4 o = make_float4(fmaxf(fmaf(v.x, sc[0], sh[0]), 0.f), fmaxf(fmaf(v.y, sc[1], sh[1]), 0.f),
                                     fmaxf(fmaf(v.z, sc[2], sh[2]), 0.f), fmaxf(fmaf(v.w, sc[3], sh[3]), 0.f));
        const long long oo = n * go.sn + (long long)(2 * hv + a) * go.sh + (long long)(2 * wv + b) * go.sw + cq * 4;
        st4(out + oo, qeb_tf32r4(o));
        if (out16) st4h(out16 + oo, o);
        if (a == 0 && b == 0) {
          m = o;
        } else {   // NaN-propagating max, the first maximum wins (as maxpool_fwd_kernel)
          m.x = (o.x > m.x || o.x != o.x) ? o.x : m.x;
          m.y = (o.y > m.y || o.y != o.y) ? o.y : m.y;
          m.z = (o.z > m.z || o.z != o.z) ? o.z : m.z;
          m.w = (o.w > m.w || o.w != o.w) ? o.w : m.w;
        }
      }
    const long long po = n * gp.sn + (long long)hv * gp.sh + (long long)wv * gp.sw + cq * 4;
    st4(pool + po, qeb_tf32r4(m));
    if (pool16) st4h(pool16 + po, m);
  }
}

__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const float* __restrict__ z, long long zs,
                                                                 const float* __restrict__ dy, long long ds, long long M, int C,
                                                                 const float* __restrict__ scsh, int relu,
                                                                 double* __restrict__ red) {
  qeb_pdl_sync();
  const int cq_n = C / 4, rpb = kThreads / cq_n;
  const int cq = threadIdx.x % cq_n, rl = threadIdx.x / cq_n;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < rpb) {
    const float4 sc = ld4(scsh + cq * 4), sh = ld4(scsh + C + cq * 4);
    const float4 mu = ld4(scsh + 2 * C + cq * 4), is = ld4(scsh + 3 * C + cq * 4);
    const long long step = (long long)gridDim.x * rpb;
    long long r = (long long)blockIdx.x * rpb + rl;
    for (; r + step < M; r += 2 * step) {  // two rows (four independent loads) in flight per thread
      const float4 v0 = ld4(z + r * zs + cq * 4), v1 = ld4(z + (r + step) * zs + cq * 4);
      const float4 d0 = ld4(dy + r * ds + cq * 4), d1 = ld4(dy + (r + step) * ds + cq * 4);
      const float4 g0 = bn_masked_grad(v0, d0, sc, sh, relu), g1 = bn_masked_grad(v1, d1, sc, sh, relu);
      s[0] += g0.x + g1.x; s[1] += g0.y + g1.y; s[2] += g0.z + g1.z; s[3] += g0.w + g1.w;
      q[0] = fmaf(g0.x, (v0.x - mu.x) * is.x, q[0]); q[1] = fmaf(g0.y, (v0.y - mu.y) * is.y, q[1]);
      q[2] = fmaf(g0.z, (v0.z - mu.z) * is.z, q[2]); q[3] = fmaf(g0.w, (v0.w - mu.w) * is.w, q[3]);
      q[0] = fmaf(g1.x, (v1.x - mu.x) * is.x, q[0]); q[1] = fmaf(g1.y, (v1.y - mu.y) * is.y, q[1]);
      q[2] = fmaf(g1.z, (v1.z - mu.z) * is.z, q[2]); q[3] = fmaf(g1.w, (v1.w - mu.w) * is.w, q[3]);
    }
    for (; r < M; r += step) {
      const float4 v = ld4(z + r * zs + cq * 4);
      const float4 g = bn_masked_grad(v, ld4(dy + r * ds + cq * 4), sc, sh, relu);
      s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
      q[0] = fmaf(g.x, (v.x - mu.x) * is.x, q[0]); q[1] = fmaf(g.y, (v.y - mu.y) * is.y, q[1]);
      q[2] = fmaf(g.z, (v.z - mu.z) * is.z, q[2]); q[3] = fmaf(g.w, (v.w - mu.w) * is.w, q[3]);
    }
  }
  extern __shared__ float sm[];
  float* ss = sm;
  float* sq = sm + rpb * C;
  if (rl < rpb) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { ss[rl * C + cq * 4 + j] = s[j]; sq[rl * C + cq * 4 + j] = q[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kThreads) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < rpb; ++i) { a += ss[i * C + c]; b += sq[i * C + c]; }
    atomicAdd(red + c, a);
    atomicAdd(red + C + c, b);
  }
}

// mode 0 (train): dz = gamma*invstd*(g - mean_g - xhat*mean_gx); mode 1 (eval): dz = g*scale
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const float* __restrict__ z, long long zs,
                                                                const float* __restrict__ dy, long long ds, long long M, int C,
                                                                const float* __restrict__ scsh, int relu,
                                                                const double* __restrict__ red, int mode,
                                                                float* __restrict__ dz, long long dzs, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, __half* __restrict__ dz16,
                                                                const float* __restrict__ gscale, unsigned* __restrict__ gamax, int only16) {
  qeb_pdl_sync();
  const float gs_s = dz16 ? __ldg(gscale) : 1.f;
  float gs_m = 0.f;
  const int cq_n = C / 4;
  if (blockIdx.x == 0 && dgamma) {
    for (int c = threadIdx.x; c < C; c += kThreads) {
      dgamma[c] += (float)red[C + c];
      dbeta[c] += (float)red[c];
    }
  }
  // kThreads is a multiple of C/4 (checked on the host): a thread keeps its channel quad over the grid-stride loop, so the
  // per-channel constants - including the fp64 -> fp32 conversions of the batch reductions - are formed once per thread and
  // the row index advances without a division
  const float invM = 1.f / (float)M;
  const int cq = threadIdx.x % cq_n;
  const float4 sc = ld4(scsh + cq * 4), sh = ld4(scsh + C + cq * 4);
  float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), is = mu, mg = mu, mx = mu;
  if (mode != 1) {
    mu = ld4(scsh + 2 * C + cq * 4); is = ld4(scsh + 3 * C + cq * 4);
    mg = make_float4((float)red[cq * 4] * invM, (float)red[cq * 4 + 1] * invM, (float)red[cq * 4 + 2] * invM, (float)red[cq * 4 + 3] * invM);
    mx = make_float4((float)red[C + cq * 4] * invM, (float)red[C + cq * 4 + 1] * invM, (float)red[C + cq * 4 + 2] * invM,
                     (float)red[C + cq * 4 + 3] * invM);
  }
  const long long rstep = (long long)gridDim.x * (kThreads / cq_n);
  for (long long r = (long long)blockIdx.x * (kThreads / cq_n) + threadIdx.x / cq_n; r < M; r += rstep) {
    const float4 v = ld4(z + r * zs + cq * 4);
    const float4 g = bn_masked_grad(v, ld4(dy + r * ds + cq * 4), sc, sh, relu);
    float4 o;
    if (mode == 1) {
      o = make_float4(g.x * sc.x, g.y * sc.y, g.z * sc.z, g.w * sc.w);
    } else {
      o.x = sc.x * (g.x - mg.x - (v.x - mu.x) * is.x * mx.x);
      o.y = sc.y * (g.y - mg.y - (v.y - mu.y) * is.y * mx.y);
      o.z = sc.z * (g.z - mg.z - (v.z - mu.z) * is.z * mx.z);
      o.w = sc.w * (g.w - mg.w - (v.w - mu.w) * is.w * mx.w);
    }
    if (!only16) st4(dz + r * dzs + cq * 4, qeb_tf32r4(o));   // dgrad / wgrad operand: rounded to tf32 here
    if (dz16) st4h_scaled(dz16 + r * dzs + cq * 4, o, gs_s);
    gs_m = amax4(gs_m, o);
  }
  if (gamax) amax_flush(gamax, gs_m);
}

__global__ void __launch_bounds__(kThreads) colsum_kernel(const float* __restrict__ x, long long xs, long long M, int C,
                                                          float* __restrict__ out, float* __restrict__ out2) {
  qeb_pdl_sync();
  const int cq_n = (C + 3) / 4, rpb = max(1, kThreads / cq_n);
  extern __shared__ float sm[];  // [rpb][cq_n*4]
  const int Cp = cq_n * 4;
  for (int cq0 = 0; cq0 < cq_n; cq0 += kThreads) {  // C up to 4*kThreads per pass
    const int cq = cq0 + threadIdx.x % min(cq_n, kThreads), rl = threadIdx.x / min(cq_n, kThreads);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (rl < rpb && cq < cq_n) {
      const long long step = (long long)gridDim.x * rpb;
      long long r = (long long)blockIdx.x * rpb + rl;
      if (cq * 4 + 3 < C) {
        for (; r + 3 * step < M; r += 4 * step) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ld4(x + (r + u * step) * xs + cq * 4);
#pragma unroll
          for (int u = 0; u < 4; ++u) { s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w; }
        }
        for (; r < M; r += step) {
          const float4 v = ld4(x + r * xs + cq * 4);
          s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
        }
      } else {
        for (; r < M; r += step) {
          const float* p = x + r * xs + cq * 4;
          for (int j = 0; cq * 4 + j < C; ++j) s[j] += p[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) sm[rl * Cp + cq * 4 + j] = s[j];
    }
    __syncthreads();
    for (int c = cq0 * 4 + threadIdx.x; c < min(C, (cq0 + kThreads) * 4); c += kThreads) {
      float a = 0.f;
      for (int i = 0; i < rpb; ++i) a += sm[i * Cp + c];
      atomicAdd(out + c, a);
      if (out2) atomicAdd(out2 + c, a);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) relu_bwd_kernel(const float* __restrict__ a, Geo ga, const float* __restrict__ dy,
                                                            Geo gd, float* __restrict__ dx, Geo gx, int cq_n, long long total) {
  qeb_pdl_sync();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int cq = (int)(i % cq_n);
    const long long pix = i / cq_n;
    const float4 av = ld4(a + pix_off(ga, pix) + cq * 4);
    const float4 g = ld4(dy + pix_off(gd, pix) + cq * 4);
    st4(dx + pix_off(gx, pix) + cq * 4,
        make_float4(av.x > 0.f ? g.x : 0.f, av.y > 0.f ? g.y : 0.f, av.z > 0.f ? g.z : 0.f, av.w > 0.f ? g.w : 0.f));
  }
}

__global__ void __launch_bounds__(kThreads) pack3d_kernel(const float* __restrict__ src, float* __restrict__ dst, int n0, int n1,
                                                          int n2, long long s0, long long s1, long long s2, long long d0,
                                                          long long d1) {
  qeb_pdl_sync();
  const long long total = (long long)n0 * n1 * n2;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    const int i2 = (int)(i % n2);
    const long long t = i / n2;
    const int i1 = (int)(t % n1);
    const long long i0 = t / n1;
    dst[i0 * d0 + i1 * d1 + i2] = src[i0 * s0 + i1 * s1 + (long long)i2 * s2];
  }
}

struct PackJobsParam {
  PackJob j[PackBatch::kMax];
  int n;
};
__global__ void __launch_bounds__(kThreads) pack_multi_kernel(const __grid_constant__ PackJobsParam jobs, long long total,
                                                              int accumulate) {
  qeb_pdl_sync();
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
    int lo = 0, hi = jobs.n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs.j[mid].start <= i) lo = mid; else hi = mid - 1;
    }
    const PackJob& j = jobs.j[lo];
    const long long e = i - j.start;
    const int i2 = (int)(e % j.n2);
    const long long t = e / j.n2;
    const int i1 = (int)(t % j.n1);
    const long long i0 = t / j.n1;
    const float v = j.src[i0 * j.s0 + i1 * j.s1 + (long long)i2 * j.s2];
    if (j.half_out) {   // fp16 operand copy (forward weights)
      reinterpret_cast<__half*>(j.dst)[i0 * j.d0 + i1 * j.d1 + i2] = __float2half_rn(v);
      continue;
    }
    float* d = j.dst + i0 * j.d0 + i1 * j.d1 + i2;
    *d = accumulate ? *d + v : qeb_tf32r(v);   // plain re-layouts are tensor-core operands: rounded to tf32 here
  }
}

struct ConvPackJobsParam {
  ConvPackJob j[PackBatch::kMax];
  int n;
};
// One block = one tile of 32 output channels (a) x 32 input channels (b) x taps of one job. Phase 1 reads the source
// in runs of 32 * taps (modes 0, 1) or 32 (mode 2) contiguous floats into smem[a][b * taps + tap] (row pitch 32 * taps + 1);
// phase 2 writes runs of 32 (modes 0, 1) or 32 * taps (mode 2) contiguous floats. All shared-memory strides (taps, pitch)
// are odd, so every access is bank-conflict free.
__global__ void __launch_bounds__(kThreads) conv_pack_kernel(const __grid_constant__ ConvPackJobsParam jobs, int accumulate) {
  qeb_pdl_sync();
  extern __shared__ float tile[];
  int ji = 0;
  while (ji + 1 < jobs.n && jobs.j[ji + 1].tile_start <= (int)blockIdx.x) ++ji;
  const ConvPackJob& j = jobs.j[ji];
  const int T = j.taps, pitch = 32 * T + 1, run = 32 * T;
  const int t = blockIdx.x - j.tile_start;
  const int tb = t % (j.B / 32), ta = t / (j.B / 32);
  const int a0 = ta * 32, b0 = tb * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = kThreads / 32;
  if (j.mode != 2) {
    for (int a = warp; a < 32; a += nwarp) {
      const float* s = j.src + ((long long)(a0 + a) * j.B + b0) * T;
      for (int e = lane; e < run; e += 32) tile[a * pitch + e] = __ldg(s + e);
    }
  } else {
    for (int r = warp; r < 32 * T; r += nwarp) {   // r = a * T + tap
      const int a = r / T, tap = r - a * T;
      tile[a * pitch + lane * T + tap] = __ldg(j.src + ((long long)(a0 + a) * T + tap) * j.B + b0 + lane);
    }
  }
  __syncthreads();
  if (j.mode == 0) {
    for (int r = warp; r < 32 * T; r += nwarp) {
      const int a = r / T, tap = r - a * T;
      const long long o = ((long long)(a0 + a) * T + tap) * j.B + b0 + lane;
      if (j.half_out) reinterpret_cast<__half*>(j.dst)[o] = __float2half_rn(tile[a * pitch + lane * T + tap]);
      else j.dst[o] = qeb_tf32r(tile[a * pitch + lane * T + tap]);
    }
  } else if (j.mode == 1) {
    for (int r = warp; r < 32 * T; r += nwarp) {
      const int b = r / T, ft = r - b * T;
      const long long o = ((long long)(b0 + b) * T + ft) * j.A + a0 + lane;
      if (j.half_out) reinterpret_cast<__half*>(j.dst)[o] = __float2half_rn(tile[lane * pitch + b * T + (T - 1 - ft)]);
      else j.dst[o] = qeb_tf32r(tile[lane * pitch + b * T + (T - 1 - ft)]);
    }
  } else {
    for (int a = warp; a < 32; a += nwarp) {
      float* d = j.dst + ((long long)(a0 + a) * j.B + b0) * T;
      for (int e = lane; e < run; e += 32) d[e] = accumulate ? d[e] + tile[a * pitch + e] : tile[a * pitch + e];
    }
  }
}

__global__ void vec_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  qeb_pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

// Grid of a grid-stride kernel = enough blocks for `total` threads, at most ONE resident wave (occupancy API, cached per
// kernel): with a fixed "8 blocks per SM" cap a kernel that fits 3 blocks per SM ran 2.7 waves, the last one at a third of
// the occupancy (ncu: c1_* kernels at 22-32 % active warps).
int resident_grid(const void* kernel, long long total, int threads, size_t smem = 0) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  int bps = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) bps = it->second;
  }
  if (!bps) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, smem) != cudaSuccess || bps < 1) { cudaGetLastError(); bps = 4; }
    std::lock_guard<std::mutex> lk(mu);
    cache[kernel] = bps;
  }
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * bps;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}
#define RGRID(kernel, total) resident_grid(reinterpret_cast<const void*>(kernel), (total), kThreads)

#define REQ_FLAT(img, what) QEB_REQUIRE(img_flat(img) && vec4_ok(img), what ": tensor must be a packed NHWC view with 16-byte aligned rows")

}  // namespace

int fill_zero(void* p, size_t bytes, cudaStream_t st) {
  ProfScope prof("memset", st, 0.0, (double)bytes);
  QEB_CUDA(cudaMemsetAsync(p, 0, bytes, st));
  return QEB_OK;
}

int pack_3d(const float* src, float* dst, int n0, int n1, int n2, long long s0, long long s1, long long s2, long long d0,
            long long d1, cudaStream_t st) {
  ProfScope prof("pack", st, 0.0, 8.0 * n0 * n1 * n2);
  const long long total = (long long)n0 * n1 * n2;
  if (total == 0) return QEB_OK;
  QEB_CUDA(qeb_launch(pack3d_kernel, qeb_grid(total, kThreads), kThreads, 0, st, src, dst, n0, n1, n2, s0, s1, s2, d0, d1));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int c1_conv_fwd(const Img& x, const float* w, const float* bias, int relu, const Img& out, cudaStream_t st) {
  ProfScope prof("c1_conv_fwd", st, 18.0 * (double)img_pixels(x) * out.c, 4.0 * (double)img_pixels(x) * (1 + out.c));
  QEB_REQUIRE(x.c == 1 && (out.c == 32 || out.c == 64), "c1_conv_fwd: 1 -> 32/64 channels only (got %d -> %d)", x.c, out.c);
  QEB_REQUIRE(x.n == out.n && x.h == out.h && x.w == out.w && vec4_ok(out), "c1_conv_fwd: geometry/alignment");
  QEB_REQUIRE(x.w % 4 == 0 && img_pixels(x) / 4 < (1ll << 30), "c1_conv_fwd: the width must be a multiple of 4");
  const int n_quads = (int)(img_pixels(x) / 4);
  const long long nthr = (long long)n_quads * (out.c / 4);
  const int g = out.c == 32 ? RGRID(c1_fwd_kernel<32>, nthr) : RGRID(c1_fwd_kernel<64>, nthr);
  if (out.c == 32) QEB_CUDA(qeb_launch(c1_fwd_kernel<32>, g, kThreads, 0, st, x.p, geo(x), w, bias, relu, out.p, geo(out), n_quads));
  else QEB_CUDA(qeb_launch(c1_fwd_kernel<64>, g, kThreads, 0, st, x.p, geo(x), w, bias, relu, out.p, geo(out), n_quads));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int c1_conv_fwd_jitter(const float* x, int n_img, int h, int w, const JitterArgs& j, const float* wgt, const float* bias, int relu,
                       const Img& out, cudaStream_t st) {
  ProfScope prof("c1_conv_fwd", st, 18.0 * (double)n_img * h * w * out.c, 4.0 * (double)n_img * h * w * (2 + out.c));
  QEB_REQUIRE(x && j.sigma && (out.c == 32 || out.c == 64), "c1_conv_fwd_jitter: 1 -> 32/64 channels, sigma per image required");
  QEB_REQUIRE(out.n == n_img && out.h == h && out.w == w && vec4_ok(out), "c1_conv_fwd_jitter: geometry/alignment");
  QEB_REQUIRE(w % 4 == 0 && w <= 2048 && (long long)h * (w / 4) < (1ll << 31), "c1_conv_fwd_jitter: the width must be a multiple of 4, <= 2048");
  QEB_REQUIRE((((uintptr_t)x | (uintptr_t)j.noisy_out | (uintptr_t)j.noise_out) & 15) == 0, "c1_conv_fwd_jitter: 16-byte aligned images");
  const int bands = (h + kJitRows - 1) / kJitRows;
  const size_t smem = (size_t)(kJitRows + 2) * (w + 8) * sizeof(float);
  const long long grid = (long long)n_img * bands;
  QEB_REQUIRE(grid < (1ll << 31), "c1_conv_fwd_jitter: too many images");
  if (out.c == 32) {
    if (smem > 48 * 1024) QEB_CUDA(cudaFuncSetAttribute(c1_jitter_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QEB_CUDA(qeb_launch(c1_jitter_fwd_kernel<32>, (int)grid, kThreads, smem, st, x, h, w, j.sigma, j.mean, j.coef, j.seed, j.seed_dev,
                        j.noisy_out, j.noise_out, wgt, bias, relu, out.p, geo(out)));
  } else {
    if (smem > 48 * 1024) QEB_CUDA(cudaFuncSetAttribute(c1_jitter_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QEB_CUDA(qeb_launch(c1_jitter_fwd_kernel<64>, (int)grid, kThreads, smem, st, x, h, w, j.sigma, j.mean, j.coef, j.seed, j.seed_dev,
                        j.noisy_out, j.noise_out, wgt, bias, relu, out.p, geo(out)));
  }
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int c1_conv_wgrad(const Img& x, const Img& dy, float* dw, float* dbias, cudaStream_t st) {
  ProfScope prof("c1_conv_wgrad", st, 18.0 * (double)img_pixels(x) * dy.c, 4.0 * (double)img_pixels(x) * (1 + dy.c));
  QEB_REQUIRE(x.c == 1 && (dy.c == 32 || dy.c == 64), "c1_conv_wgrad: 1 -> 32/64 channels only");
  QEB_REQUIRE(x.n == dy.n && x.h == dy.h && x.w == dy.w && vec4_ok(dy), "c1_conv_wgrad: geometry/alignment");
  QEB_REQUIRE(x.w % 4 == 0 && img_pixels(x) / 4 < (1ll << 30), "c1_conv_wgrad: the width must be a multiple of 4");
  const int n_quads = (int)(img_pixels(x) / 4);
  const long long nthr = (long long)n_quads * (dy.c / 4);
  const int g = dy.c == 32 ? RGRID(c1_wgrad_kernel<32>, nthr) : RGRID(c1_wgrad_kernel<64>, nthr);
  if (dy.c == 32) QEB_CUDA(qeb_launch(c1_wgrad_kernel<32>, g, kThreads, 0, st, x.p, geo(x), dy.p, geo(dy), dw, dbias, n_quads));
  else QEB_CUDA(qeb_launch(c1_wgrad_kernel<64>, g, kThreads, 0, st, x.p, geo(x), dy.p, geo(dy), dw, dbias, n_quads));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int c1_conv_dgrad(const Img& dy, const float* w, const Img& dx, cudaStream_t st) {
  ProfScope prof("c1_conv_dgrad", st, 18.0 * (double)img_pixels(dx) * dy.c, 4.0 * (double)img_pixels(dx) * (1 + dy.c));
  QEB_REQUIRE(dx.c == 1 && (dy.c == 32 || dy.c == 64), "c1_conv_dgrad: 1 <- 32/64 channels only");
  QEB_REQUIRE(dx.n == dy.n && dx.h == dy.h && dx.w == dy.w && vec4_ok(dy), "c1_conv_dgrad: geometry/alignment");
  QEB_REQUIRE(dx.w % 4 == 0 && img_pixels(dx) / 4 < (1ll << 30), "c1_conv_dgrad: the width must be a multiple of 4");
  const int n_quads = (int)(img_pixels(dx) / 4);
  const long long nthr = (long long)n_quads * (dy.c / 4);
  const int g = dy.c == 32 ? RGRID(c1_dgrad_kernel<32>, nthr) : RGRID(c1_dgrad_kernel<64>, nthr);
  if (dy.c == 32) QEB_CUDA(qeb_launch(c1_dgrad_kernel<32>, g, kThreads, 0, st, dy.p, geo(dy), w, dx.p, geo(dx), n_quads));
  else QEB_CUDA(qeb_launch(c1_dgrad_kernel<64>, g, kThreads, 0, st, dy.p, geo(dy), w, dx.p, geo(dx), n_quads));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int o1_conv_sigmoid_fwd(const Img& x, const float* w, const float* b, float* y, cudaStream_t st) {
  ProfScope prof("o1_conv_sigmoid_fwd", st, 2.0 * (double)img_pixels(x) * x.c, 4.0 * (double)img_pixels(x) * (1 + x.c));
  QEB_REQUIRE(x.c == 32 && vec4_ok(x), "o1_conv_sigmoid_fwd: 32 input channels, aligned");
  const long long n_pix = img_pixels(x);
  QEB_CUDA(qeb_launch(o1_fwd_kernel<32>, RGRID(o1_fwd_kernel<32>, n_pix * 8), kThreads, 0, st, x.p, geo(x), w, b, y, n_pix));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int o1_conv_sigmoid_bwd(const Img& x, const float* w, const float* y, const float* dy, const Img& dx, float* dw, float* db,
                        cudaStream_t st) {
  ProfScope prof("o1_conv_sigmoid_bwd", st, 4.0 * (double)img_pixels(x) * x.c, 4.0 * (double)img_pixels(x) * (2 + 2 * x.c));
  QEB_REQUIRE(x.c == 32 && dx.c == 32 && vec4_ok(x) && vec4_ok(dx), "o1_conv_sigmoid_bwd: 32 channels, aligned");
  const long long n_pix = img_pixels(x);
  QEB_CUDA(qeb_launch(o1_bwd_kernel<32>, RGRID(o1_bwd_kernel<32>, n_pix * 8), kThreads, 0, st, x.p, geo(x), w, y, dy, dx.p, geo(dx), dw, db, n_pix));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int maxpool_fwd(const Img& x, int ph, int pw, const Img& out, cudaStream_t st, void* out16) {
  ProfScope prof("maxpool_fwd", st, 0.0, 4.0 * x.c * ((double)img_pixels(x) + (double)img_pixels(out)));
  QEB_REQUIRE(vec4_ok(x) && vec4_ok(out) && x.c == out.c, "maxpool_fwd: channel count / alignment");
  QEB_REQUIRE(x.h == out.h * ph && x.w == out.w * pw && x.n == out.n, "maxpool_fwd: input must be a multiple of the window");
  const int cq_n = x.c / 4;
  const long long total = img_pixels(out) * cq_n;
  QEB_REQUIRE(total < (1ll << 31), "maxpool_fwd: tensor too large for 32-bit indexing");
  const int g = qeb_grid(total, kThreads);
  __half* o16 = static_cast<__half*>(out16);
  if (ph == 2 && pw == 2) QEB_CUDA(qeb_launch(maxpool_fwd_kernel<2, 2>, g, kThreads, 0, st, x.p, geo(x), out.p, geo(out), cq_n, total, o16));
  else if (ph == 2 && pw == 1) QEB_CUDA(qeb_launch(maxpool_fwd_kernel<2, 1>, g, kThreads, 0, st, x.p, geo(x), out.p, geo(out), cq_n, total, o16));
  else QEB_REQUIRE(false, "maxpool_fwd: window %dx%d not supported", ph, pw);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int maxpool_bwd(const Img& x, const Img& dy, int ph, int pw, int relu_mask, const float* chan_scale, const Img* add,
                const Img& dx, cudaStream_t st, const Img* bn_z, const float* bn_scsh, double* bn_red, const GradShadow* gs) {
  __half* dx16 = gs && gs->out16 && gs->scale ? static_cast<__half*>(gs->out16) : nullptr;
  const float* gsc = gs ? gs->scale : nullptr;
  unsigned* gam = gs ? gs->amax : nullptr;
  ProfScope prof("maxpool_bwd", st, 0.0, 4.0 * x.c * (2 * (double)img_pixels(x) + (double)img_pixels(dy) + (add ? (double)img_pixels(x) : 0.0)));
  QEB_REQUIRE(vec4_ok(x) && vec4_ok(dy) && vec4_ok(dx) && x.c == dy.c && x.c == dx.c, "maxpool_bwd: channel count / alignment");
  QEB_REQUIRE(x.h == dy.h * ph && x.w == dy.w * pw && x.n == dy.n, "maxpool_bwd: input must be a multiple of the window");
  QEB_REQUIRE(!add || (vec4_ok(*add) && add->c == x.c), "maxpool_bwd: add tensor");
  const int cq_n = x.c / 4;
  const long long total = img_pixels(dy) * cq_n;
  QEB_REQUIRE(total < (1ll << 31), "maxpool_bwd: tensor too large for 32-bit indexing");
  const bool red = bn_z && bn_scsh && bn_red;
  if (!red) bn_red = nullptr;
  int g = (ph == 2 && pw == 2) ? (red ? RGRID((maxpool_bwd_kernel<2, 2, true>), total) : RGRID((maxpool_bwd_kernel<2, 2, false>), total))
                               : RGRID((maxpool_bwd_kernel<2, 1, false>), total);
  Geo ga = add ? geo(*add) : geo(x);
  const float* ap = add ? add->p : nullptr;
  size_t smem = 0;
  if (red) {
    QEB_REQUIRE(vec4_ok(*bn_z) && bn_z->c == x.c && bn_z->n == x.n && bn_z->h == x.h && bn_z->w == x.w && kThreads % cq_n == 0,
                "maxpool_bwd: fused BatchNorm reductions need z of the pooled tensor's shape and C/4 dividing %d", kThreads);
    smem = (size_t)2 * (kThreads / cq_n) * x.c * sizeof(float);
    g = min(g, 4 * kNumSMs);   // every block ends with 2C double atomics: a few blocks per SM, grid-stride over the rest
  }
  const Geo gz = red ? geo(*bn_z) : geo(x);
  const float* zp = red ? bn_z->p : nullptr;
  QEB_REQUIRE(!red || (ph == 2 && pw == 2), "maxpool_bwd: fused BatchNorm reductions exist for the 2x2 window only");
  if (ph == 2 && pw == 2 && red)
    QEB_CUDA(qeb_launch(maxpool_bwd_kernel<2, 2, true>, g, kThreads, smem, st, x.p, geo(x), dy.p, geo(dy), relu_mask, chan_scale, ap, ga, dx.p, geo(dx),
                        cq_n, total, zp, gz, bn_scsh, bn_red, dx16, gsc, gam));
  else if (ph == 2 && pw == 2)
    QEB_CUDA(qeb_launch(maxpool_bwd_kernel<2, 2, false>, g, kThreads, smem, st, x.p, geo(x), dy.p, geo(dy), relu_mask, chan_scale, ap, ga, dx.p, geo(dx),
                        cq_n, total, zp, gz, bn_scsh, (double*)nullptr, dx16, gsc, gam));
  else if (ph == 2 && pw == 1)
    QEB_CUDA(qeb_launch(maxpool_bwd_kernel<2, 1, false>, g, kThreads, smem, st, x.p, geo(x), dy.p, geo(dy), relu_mask, chan_scale, ap, ga, dx.p, geo(dx),
                        cq_n, total, zp, gz, bn_scsh, (double*)nullptr, dx16, gsc, gam));
  else QEB_REQUIRE(false, "maxpool_bwd: window %dx%d not supported", ph, pw);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

// grid of a column-reduction kernel. Every block ends with one atomic per channel (and statistic), so the block count is
// what the serialised L2 atomics scale with: give each block at least 64 KB of input (measured: with one block per
// `rpb` rows the atomics, not the reads, set the run time of the smaller tensors).
static int reduce_grid(long long M, int rpb, int C) {
  const long long min_rows = max(1LL, (64LL << 10) / ((long long)C * 4));
  long long g = (M + min_rows - 1) / min_rows;
  const long long by_rows = (M + rpb - 1) / rpb;
  if (g > by_rows) g = by_rows;
  const long long cap = 4 * kNumSMs;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

int bn_train_stats(const Img& z, double* stats, cudaStream_t st) {
  ProfScope prof("bn_stats", st, 0.0, 4.0 * z.c * (double)img_pixels(z));
  REQ_FLAT(z, "bn_train_stats");
  QEB_REQUIRE(z.c <= 1024, "bn_train_stats: at most 1024 channels");
  const long long M = img_pixels(z);
  const int rpb = kThreads / (z.c / 4);
  QEB_CUDA(qeb_launch(bn_stats_kernel, reduce_grid(M, rpb, z.c), kThreads, 2 * rpb * z.c * sizeof(float), st, z.p, z.sw, M, z.c, stats));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_train_finalize(const double* stats, long long count, int c, const BnParams& bn, float* scsh, cudaStream_t st) {
  ProfScope prof("bn_finalize", st);
  QEB_CUDA(qeb_launch(bn_finalize_kernel, qeb_cdiv(c, 128), 128, 0, st, stats, count, c, bn.gamma, bn.beta, bn.running_mean, bn.running_var,
                                                       bn.num_batches_tracked, bn.eps, bn.momentum, scsh));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_eval_scsh(int c, const BnParams& bn, const float* conv_bias, float* scsh, cudaStream_t st) {
  ProfScope prof("bn_finalize", st);
  QEB_CUDA(qeb_launch(bn_eval_scsh_kernel, qeb_cdiv(c, 128), 128, 0, st, c, bn.gamma, bn.beta, bn.running_mean, bn.running_var, bn.eps, conv_bias,
                                                        scsh));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_train_finalize_apply(const Img& z, const double* stats, const BnParams& bn, float* scsh, int relu, const Img& out,
                            cudaStream_t st, void* out16, int only16) {
  ProfScope prof("bn_apply", st, 0.0, 8.0 * z.c * (double)img_pixels(z));
  REQ_FLAT(z, "bn_train_finalize_apply");
  REQ_FLAT(out, "bn_train_finalize_apply");
  QEB_REQUIRE(z.c == out.c && img_pixels(z) == img_pixels(out), "bn_train_finalize_apply: shape mismatch");
  QEB_REQUIRE(kThreads % (z.c / 4) == 0, "bn_train_finalize_apply: C/4 must divide %d", kThreads);
  const long long M = img_pixels(z);
  QEB_CUDA(qeb_launch(bn_apply_train_kernel, RGRID(bn_apply_train_kernel, M * (z.c / 4)), kThreads, 0, st, z.p, z.sw, M, z.c, stats, bn.gamma, bn.beta,
                                                                               bn.running_mean, bn.running_var,
                                                                               bn.num_batches_tracked, bn.eps, bn.momentum, scsh,
                                                                               relu, out.p, out.sw, static_cast<__half*>(out16),
                                                                               (only16 && out16) ? 1 : 0));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_train_finalize_apply_pool(const Img& z, const double* stats, const BnParams& bn, float* scsh, const Img& out, void* out16,
                                 const Img& pool, void* pool16, cudaStream_t st) {
  ProfScope prof("bn_apply", st, 0.0, (8.0 * z.c + 1.0 * z.c) * (double)img_pixels(z));
  QEB_REQUIRE(vec4_ok(z) && vec4_ok(out) && vec4_ok(pool) && z.c == out.c && z.c == pool.c, "bn_train_finalize_apply_pool: channels / alignment");
  QEB_REQUIRE(z.n == out.n && z.h == out.h && z.w == out.w && pool.n == z.n && z.h == 2 * pool.h && z.w == 2 * pool.w,
              "bn_train_finalize_apply_pool: shapes (the pooled tensor is half the unit's height and width)");
  QEB_REQUIRE(kThreads % (z.c / 4) == 0, "bn_train_finalize_apply_pool: C/4 must divide %d", kThreads);
  const long long M = img_pixels(z), np = img_pixels(pool);
  QEB_REQUIRE(np * (z.c / 4) < (1ll << 31), "bn_train_finalize_apply_pool: tensor too large for 32-bit indexing");
  QEB_CUDA(qeb_launch(bn_apply_train_pool_kernel, RGRID(bn_apply_train_pool_kernel, np * (z.c / 4)), kThreads, 0, st, z.p, geo(z), M, z.c, stats, bn.gamma,
                      bn.beta, bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.eps, bn.momentum, scsh, out.p, geo(out),
                      static_cast<__half*>(out16), pool.p, geo(pool), static_cast<__half*>(pool16), (int)np));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_apply(const Img& z, const float* scsh, int relu, const Img& out, cudaStream_t st, void* out16) {
  ProfScope prof("bn_apply", st, 0.0, 8.0 * z.c * (double)img_pixels(z));
  REQ_FLAT(z, "bn_apply");
  REQ_FLAT(out, "bn_apply");
  QEB_REQUIRE(z.c == out.c && img_pixels(z) == img_pixels(out), "bn_apply: shape mismatch");
  QEB_REQUIRE(kThreads % (z.c / 4) == 0, "bn_apply: C/4 must divide %d", kThreads);
  const long long M = img_pixels(z);
  QEB_CUDA(qeb_launch(bn_apply_kernel, RGRID(bn_apply_kernel, M * (z.c / 4)), kThreads, 0, st, z.p, z.sw, M, z.c, scsh, relu, out.p, out.sw,
                                                                          static_cast<__half*>(out16)));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_bwd_reduce(const Img& z, const Img& dy, const float* scsh, int relu, double* red, cudaStream_t st) {
  ProfScope prof("bn_bwd_reduce", st, 0.0, 8.0 * z.c * (double)img_pixels(z));
  REQ_FLAT(z, "bn_bwd_reduce");
  REQ_FLAT(dy, "bn_bwd_reduce");
  QEB_REQUIRE(z.c == dy.c && img_pixels(z) == img_pixels(dy) && z.c <= 1024, "bn_bwd_reduce: shape mismatch");
  const long long M = img_pixels(z);
  const int rpb = kThreads / (z.c / 4);
  QEB_CUDA(qeb_launch(bn_bwd_reduce_kernel, reduce_grid(M, rpb, 2 * z.c), kThreads, 2 * rpb * z.c * sizeof(float), st, z.p, z.sw, dy.p, dy.sw, M, z.c, scsh,
                                                                                           relu, red));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

static int bn_bwd_apply(const Img& z, const Img& dy, const float* scsh, int relu, const double* red, int mode, const Img& dz,
                        float* dgamma, float* dbeta, cudaStream_t st, const GradShadow* gs) {
  __half* dz16 = gs && gs->out16 && gs->scale ? static_cast<__half*>(gs->out16) : nullptr;
  QEB_REQUIRE(!dz16 || (((uintptr_t)dz16 & 7) == 0), "bn_bwd_apply: fp16 shadow alignment");
  ProfScope prof("bn_bwd_apply", st, 0.0, 12.0 * z.c * (double)img_pixels(z));
  REQ_FLAT(z, "bn_bwd_apply");
  REQ_FLAT(dy, "bn_bwd_apply");
  REQ_FLAT(dz, "bn_bwd_apply");
  QEB_REQUIRE(z.c == dy.c && z.c == dz.c && img_pixels(z) == img_pixels(dy) && img_pixels(z) == img_pixels(dz),
              "bn_bwd_apply: shape mismatch");
  QEB_REQUIRE(kThreads % (z.c / 4) == 0, "bn_bwd_apply: C/4 must divide %d", kThreads);
  const long long M = img_pixels(z);
  QEB_CUDA(qeb_launch(bn_bwd_apply_kernel, RGRID(bn_bwd_apply_kernel, M * (z.c / 4)), kThreads, 0, st, z.p, z.sw, dy.p, dy.sw, M, z.c, scsh, relu, red, mode,
                                                                             dz.p, dz.sw, dgamma, dbeta, dz16, gs ? gs->scale : nullptr,
                                                                             gs ? gs->amax : nullptr, (dz16 && gs->only16) ? 1 : 0));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int bn_bwd_apply_train(const Img& z, const Img& dy, const float* scsh, int relu, const double* red, const float* gamma,
                       const Img& dz, float* dgamma, float* dbeta, cudaStream_t st, const GradShadow* gs) {
  (void)gamma;  // scale = gamma*invstd is already in scsh
  return bn_bwd_apply(z, dy, scsh, relu, red, 0, dz, dgamma, dbeta, st, gs);
}

int bn_bwd_apply_eval(const Img& z, const Img& dy, const float* scsh, int relu, const double* red, const Img& dz,
                      float* dgamma, float* dbeta, cudaStream_t st, const GradShadow* gs) {
  return bn_bwd_apply(z, dy, scsh, relu, red, 1, dz, red ? dgamma : nullptr, red ? dbeta : nullptr, st, gs);
}

int colsum_acc(const Img& x, float* out, cudaStream_t st, float* out2) {
  ProfScope prof("colsum", st, 0.0, 4.0 * x.c * (double)img_pixels(x));
  QEB_REQUIRE(img_flat(x) && ((uintptr_t)x.p & 15) == 0 && x.sw % 4 == 0, "colsum_acc: packed, 16-byte aligned rows");
  const long long M = img_pixels(x);
  const int cq_n = (x.c + 3) / 4;
  const int rpb = max(1, kThreads / cq_n);
  QEB_CUDA(qeb_launch(colsum_kernel, reduce_grid(M, rpb, x.c), kThreads, (size_t)rpb * cq_n * 4 * sizeof(float), st, x.p, x.sw, M, x.c, out, out2));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int relu_bwd(const Img& a, const Img& dy, const Img& dx, cudaStream_t st) {
  ProfScope prof("relu_bwd", st, 0.0, 12.0 * a.c * (double)img_pixels(a));
  QEB_REQUIRE(vec4_ok(a) && vec4_ok(dy) && vec4_ok(dx) && a.c == dy.c && a.c == dx.c, "relu_bwd: channel count / alignment");
  const int cq_n = a.c / 4;
  const long long total = img_pixels(a) * cq_n;
  QEB_CUDA(qeb_launch(relu_bwd_kernel, qeb_grid(total, kThreads), kThreads, 0, st, a.p, geo(a), dy.p, geo(dy), dx.p, geo(dx), cq_n, total));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int vec_add(const float* a, const float* b, float* out, int n, cudaStream_t st) {
  ProfScope prof("vec_add", st, 0.0, 12.0 * n);
  QEB_CUDA(qeb_launch(vec_add_kernel, qeb_cdiv(n, 256), 256, 0, st, a, b, out, n));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int pack_flush(PackBatch& b, cudaStream_t st) {
  QEB_REQUIRE(b.n <= PackBatch::kMax && b.nc <= PackBatch::kMax, "pack_flush: too many jobs");
  long long conv_elems = 0;
  for (int i = 0; i < b.nc; ++i) conv_elems += (long long)b.cjobs[i].A * b.cjobs[i].B * b.cjobs[i].taps;
  ProfScope prof("pack", st, 0.0, 8.0 * (b.total + conv_elems));
  if (b.n > 0) {
    PackJobsParam p;
    for (int i = 0; i < b.n; ++i) p.j[i] = b.jobs[i];
    p.n = b.n;
    QEB_CUDA(qeb_launch(pack_multi_kernel, qeb_grid(b.total, kThreads), kThreads, 0, st, p, b.total, b.accumulate));
    QEB_LAUNCH_CHECK();
    qeb_count_launch();
  }
  if (b.nc > 0) {
    ConvPackJobsParam p;
    for (int i = 0; i < b.nc; ++i) p.j[i] = b.cjobs[i];
    p.n = b.nc;
    QEB_CUDA(qeb_launch(conv_pack_kernel, b.ctiles, kThreads, (32 * (32 * 9 + 1)) * sizeof(float), st, p, b.accumulate));
    QEB_LAUNCH_CHECK();
    qeb_count_launch();
  }
  b.n = b.nc = b.ctiles = 0;
  b.total = 0;
  b.accumulate = 0;
  return QEB_OK;
}
