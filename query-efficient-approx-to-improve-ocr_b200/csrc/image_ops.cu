// Gaussian jitter and crop/pad of text strips. HBM-bound elementwise/gather kernels.
//
// Replaces
//   AddGaussianNoice.__call__  transform_helper.py:33-45 (+ add_noise loops train_nn_patch.py:187-191,
//                              train_nn_area.py:184-191): out = clamp(img - coef * N(mean, sigma_img), 0, 1)
//   get_text_stack / padder    utils.py:118-141: crop each bbox of a (1,H,W) image, centre-pad with 1.0 to (h,w)
//
// Jitter: one sigma per image (the reference draws r_std per image on the host; that stays a host draw).
// Noise is either supplied (exact restatement of the reference arithmetic on a given noise tensor) or generated
// in-kernel with Philox4x32-10 + Box-Muller: counter = (group index within image, image index, 0, 0),
// key = (seed_lo, seed_hi), each counter yields the 4 normals of 4 consecutive pixels (float4 load/store).
#include "common.cuh"
#include "philox.cuh"

namespace {

using namespace qebrng;

// hw4 = pixels per image / 4
__global__ void jitter_kernel(const float4* __restrict__ img, const float* __restrict__ sigma, float mean, float coef,
                              const float4* __restrict__ noise_in, unsigned long long seed,
                              const unsigned long long* __restrict__ seed_dev, long long n_img, int hw4,
                              float4* __restrict__ out, float4* __restrict__ noise_out) {
  if (seed_dev) seed += *seed_dev;   // seed kept in device memory: a captured CUDA graph draws fresh noise on every replay
  const long long total = n_img * hw4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long im = i / hw4;
    const unsigned int g = (unsigned int)(i - im * hw4);
    float4 z;
    if (noise_in) {
      z = noise_in[i];
    } else {
      z = noise4(g, im, seed, mean, sigma[im]);
    }
    const float4 p = img[i];
    out[i] = make_float4(jitter1(p.x, z.x, coef), jitter1(p.y, z.y, coef), jitter1(p.z, z.z, coef),
                         jitter1(p.w, z.w, coef));
    if (noise_out) noise_out[i] = z;
  }
}

// boxes: (n,4) int32 x_min,y_min,x_max,y_max (python slice semantics, clipped to the image).
// out[i,y,x] = crop_i[y - pad_top, x - pad_left] inside the crop, else 1.0 (negative pads crop, as ConstantPad2d does)
__global__ void crop_pad_gather_kernel(const float* __restrict__ img, int H, int W, const int* __restrict__ boxes, int n,
                                       int oh, int ow, float* __restrict__ out) {
  const int per = oh * ow;
  const long long total = (long long)n * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per), r = (int)(i - (long long)b * per);
    const int y = r / ow, x = r - y * ow;
    const int x0 = min(max(boxes[4 * b + 0], 0), W), y0 = min(max(boxes[4 * b + 1], 0), H);
    const int x1 = min(max(boxes[4 * b + 2], x0), W), y1 = min(max(boxes[4 * b + 3], y0), H);
    const int cw = x1 - x0, ch = y1 - y0;
    // floor division like python's //
    const int dl = ow - cw, dt = oh - ch;
    const int pl = (dl >= 0) ? dl / 2 : -((-dl + 1) / 2);
    const int pt = (dt >= 0) ? dt / 2 : -((-dt + 1) / 2);
    const int cy = y - pt, cx = x - pl;
    float v = 1.0f;
    if (cy >= 0 && cy < ch && cx >= 0 && cx < cw) v = img[(long long)(y0 + cy) * W + (x0 + cx)];
    out[i] = v;
  }
}

// backward: scatter-add grad of the strips into the image gradient (boxes may overlap -> atomics)
__global__ void crop_pad_scatter_kernel(const float* __restrict__ gout, int H, int W, const int* __restrict__ boxes,
                                        int n, int oh, int ow, float* __restrict__ gimg) {
  const int per = oh * ow;
  const long long total = (long long)n * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per), r = (int)(i - (long long)b * per);
    const int y = r / ow, x = r - y * ow;
    const int x0 = min(max(boxes[4 * b + 0], 0), W), y0 = min(max(boxes[4 * b + 1], 0), H);
    const int x1 = min(max(boxes[4 * b + 2], x0), W), y1 = min(max(boxes[4 * b + 3], y0), H);
    const int cw = x1 - x0, ch = y1 - y0;
    const int dl = ow - cw, dt = oh - ch;
    const int pl = (dl >= 0) ? dl / 2 : -((-dl + 1) / 2);
    const int pt = (dt >= 0) ? dt / 2 : -((-dt + 1) / 2);
    const int cy = y - pt, cx = x - pl;
    if (cy >= 0 && cy < ch && cx >= 0 && cx < cw) atomicAdd(&gimg[(long long)(y0 + cy) * W + (x0 + cx)], gout[i]);
  }
}

}  // namespace

// img/out/noise: (n_img, hw) fp32 contiguous, hw % 4 == 0, 16-byte aligned. noise_in NULL => Philox noise from seed.
static int gauss_jitter_impl(const float* img, const float* sigma, float mean, float coef, const float* noise_in,
                             unsigned long long seed, const unsigned long long* seed_dev, long long n_img, int hw, float* out,
                             float* noise_out, void* stream) {
  if (n_img == 0) return QEB_OK;
  QEB_REQUIRE(img && out && n_img > 0 && hw > 0, "gauss_jitter: bad args");
  QEB_REQUIRE(noise_in || sigma, "gauss_jitter: need sigma when noise is generated");
  QEB_REQUIRE(hw % 4 == 0, "gauss_jitter: pixels per image (%d) must be a multiple of 4", hw);
  QEB_REQUIRE(((uintptr_t)img | (uintptr_t)out | (uintptr_t)noise_in | (uintptr_t)noise_out) % 16 == 0,
              "gauss_jitter: buffers must be 16-byte aligned");
  const long long total = n_img * (hw / 4);
  const int grid = qeb_grid(total, 256);
  ProfScope prof("jitter", (cudaStream_t)stream, 0.0, 8.0 * n_img * hw);
  jitter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)img, sigma, mean, coef, (const float4*)noise_in,
                                                        seed, seed_dev, n_img, hw / 4, (float4*)out, (float4*)noise_out);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

QEB_API int qeb_gauss_jitter(const float* img, const float* sigma, float mean, float coef, const float* noise_in,
                             unsigned long long seed, long long n_img, int hw, float* out, float* noise_out,
                             void* stream) {
  return gauss_jitter_impl(img, sigma, mean, coef, noise_in, seed, nullptr, n_img, hw, out, noise_out, stream);
}

// the same with the Philox key taken from DEVICE memory (key = seed + *seed_dev): the form a captured CUDA graph needs,
// where a by-value seed would be frozen into the graph and every replay would repeat the noise
QEB_API int qeb_gauss_jitter_devseed(const float* img, const float* sigma, float mean, float coef,
                                     const unsigned long long* seed_dev, unsigned long long seed_offset, long long n_img,
                                     int hw, float* out, float* noise_out, void* stream) {
  QEB_REQUIRE(seed_dev, "gauss_jitter_devseed: null seed pointer");
  return gauss_jitter_impl(img, sigma, mean, coef, nullptr, seed_offset, seed_dev, n_img, hw, out, noise_out, stream);
}

QEB_API int qeb_crop_pad_gather(const float* img, int H, int W, const int* boxes, int n, int oh, int ow, float* out,
                                void* stream) {
  if (n == 0) return QEB_OK;
  QEB_REQUIRE(img && boxes && out && H > 0 && W > 0 && oh > 0 && ow > 0 && n > 0, "crop_pad_gather: bad args");
  const long long total = (long long)n * oh * ow;
  const int grid = qeb_grid(total, 256);
  crop_pad_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, H, W, boxes, n, oh, ow, out);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

// gimg (H,W) must be zero-initialised (or hold a gradient to accumulate into)
QEB_API int qeb_crop_pad_scatter(const float* gout, int H, int W, const int* boxes, int n, int oh, int ow, float* gimg,
                                 void* stream) {
  if (n == 0) return QEB_OK;
  QEB_REQUIRE(gout && boxes && gimg && H > 0 && W > 0 && oh > 0 && ow > 0 && n > 0, "crop_pad_scatter: bad args");
  const long long total = (long long)n * oh * ow;
  const int grid = qeb_grid(total, 256);
  crop_pad_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, H, W, boxes, n, oh, ow, gimg);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// OCR hand-off: fp32 images in [0,1] -> the uint8 pixels torchvision's ToPILImage produces for a float tensor
// (pic.mul(255).byte(): multiply in fp32, truncate), ocr_helper/tess_helper.py:20-24. 16 pixels per thread.
namespace {
__global__ void to_uint8_kernel(const float* __restrict__ x, long long n, uint8_t* __restrict__ out) {
  // grid-stride over groups of 16 pixels: the grid is capped at a few waves of the SMs, the loop covers any n
  const long long groups = (n + 15) / 16;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    const long long i = g * 16;
    if (i + 16 <= n) {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + i) + q);
        const float f[4] = {v.x, v.y, v.z, v.w};
        uint32_t pk = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) pk |= (uint32_t)(int)fminf(fmaxf(f[e] * 255.f, 0.f), 255.f) << (8 * e);
        w[q] = pk;
      }
      *reinterpret_cast<uint4*>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (long long j = i; j < n; ++j) out[j] = (uint8_t)(int)fminf(fmaxf(x[j] * 255.f, 0.f), 255.f);
    }
  }
}
}  // namespace

QEB_API int qeb_to_uint8(const float* x, long long n, unsigned char* out, void* stream) {
  QEB_REQUIRE(x && out && n >= 0, "to_uint8: bad arguments");
  QEB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "to_uint8: 16-byte aligned buffers");
  if (n == 0) return QEB_OK;
  ProfScope prof("to_uint8", (cudaStream_t)stream, 0.0, 5.0 * n);
  to_uint8_kernel<<<qeb_grid(qeb_cdiv(n, 16), 256), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}
