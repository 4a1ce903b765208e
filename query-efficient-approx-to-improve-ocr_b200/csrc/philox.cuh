// Philox4x32-10 + Box-Muller + the jitter arithmetic of AddGaussianNoice (transform_helper.py:33-45), shared by the
// standalone jitter kernel (image_ops.cu) and the fused jitter + conv1 kernel (nn_ops.cu): both draw the SAME noise for a
// given (seed, image index, pixel group), so the fused path materialises exactly the image the standalone kernel writes.
#pragma once
#include "common.cuh"

namespace qebrng {

struct Philox {
  static constexpr unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __device__ static uint4 rand4(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned int hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
      const unsigned int hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      k.x += W0;
      k.y += W1;
    }
    return c;
  }
};

__device__ __forceinline__ float2 box_muller(unsigned int a, unsigned int b) {
  // u1 in (0,1], u2 in [0,1)
  const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__device__ __forceinline__ float jitter1(float img, float noise, float coef) {
  // img - coef*noise as two roundings (torch: mul then sub), then clamp
  const float v = __fsub_rn(img, __fmul_rn(coef, noise));
  return fminf(fmaxf(v, 0.f), 1.f);
}

// the 4 normals N(mean, sigma) of pixel group g (4 consecutive pixels) of image im under key `seed`
__device__ __forceinline__ float4 noise4(unsigned int g, long long im, unsigned long long seed, float mean, float sg) {
  const uint4 r = Philox::rand4(make_uint4(g, (unsigned int)im, (unsigned int)(im >> 32), 0u),
                                make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
  const float2 n01 = box_muller(r.x, r.y), n23 = box_muller(r.z, r.w);
  return make_float4(__fadd_rn(mean, __fmul_rn(sg, n01.x)), __fadd_rn(mean, __fmul_rn(sg, n01.y)),
                     __fadd_rn(mean, __fmul_rn(sg, n23.x)), __fadd_rn(mean, __fmul_rn(sg, n23.y)));
}

}  // namespace qebrng
