// Layout conversions around the tensor-core kernels: weight packing (torch OIHW -> K-major rows) and
// NCHW <-> NHWC activation transposes at the module boundary. HBM-bound copy kernels.
#include "common.cuh"

namespace {

// src: (A, B, KH, KW) contiguous (torch Conv2d weight: A=Cout, B=Cin; ConvTranspose2d: A=Cin, B=Cout)
//  mode 0: dst[a][kh*KW+kw][b]                       conv fprop B-operand; ConvTranspose dgrad B-operand
//  mode 1: dst[b][(KH-1-kh)*KW+(KW-1-kw)][a]         conv dgrad B-operand (flipped taps, channels transposed)
//  mode 2: dst[(kh*KW+kw)*B + b][a]                  ConvTranspose fprop B-operand (N index = (dh,dw,co))
__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int B, int KH, int KW,
                                   int mode) {
  const long long total = (long long)A * B * KH * KW;
  const int taps = KH * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes dst linearly so that writes are coalesced
    int a, b, tap;
    if (mode == 0) {
      b = (int)(i % B); tap = (int)((i / B) % taps); a = (int)(i / ((long long)B * taps));
    } else if (mode == 1) {
      a = (int)(i % A);
      const int ft = (int)((i / A) % taps);
      b = (int)(i / ((long long)A * taps));
      tap = taps - 1 - ft;  // (KH-1-kh)*KW + (KW-1-kw) == taps-1-(kh*KW+kw)
    } else {
      a = (int)(i % A); b = (int)((i / A) % B); tap = (int)(i / ((long long)A * B));
    }
    dst[i] = src[((long long)a * B + b) * taps + tap];
  }
}

// (N, C, H, W) -> (N, H, W, C) with destination channel stride (tile transpose through shared memory)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int dst_cstride) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (long long)n * C * HW;
  float* d = dst + (long long)n * HW * dst_cstride;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? s[(long long)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    if (p < HW && c < C) d[(long long)p * dst_cstride + c] = tile[threadIdx.x][j];
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW, int src_cstride) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (long long)n * HW * src_cstride;
  float* d = dst + (long long)n * C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? s[(long long)p * src_cstride + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < HW) d[(long long)c * HW + p] = tile[threadIdx.x][j];
  }
}

}  // namespace

QEB_API int qeb_pack_weight(const float* src, float* dst, int A, int B, int KH, int KW, int mode, void* stream) {
  QEB_REQUIRE(src && dst && A > 0 && B > 0 && KH > 0 && KW > 0 && mode >= 0 && mode <= 2, "pack_weight: bad args");
  const long long total = (long long)A * B * KH * KW;
  pack_weight_kernel<<<qeb_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, A, B, KH, KW, mode);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

QEB_API int qeb_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int dst_cstride, void* stream) {
  QEB_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && dst_cstride >= C, "nchw_to_nhwc: bad args");
  const int HW = H * W;
  nchw_to_nhwc_kernel<<<dim3(qeb_cdiv(HW, 32), qeb_cdiv(C, 32), N), dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, C, HW, dst_cstride);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

QEB_API int qeb_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, int src_cstride, void* stream) {
  QEB_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && src_cstride >= C, "nhwc_to_nchw: bad args");
  const int HW = H * W;
  nhwc_to_nchw_kernel<<<dim3(qeb_cdiv(HW, 32), qeb_cdiv(C, 32), N), dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, C, HW, src_cstride);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}
