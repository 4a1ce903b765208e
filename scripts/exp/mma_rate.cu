// Experiment: cycles per tcgen05.mma as a function of N, kind (tf32 / f16) and A source (shared memory / TMEM), M = 128.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../query-efficient-approx-to-improve-ocr_b200/csrc -I ../../include mma_rate.cu -o mma_rate
#include "tc_common.cuh"
#include <cstdio>
using namespace tc;
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc, int f16) {
  if (f16) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, int f16) {
  if (f16) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void k(int M, int N, int f16, int ts, int n_mma, int chains, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 16384 + 32768);
  uint32_t* slot = (uint32_t*)(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  asm volatile("fence.proxy.async;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (!f16) idesc |= (2u << 7) | (2u << 10);
    const uint64_t ad = smem_desc_kmajor_sw128(smem_u32(smem)), bd = smem_desc_kmajor_sw128(smem_u32(smem + 16384));
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t d = tb + 256 + (uint32_t)((i % chains) * N) % 256;
        if (ts) mma_ts(d, tb + (uint32_t)((i & 3) * 8), bd + 2 * (i & 3), idesc, i >= chains, f16);
        else mma_ss(d, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, i >= chains, f16);
      }
      const long long t1 = clock64();
      mma_commit(bar);
      mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}
int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  printf("%4s %4s %5s %3s %6s %7s | issue cyc/mma  total cyc/mma\n", "M", "N", "kind", "A", "n_mma", "chains");
  for (int M : {128, 64}) for (int f16 = 0; f16 < 2; ++f16) for (int ts = 0; ts < 2; ++ts) for (int N : {8, 16, 32, 64, 128, 256}) for (int chains : {1, 4}) {
    if (M == 128 && N < 16) continue;
    if (chains * N > 256) continue;
    if (chains == 4 && N > 32) continue;
    const int n = 64;
    k<<<1, 128, 60000>>>(M, N, f16, ts, n, chains, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M %d N %d f16 %d ts %d: %s\n", M, N, f16, ts, cudaGetErrorString(e)); return 1; }
    printf("%4d %4d %5s %3s %6d %7d | %8.1f %12.1f\n", M, N, f16 ? "f16" : "tf32", ts ? "T" : "S", n, chains, out[0] / (double)n, out[1] / (double)n);
  }
  return 0;
}
