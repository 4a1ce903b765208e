// Experiment: does a K-major SWIZZLE_128B UMMA descriptor whose start address is a multiple of 128 B but not of 1024 B
// (a tile shifted by r rows) read the rows TMA wrote, with base_offset = (addr >> 7) & 7 or with base_offset = 0?
#include "../../query-efficient-approx-to-improve-ocr_b200/csrc/tc_common.cuh"
#include <vector>
#include <cstdio>
#include <cstdlib>
using namespace tc;
void qeb_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }

static int make_tmap2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t bc, uint32_t br) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  cuuint64_t d[2] = {cols, rows}, s[1] = {cols * 4}; cuuint32_t b[2] = {bc, br}, e[2] = {1, 1};
  return ((EncodeFn)fn)(out, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

constexpr int kRowsA = 144;  // rows loaded (>= 128 + 8)
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int shift,
                                           int use_base_offset, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                       // 144 x 128 B
  uint8_t* sb = smem + 20 * 1024;           // 32 x 128 B
  uint64_t* bar = (uint64_t*)(smem + 28 * 1024);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, kRowsA * 128 + 32 * 128);
    tma_load_2d(sa, &ta, bar, 0, 0);
    tma_load_2d(sb, &tb, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    uint64_t adesc = smem_desc_kmajor_sw128(a_addr);
    if (use_base_offset) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = smem_desc_kmajor_sw128(smem_u32(sb));
    constexpr uint32_t idesc = instr_desc_tf32(128, 32, 0, 0);
    for (int kk = 0; kk < 4; ++kk) mma_tf32_ss(tm, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk != 0);
    mma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  float v[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = v[j];
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}

int main() {
  std::vector<float> A(kRowsA * 32), B(32 * 32, 0.f);
  for (int r = 0; r < kRowsA; ++r) for (int c = 0; c < 32; ++c) A[r * 32 + c] = (float)(r * 32 + c);   // exactly representable in tf32? up to 4607 < 2^13: 11 bits mantissa needed -> use smaller
  for (int r = 0; r < kRowsA; ++r) for (int c = 0; c < 32; ++c) A[r * 32 + c] = (float)((r * 7 + c * 37) % 509);  // < 2^9: exact in tf32
  for (int i = 0; i < 32; ++i) B[i * 32 + i] = 1.f;  // D = A (identity)
  float *dA, *dB, *dO; cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  if (make_tmap2d(&ta, dA, 32, kRowsA, 32, kRowsA) || make_tmap2d(&tb, dB, 32, 32, 32, 32)) { printf("tmap fail\n"); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
  std::vector<float> O(128 * 32);
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int shift = 0; shift < 10; ++shift) {
      k<<<1, 128, 32 * 1024>>>(ta, tb, shift, ubo, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d base_offset %d: CUDA error %s\n", shift, ubo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0; int first_bad = -1;
      for (int r = 0; r < 128; ++r) for (int c = 0; c < 32; ++c) {
        const float want = A[(r + shift) * 32 + c];
        if (fabsf(O[r * 32 + c] - want) > 1e-3f * fabsf(want) + 1e-3f) { if (first_bad < 0) first_bad = r * 32 + c; ++bad; }
      }
      printf("shift %d base_offset_field %d: %s (%d wrong)", shift, ubo, bad ? "MISMATCH" : "ok", bad);
      if (bad) printf("  e.g. out[%d][%d] = %.4f want %.4f", first_bad / 32, first_bad % 32, O[first_bad], A[(first_bad / 32 + shift) * 32 + first_bad % 32]);
      printf("\n");
    }
  return 0;
}
