// Shared helpers for the qeb sm_100a kernels: error reporting, launch checks, warp primitives.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

#define QEB_OK 0
#define QEB_ERR_INVALID -1   // bad shape / alignment / null pointer, nothing launched
#define QEB_ERR_CUDA -2      // CUDA runtime / driver error at launch
#define QEB_ERR_UNSUPPORTED -3

#define QEB_API extern "C" __attribute__((visibility("default")))

// thread-local last error message (qeb_last_error)
void qeb_set_error(const char* fmt, ...);

#define QEB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      qeb_set_error(__VA_ARGS__);         \
      return QEB_ERR_INVALID;             \
    }                                     \
  } while (0)

#define QEB_CUDA(expr)                                                             \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      qeb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return QEB_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

#define QEB_LAUNCH_CHECK() QEB_CUDA(cudaGetLastError())

static inline int qeb_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
constexpr int kNumSMs = 148;  // B200
// grid for a grid-stride elementwise kernel: enough blocks to cover `total`, capped at a few waves of the 148 SMs
static inline int qeb_grid(long long total, int threads, int blocks_per_sm = 8) {
  long long g = (total + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// launch-counter (bench.py reports gpu_launches from it)
void qeb_count_launch(int n = 1);
long long qeb_launch_count_now();

// per-launch profiling scope (abi.cu): no-op unless qeb_prof_enable(1)
int qeb_prof_on();
int qeb_prof_begin(const char* tag, cudaStream_t st, double flops, double bytes);
void qeb_prof_end(int idx, cudaStream_t st);
// Ablation aid (abi.cu): QEB_DBG_SKIP="tag1,tag2" makes every kernel launched through qeb_launch inside a ProfScope whose
// tag starts with one of the listed prefixes a no-op. Results are garbage; the step time that disappears is that kernel
// family's true share of the critical path inside the CUDA-graph replay (which per-launch events cannot show). Unset: free.
bool qeb_dbg_skip_active();
bool qeb_dbg_skip_tag(const char* tag);
int& qeb_skip_flag();
struct ProfScope {
  int idx = -1;
  int prev_skip = 0;
  cudaStream_t st;
  ProfScope(const char* tag, cudaStream_t s, double flops = 0.0, double bytes = 0.0) : st(s) {
    if (qeb_prof_on()) idx = qeb_prof_begin(tag, s, flops, bytes);
    if (qeb_dbg_skip_active()) {
      prev_skip = qeb_skip_flag();
      if (qeb_dbg_skip_tag(tag)) qeb_skip_flag() = 1;
    }
  }
  ~ProfScope() {
    if (idx >= 0) qeb_prof_end(idx, st);
    if (qeb_dbg_skip_active()) qeb_skip_flag() = prev_skip;
  }
};

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------------
// A training step is ~230 short kernels in a row; with plain stream order every kernel pays launch latency + the ramp of
// its first wave after the previous grid has drained. Kernels of the step are launched with the programmatic-stream-
// serialization attribute (also captured into CUDA graphs as programmatic edges): the next grid may be scheduled while the
// previous one is still running, and blocks in qeb_pdl_sync() - the FIRST statement of every kernel launched this way -
// until the previous grid has completed and its memory is visible. QEB_PDL=0 launches everything with full serialization.
bool qeb_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void qeb_pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// the same launch as a grid of thread-block clusters of `cluster_x` CTAs along x (gridDim.x a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t qeb_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = qeb_pdl_enabled() ? 2 : 1;
  if (qeb_dbg_skip_active() && qeb_skip_flag()) return cudaSuccess;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t qeb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = qeb_pdl_enabled() ? 1 : 0;
  if (qeb_dbg_skip_active() && qeb_skip_flag()) return cudaSuccess;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#ifdef __CUDACC__
constexpr unsigned FULL_MASK = 0xffffffffu;

// ---- tf32 operands, rounded where they are PRODUCED ---------------------------------------------------------------------
// tcgen05.mma.kind::tf32 reads fp32 words from shared memory and TRUNCATES them to tf32 (10 mantissa bits): a bias of half an
// ulp per operand that the backward pass of the random-init networks amplifies (weight gradients 1-2e-2 off the fp32 oracle,
// profiles/r1_notes.md 7 / 21). TMA moves the bytes untouched, so the only place to round is the kernel that WRITES a tensor
// which a tf32 contraction will read: it stores round-to-nearest(tf32) - an fp32 word whose low 13 mantissa bits are zero, so
// the tensor core's truncation is exact. Same 11-bit significand as the fp16 operand copies of the forward pass, full fp32
// range, no extra bytes. Applies to: activations kept for the weight gradients, every gradient tensor that feeds a dgrad /
// wgrad contraction, the packed dgrad weights. QEB_TF32_RN=0 at compile time restores plain stores.
#ifndef QEB_TF32_RN
#define QEB_TF32_RN 1
#endif
__device__ __forceinline__ float qeb_tf32r(float x) {
#if QEB_TF32_RN
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
#else
  return x;
#endif
}
__device__ __forceinline__ float4 qeb_tf32r4(float4 v) { return make_float4(qeb_tf32r(v.x), qeb_tf32r(v.y), qeb_tf32r(v.z), qeb_tf32r(v.w)); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
#endif
