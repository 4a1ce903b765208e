// Minibatch-subset selection keyed on CER: TopKCER and rangeCER, segmented (one warp per minibatch).
//
// Replaces the CPU paths of
//   TopKCERSampler.query   selection_utils.py:144-151   argsort(cers, descending)[:k]
//   CerRangeSampler.query  selection_utils.py:107-135   k points (max-min)*rand+min, sequential nearest
//                                                       neighbour without replacement (sentinel 100)
// Values are fp32 (the reference builds torch.tensor(list_of_python_floats) -> float32).
// Tie-break contract: equal CERs are ordered lowest-index-first (= argsort(stable=True));
// argmin returns the first minimal index (torch.argmin). The range points are computed as a separate fp32
// multiply and add (__fmul_rn/__fadd_rn, no FMA contraction), like the two ATen ops.
#include "common.cuh"

namespace {

// descending order key: larger float first; NaN sorts as largest (torch.sort semantics); stable on index.
__device__ __forceinline__ unsigned int desc_key(float v) {
  if (v != v) return 0xffffffffu;
  unsigned int u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (u == 0x7fffffffu) u = 0x80000000u;  // -0.0 == +0.0
  return u;
}

// one warp per segment; rank of element i = #{j : key_j > key_i or (key_j == key_i and j < i)}.
// The segment's keys are staged in shared memory (coalesced read) when they fit.
constexpr int kTopkCap = 1024;  // keys staged per warp

__global__ void __launch_bounds__(128) topk_segmented_kernel(const float* __restrict__ vals,
                                                            const int* __restrict__ seg_off,
                                                            const int* __restrict__ seg_k,
                                                            const int* __restrict__ out_off, int n_seg,
                                                            long long* __restrict__ out_idx) {
  __shared__ unsigned int skeys[4][kTopkCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.x * 4 + warp;
  if (seg >= n_seg) return;
  const int s0 = seg_off[seg], n = seg_off[seg + 1] - s0;
  const int k = min(seg_k[seg], n);
  long long* out = out_idx + out_off[seg];
  const bool staged = n <= kTopkCap;
  if (staged) {
    for (int i = lane; i < n; i += 32) skeys[warp][i] = desc_key(vals[s0 + i]);
    __syncwarp();
  }
  for (int i = lane; i < n; i += 32) {
    const unsigned int ki = staged ? skeys[warp][i] : desc_key(vals[s0 + i]);
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned int kj = staged ? skeys[warp][j] : desc_key(vals[s0 + j]);
      rank += (kj > ki) || (kj == ki && j < i);
    }
    if (rank < k) out[rank] = i;
  }
}

// one warp per segment. points[j] = (max-min)*rand[j] + min; for each point in order: first index of the minimum of
// |point - copy|, then copy[index] = 100.
__global__ void range_segmented_kernel(const float* __restrict__ vals, const int* __restrict__ seg_off,
                                       const int* __restrict__ seg_k, const int* __restrict__ out_off,
                                       const float* __restrict__ rands, int n_seg, float* __restrict__ work,
                                       long long* __restrict__ out_idx, float* __restrict__ points_out) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.x * warps + warp;
  if (seg >= n_seg) return;
  const int s0 = seg_off[seg], n = seg_off[seg + 1] - s0;
  const int k = seg_k[seg];
  const int o0 = out_off[seg];
  if (n == 0) return;
  float* copy = work + s0;
  // max / min: torch.max/min propagate NaN; CERs are never NaN, plain fmax/fmin is exact for ordered floats
  float mx = -INFINITY, mn = INFINITY;
  for (int i = lane; i < n; i += 32) {
    const float v = vals[s0 + i];
    copy[i] = v;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
    mn = fminf(mn, __shfl_xor_sync(FULL_MASK, mn, o));
  }
  const float span = __fsub_rn(mx, mn);
  __syncwarp();
  for (int j = 0; j < k; ++j) {
    const float point = __fadd_rn(__fmul_rn(span, rands[o0 + j]), mn);
    if (points_out && lane == 0) points_out[o0 + j] = point;
    float best = INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      const float d = fabsf(__fsub_rn(point, copy[i]));
      if (bi == 0x7fffffff || d < best) { best = d; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL_MASK, best, o);
      const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      out_idx[o0 + j] = bi;
      copy[bi] = 100.f;
    }
    __syncwarp();
  }
}

// ---- one LARGE segment (dataset-wide top-k: pruning/methods.py:5-8, the per-rank shard reduction of dist.global_topk) --------
// The warp kernel above ranks a segment with n^2 comparisons - right for minibatch-sized segments (n <= 1024), hopeless for
// 10^5..10^7 CERs. Here every element becomes ONE 64-bit key (~desc_key(value) << 32 | index): ascending key order is
// "value descending, index ascending", keys are distinct, so any comparison sort gives exactly the stable order. The keys
// are sorted with a bitonic network: chunks of 2048 keys in shared memory (all sub-stages with stride < 2048), strides
// >= 2048 as global compare-exchange passes. O(n log^2 n) compare-exchanges, 8 MB per pass at n = 1 M: HBM-bound passes.
constexpr int kSortChunk = 2048;   // keys per block in the shared-memory kernels (512 threads x 4)

__global__ void topk_keys_kernel(const float* __restrict__ vals, long long n, long long n_pad, unsigned long long* __restrict__ keys) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (long long)gridDim.x * blockDim.x)
    keys[i] = i < n ? ((unsigned long long)(~desc_key(vals[i])) << 32) | (unsigned long long)i : ~0ull;
}

__device__ __forceinline__ void cmp_swap(unsigned long long& a, unsigned long long& b, bool up) {
  if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

// sub-stages j = j_hi, j_hi/2, ..., 1 of stage k on one 2048-key chunk held in shared memory; full_sort: all stages
// k = 2 .. 2048 (the initial chunk sort)
__global__ void __launch_bounds__(512) bitonic_shared_kernel(unsigned long long* __restrict__ keys, long long k, int full_sort) {
  __shared__ unsigned long long s[kSortChunk];
  const long long base = (long long)blockIdx.x * kSortChunk;
  for (int i = threadIdx.x; i < kSortChunk; i += 512) s[i] = keys[base + i];
  __syncthreads();
  for (long long kk = full_sort ? 2 : k; kk <= k; kk <<= 1) {
    for (int j = (int)(kk < kSortChunk ? kk >> 1 : kSortChunk >> 1); j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < kSortChunk / 2; t += 512) {
        const int lo = 2 * t - (t & (j - 1));        // index with bit j clear
        const bool up = ((base + lo) & kk) == 0;
        cmp_swap(s[lo], s[lo + j], up);
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < kSortChunk; i += 512) keys[base + i] = s[i];
}

// one global compare-exchange pass: stage k, stride j (>= kSortChunk)
__global__ void bitonic_global_kernel(unsigned long long* __restrict__ keys, long long n_pad, long long k, long long j) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_pad / 2; t += (long long)gridDim.x * blockDim.x) {
    const long long lo = 2 * t - (t & (j - 1));
    unsigned long long a = keys[lo], b = keys[lo + j];
    const bool up = (lo & k) == 0;
    if ((a > b) == up) { keys[lo] = b; keys[lo + j] = a; }
  }
}

__global__ void topk_emit_kernel(const unsigned long long* __restrict__ keys, long long k, long long* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (long long)gridDim.x * blockDim.x)
    out[i] = (long long)(keys[i] & 0xffffffffull);
}

long long sort_pad(long long n) {
  long long p = kSortChunk;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace

// Stable top-k of ONE segment of any size: out_idx[r] = index of the r-th largest value, equal values lowest index first
// (= argsort(descending, stable)[:k]); k is clamped to n. work: qeb_cer_topk_global_workspace_bytes(n) bytes.
QEB_API size_t qeb_cer_topk_global_workspace_bytes(long long n) { return n > 0 ? (size_t)sort_pad(n) * 8 : 0; }

QEB_API int qeb_cer_topk_global(const float* vals, long long n, long long k, void* work, long long* out_idx, void* stream) {
  if (n <= 0 || k <= 0) return QEB_OK;
  QEB_REQUIRE(vals && work && out_idx, "cer_topk_global: null pointer");
  QEB_REQUIRE(n < (1ll << 32), "cer_topk_global: n=%lld does not fit the 32-bit index field", n);
  QEB_REQUIRE(((uintptr_t)work & 7) == 0, "cer_topk_global: workspace must be 8-byte aligned");
  if (k > n) k = n;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* keys = static_cast<unsigned long long*>(work);
  const long long n_pad = sort_pad(n);
  ProfScope prof("topk_global", st, 0.0, 4.0 * n + 8.0 * k);
  topk_keys_kernel<<<qeb_grid(n_pad, 256), 256, 0, st>>>(vals, n, n_pad, keys);
  QEB_LAUNCH_CHECK();
  const int chunks = (int)(n_pad / kSortChunk);
  bitonic_shared_kernel<<<chunks, 512, 0, st>>>(keys, kSortChunk, 1);
  QEB_LAUNCH_CHECK();
  int launches = 2;
  for (long long kk = 2 * kSortChunk; kk <= n_pad; kk <<= 1) {
    for (long long j = kk >> 1; j >= kSortChunk; j >>= 1) {
      bitonic_global_kernel<<<qeb_grid(n_pad / 2, 256), 256, 0, st>>>(keys, n_pad, kk, j);
      QEB_LAUNCH_CHECK();
      ++launches;
    }
    bitonic_shared_kernel<<<chunks, 512, 0, st>>>(keys, kk, 0);
    QEB_LAUNCH_CHECK();
    ++launches;
  }
  topk_emit_kernel<<<qeb_grid(k, 256), 256, 0, st>>>(keys, k, out_idx);
  QEB_LAUNCH_CHECK();
  qeb_count_launch(launches + 1);
  return QEB_OK;
}

// vals: concatenated fp32 CERs of all segments; seg_off (n_seg+1); seg_k (n_seg) requested picks per segment;
// out_off (n_seg) exclusive prefix sum of min(k, n); out_idx: int64 indices local to the segment.
QEB_API int qeb_cer_topk_segmented(const float* vals, const int* seg_off, const int* seg_k, const int* out_off,
                                   int n_seg, long long* out_idx, void* stream) {
  if (n_seg == 0) return QEB_OK;
  QEB_REQUIRE(vals && seg_off && seg_k && out_off && out_idx && n_seg > 0, "cer_topk_segmented: bad args");
  ProfScope prof("topk_segmented", (cudaStream_t)stream);
  topk_segmented_kernel<<<qeb_cdiv(n_seg, 4), 128, 0, (cudaStream_t)stream>>>(vals, seg_off, seg_k, out_off, n_seg, out_idx);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

// rands: concatenated torch.rand(k) draws per segment (host RNG, so the reference's random stream is kept);
// work: scratch of the same size as vals; out_off here is the exclusive prefix sum of k (k may exceed n).
QEB_API int qeb_cer_range_segmented(const float* vals, const int* seg_off, const int* seg_k, const int* out_off,
                                    const float* rands, int n_seg, float* work, long long* out_idx,
                                    float* points_out, void* stream) {
  if (n_seg == 0) return QEB_OK;
  QEB_REQUIRE(vals && seg_off && seg_k && out_off && rands && work && out_idx && n_seg > 0, "cer_range_segmented: bad args");
  ProfScope prof("range_segmented", (cudaStream_t)stream);
  range_segmented_kernel<<<qeb_cdiv(n_seg, 4), 128, 0, (cudaStream_t)stream>>>(vals, seg_off, seg_k, out_off, rands, n_seg,
                                                                               work, out_idx, points_out);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}
