// Batched Levenshtein distance + CER, greedy CTC decode. Integer paths: bit-exact against the oracle.
//
// Replaces, for a whole batch in one launch:
//   Levenshtein.distance(labels[i], preds[i]) / max(1, len(labels[i]))      utils.py:103-109 (compare_labels)
//   the per-timestep argmax / collapse / drop-blank loop                      utils.py:74-92  (pred_to_string)
// python-Levenshtein 0.12.0 (requirements.txt:70) is not vendored; its distance() is the unit-cost
// insert/delete/substitute edit distance over code points, restated in oracle/levenshtein.c.
//
// Layout: strings in CSR form: one flat array of symbols (uint8 char-set indices or int32 code points) and
// n+1 int32 offsets per side. A block of 256 pairs first stages its contiguous CSR span of both sides in
// shared memory with coalesced loads (spans are adjacent in memory because pairs are consecutive), then each
// thread runs the single-row Wagner-Fischer recurrence for its pair out of shared memory.
#include "common.cuh"

namespace {

constexpr int kLevThreads = 256;
constexpr int kLevStage = 12 * 1024;  // symbols staged per side per block (fallback: read global)
constexpr int kLevRow = 128;          // per-thread DP row (shorter string <= 128); longer pairs -> block kernel

template <typename Sym>
__global__ void __launch_bounds__(kLevThreads) lev_small_kernel(const Sym* __restrict__ a, const int* __restrict__ aoff,
                                                               const int* __restrict__ alen,
                                                               const Sym* __restrict__ b, const int* __restrict__ boff,
                                                               const int* __restrict__ blen, int n, int* __restrict__ dist, double* __restrict__ cer,
                                                               int* __restrict__ n_long) {
  extern __shared__ unsigned char smraw[];
  Sym* sa = reinterpret_cast<Sym*>(smraw);
  Sym* sb = sa + kLevStage;
  const int base = blockIdx.x * kLevThreads;
  const int cnt = min(kLevThreads, n - base);
  const int a0 = aoff[base], a1 = alen ? aoff[base + cnt - 1] + alen[base + cnt - 1] : aoff[base + cnt];
  const int b0 = boff[base], b1 = blen ? boff[base + cnt - 1] + blen[base + cnt - 1] : boff[base + cnt];
  const bool stage_a = (a1 - a0) <= kLevStage, stage_b = (b1 - b0) <= kLevStage;
  if (stage_a) for (int i = threadIdx.x; i < a1 - a0; i += kLevThreads) sa[i] = a[a0 + i];
  if (stage_b) for (int i = threadIdx.x; i < b1 - b0; i += kLevThreads) sb[i] = b[b0 + i];
  __syncthreads();
  const int p = base + threadIdx.x;
  if (p >= n) return;
  const int as = aoff[p], la = alen ? alen[p] : aoff[p + 1] - as;
  const int bs = boff[p], lb = blen ? blen[p] : boff[p + 1] - bs;
  const Sym* pa = stage_a ? sa + (as - a0) : a + as;
  const Sym* pb = stage_b ? sb + (bs - b0) : b + bs;
  // x = shorter string (DP row), y = longer (outer loop); the distance is symmetric
  const Sym* x = pa; int lx = la; const Sym* y = pb; int ly = lb;
  if (lx > ly) { x = pb; lx = lb; y = pa; ly = la; }
  int d;
  if (lx == 0) {
    d = ly;
  } else if (lx > kLevRow) {
    atomicAdd(n_long, 1);
    d = -1;  // filled in by lev_long_kernel
  } else {
    unsigned short row[kLevRow + 1];
    for (int j = 0; j <= lx; ++j) row[j] = (unsigned short)j;
    for (int i = 1; i <= ly; ++i) {
      const Sym c = y[i - 1];
      int diag = row[0];
      int left = i;
      row[0] = (unsigned short)i;
      for (int j = 1; j <= lx; ++j) {
        const int up = row[j];
        int v = diag + (x[j - 1] != c);
        v = min(v, min(up, left) + 1);
        row[j] = (unsigned short)v;
        diag = up;
        left = v;
      }
    }
    d = row[lx];
  }
  dist[p] = d;
  if (cer && d >= 0) cer[p] = (double)d / (double)(la > 1 ? la : 1);
}

// rare path: both strings longer than kLevRow. One block per pair, DP row in shared memory, computed by
// anti-diagonal-free single thread (these are far off the measured path: OCR labels <= 100 characters).
template <typename Sym>
__global__ void lev_long_kernel(const Sym* __restrict__ a, const int* __restrict__ aoff, const int* __restrict__ alen,
                                const Sym* __restrict__ b, const int* __restrict__ boff, const int* __restrict__ blen, int n, int* __restrict__ dist, double* __restrict__ cer,
                                int row_cap) {
  extern __shared__ int lrow[];
  for (int p = blockIdx.x; p < n; p += gridDim.x) {
    if (dist[p] >= 0) continue;
    if (threadIdx.x != 0) continue;
    const int as = aoff[p], la = alen ? alen[p] : aoff[p + 1] - as;
    const int bs = boff[p], lb = blen ? blen[p] : boff[p + 1] - bs;
    const Sym* x = a + as; int lx = la; const Sym* y = b + bs; int ly = lb;
    if (lx > ly) { x = b + bs; lx = lb; y = a + as; ly = la; }
    if (lx > row_cap) { dist[p] = -2; continue; }
    for (int j = 0; j <= lx; ++j) lrow[j] = j;
    for (int i = 1; i <= ly; ++i) {
      const Sym c = y[i - 1];
      int diag = lrow[0], left = i;
      lrow[0] = i;
      for (int j = 1; j <= lx; ++j) {
        const int up = lrow[j];
        int v = diag + (x[j - 1] != c);
        v = min(v, min(up, left) + 1);
        lrow[j] = v;
        diag = up;
        left = v;
      }
    }
    dist[p] = lrow[lx];
    if (cer) cer[p] = (double)lrow[lx] / (double)(la > 1 ? la : 1);
  }
}

template <typename Sym>
int lev_launch(const void* a, const int* aoff, const int* alen, const void* b, const int* boff, const int* blen, int n,
               int max_len, int* dist,
               double* cer, int* scratch, cudaStream_t st) {
  const size_t smem = 2 * (size_t)kLevStage * sizeof(Sym);
  static bool attr_set = false;
  if (!attr_set) {
    QEB_CUDA(cudaFuncSetAttribute(lev_small_kernel<Sym>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  lev_small_kernel<Sym><<<qeb_cdiv(n, kLevThreads), kLevThreads, smem, st>>>(
      (const Sym*)a, aoff, alen, (const Sym*)b, boff, blen, n, dist, cer, scratch);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  if (max_len > kLevRow) {
    const int cap = max_len;
    const size_t lsmem = (size_t)(cap + 1) * sizeof(int);
    QEB_REQUIRE(lsmem <= 200 * 1024, "levenshtein: strings longer than %d symbols unsupported", 200 * 256 - 1);
    QEB_CUDA(cudaFuncSetAttribute(lev_long_kernel<Sym>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
    lev_long_kernel<Sym><<<min(n, 148 * 4), 32, lsmem, st>>>((const Sym*)a, aoff, alen, (const Sym*)b, boff, blen, n, dist,
                                                             cer, cap);
    QEB_LAUNCH_CHECK();
    qeb_count_launch();
  }
  return QEB_OK;
}

// ---------------------------------------------------------------------------------------------
// greedy decode: one warp per sample; per timestep a coalesced read of the V scores and a warp arg-max with
// first-index tie-break (torch.argmax), then collapse repeats / drop blank.
// torch.argmax ordering: NaN counts as the largest value, the first maximal index wins
__device__ __forceinline__ bool score_gt(float a, float b) { return (a > b) || ((a != a) && !(b != b)); }

__global__ void greedy_decode_kernel(const float* __restrict__ scores, long long st_t, long long st_b, int T, int B,
                                     int V, int blank, int* __restrict__ out, int* __restrict__ out_len,
                                     int* __restrict__ raw_path) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int prev = -1, n = 0;
  for (int t = 0; t < T; ++t) {
    const float* r = scores + t * st_t + b * st_b;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
      const float v = r[c];
      if (bi == 0x7fffffff || score_gt(v, best)) { best = v; bi = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(FULL_MASK, best, o);
      const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
      if (score_gt(ov, best) || (!score_gt(best, ov) && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      if (raw_path) raw_path[(long long)b * T + t] = bi;
      if (bi != blank && bi != prev) out[(long long)b * T + n++] = bi;
      prev = bi;
    }
  }
  if (lane == 0) {
    out_len[b] = n;
    for (int i = n; i < T; ++i) out[(long long)b * T + i] = -1;
  }
}

// the collapse half of pred_to_string (utils.py:84-89) on a per-frame arg-max path that the CRNN head's epilogue already
// produced (conv_tc.cu lsm_epilogue): drop blanks, collapse repeats. One thread per sample; frame t of sample b at
// path[t * st_t + b * st_b] (coalesced over b for the (T,B) layout the head writes).
__global__ void greedy_collapse_kernel(const int* __restrict__ path, long long st_t, long long st_b, int T, int B, int blank,
                                       int* __restrict__ out, int* __restrict__ out_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int prev = -1, n = 0;
  for (int t = 0; t < T; ++t) {
    const int bi = path[t * st_t + b * st_b];
    if (bi != blank && bi != prev) out[(long long)b * T + n++] = bi;
    prev = bi;
  }
  out_len[b] = n;
  for (int i = n; i < T; ++i) out[(long long)b * T + i] = -1;
}

}  // namespace

// symbols: sym_bytes = 1 (uint8 char-set indices) or 4 (int32 code points). a = labels (CER denominator),
// b = predictions. Each side is either CSR (len NULL, offsets have n+1 entries) or padded rows (offsets = row
// starts, len = symbols used per row). dist (n) int32; cer (n) fp64 or NULL; scratch: 1 int (zeroed by callee).
QEB_API int qeb_levenshtein_batch(const void* a_syms, const int* a_off, const int* a_len, const void* b_syms,
                                  const int* b_off, const int* b_len, int n, int sym_bytes, int max_len, int* dist, double* cer, int* scratch, void* stream) {
  QEB_REQUIRE(n >= 0, "levenshtein: n < 0");
  if (n == 0) return QEB_OK;
  QEB_REQUIRE(a_off && b_off && dist && scratch, "levenshtein: null pointer");
  QEB_REQUIRE(sym_bytes == 1 || sym_bytes == 4, "levenshtein: sym_bytes must be 1 or 4, got %d", sym_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  QEB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int), st));
  ProfScope prof("levenshtein", st, 0.0, (2.0 * 4 + 4 + (cer ? 8 : 0)) * n);   // offsets + results; callers add the symbol bytes
  if (sym_bytes == 1) return lev_launch<unsigned char>(a_syms, a_off, a_len, b_syms, b_off, b_len, n, max_len, dist, cer, scratch, st);
  return lev_launch<int>(a_syms, a_off, a_len, b_syms, b_off, b_len, n, max_len, dist, cer, scratch, st);
}

// scores (T,B,V) -> out (B,T) int32 class indices (padded with -1), out_len (B); raw_path (B,T) optional.
QEB_API int qeb_greedy_decode(const float* scores, long long st_t, long long st_b, int T, int B, int V, int blank,
                              int* out, int* out_len, int* raw_path, void* stream) {
  QEB_REQUIRE(scores && out && out_len, "greedy_decode: null pointer");
  QEB_REQUIRE(T > 0 && B > 0 && V > 0, "greedy_decode: bad sizes");
  ProfScope prof("greedy_decode", (cudaStream_t)stream, 0.0, 4.0 * T * B * V + 4.0 * B * (T + 1));
  greedy_decode_kernel<<<qeb_cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(scores, st_t, st_b, T, B, V, blank, out,
                                                                         out_len, raw_path);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

// path: per-frame arg-max classes, frame t of sample b at path[t*st_t + b*st_b] -> the same out / out_len as qeb_greedy_decode.
QEB_API int qeb_greedy_collapse(const int* path, long long st_t, long long st_b, int T, int B, int blank, int* out, int* out_len,
                                void* stream) {
  QEB_REQUIRE(path && out && out_len, "greedy_collapse: null pointer");
  QEB_REQUIRE(T > 0 && B > 0, "greedy_collapse: bad sizes");
  ProfScope prof("greedy_decode", (cudaStream_t)stream, 0.0, 4.0 * T * B + 4.0 * B * (T + 1));
  greedy_collapse_kernel<<<qeb_cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(path, st_t, st_b, T, B, blank, out, out_len);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}
