// CTC loss (alpha/beta in log space, one warp per sequence) and the CRNN head's log-softmax.
//
// Replaces torch.nn.CTCLoss as the reference calls it (blank=0, reduction mean/none, zero_infinity=False):
//   ctor  train_nn_patch.py:143-144, train_nn_area.py:146-148, train_crnn.py:130-131
//   calls train_nn_patch.py:178,294, train_nn_area.py:174,265, tracking_utils.py:68,72, train_crnn.py:160
// and fn.log_softmax(self.linear(x), 2) in models/model_crnn.py:20.
//
// Semantics follow ATen's native CPU CTC (the oracle): the value returned as "gradient w.r.t. log_probs" is
//   (exp(lp) - exp(logsumexp_{s: l'_s = c}(alpha+beta) + nll - lp)) * grad_out
// i.e. already the gradient at the logits; an infeasible target gives nll = +inf and NaN
// gradients for that sample unless zero_infinity is set (the reference scrubs them with CRNN.backward_hook).
//
// Layout: log_probs (T, B, V) fp32 with arbitrary t/b strides (class stride 1). States of the extended
// label l' (2L+1) are interleaved over the 32 lanes: state s lives in lane s%32, register slot s/32, so the
// s-1 / s-2 neighbours come from warp shuffles. Each timestep's V log-probs are staged once in shared memory
// with a coalesced read and gathered from there.
#include "common.cuh"
#include <type_traits>

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxSPL = 8;  // states per lane: up to 256 extended states => target length <= 127

__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(a, fmaxf(b, c));
  if (m == neg_inf()) m = 0.f;
  return logf(expf(a - m) + expf(b - m) + expf(c - m)) + m;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct CtcArgs {
  const float* lp;
  long long st_t, st_b;     // strides of log_probs in elements
  const int* batch_index;   // optional: sample b reads/writes column batch_index[b] (weighted_ctc_loss subsets)
  const int* targets;       // concatenated targets
  const int* tgt_offsets;   // exclusive prefix sum of target_lengths, B entries
  const int* input_lengths;
  const int* target_lengths;
  int B, T, V, blank, S;    // S = allocated state stride (2*Lmax+1)
  float* log_alpha;         // (B, T, S)
  float* nll;               // (B)
};

template <int SPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ctc_alpha_kernel(CtcArgs a) {
  qeb_pdl_sync();
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerBlock + warp;
  if (b >= a.B) return;
  float* row = smem + warp * a.V;
  const int L = a.target_lengths[b];
  const int Tb = a.input_lengths[b];
  const int off = a.tgt_offsets[b];
  const int S = 2 * L + 1;
  const int col = a.batch_index ? a.batch_index[b] : b;
  const float* lp = a.lp + (long long)col * a.st_b;
  float* la_out = a.log_alpha + (long long)b * a.T * a.S;

  int cls[SPL];
  bool skip[SPL];  // transition from s-2 allowed
  float al[SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = j * 32 + lane;
    cls[j] = a.blank;
    skip[j] = false;
    if (s < S && (s & 1)) {
      cls[j] = a.targets[off + (s >> 1)];
      if (s >= 3) skip[j] = (a.targets[off + (s >> 1) - 1] != cls[j]);
    }
    al[j] = neg_inf();
  }
  if (Tb <= 0) {  // degenerate: no frames. ATen: log_alpha never initialised; treat as infeasible unless L==0.
    if (lane == 0) a.nll[b] = (L == 0) ? 0.f : INFINITY;
    return;
  }
  // t = 0
  for (int c = lane; c < a.V; c += 32) row[c] = lp[c];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = j * 32 + lane;
    if (s == 0) al[j] = row[a.blank];
    else if (s == 1 && S > 1) al[j] = row[cls[j]];
    if (s < S) la_out[s] = al[j];
  }
  // The recursion is a chain of T dependent steps of one warp: the row of log-probs of step t + 1 is loaded into registers
  // while step t is computed, so that the ~1 us global-load latency is off the chain (V <= 128; wider rows load in place).
  const bool pf = a.V <= 4 * 32;
  float nxt[4] = {0.f, 0.f, 0.f, 0.f};
  if (pf && Tb > 1) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < a.V) nxt[k] = lp[a.st_t + lane + 32 * k];
  }
  for (int t = 1; t < Tb; ++t) {
    __syncwarp();
    const float* lpt = lp + (long long)t * a.st_t;
    if (pf) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lane + 32 * k < a.V) row[lane + 32 * k] = nxt[k];
      if (t + 1 < Tb) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (lane + 32 * k < a.V) nxt[k] = lpt[a.st_t + lane + 32 * k];
      }
    } else {
      for (int c = lane; c < a.V; c += 32) row[c] = lpt[c];
    }
    __syncwarp();
    float nw[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = j * 32 + lane;
      const float up1 = __shfl_up_sync(FULL_MASK, al[j], 1);
      const float up2 = __shfl_up_sync(FULL_MASK, al[j], 2);
      float p31 = neg_inf(), p30 = neg_inf();
      if (j > 0) {
        p31 = __shfl_sync(FULL_MASK, al[j - 1], 31);
        p30 = __shfl_sync(FULL_MASK, al[j - 1], 30);
      }
      const float la1 = al[j];
      const float la2 = (lane == 0) ? p31 : up1;
      float la3 = (lane == 0) ? p30 : (lane == 1 ? p31 : up2);
      if (!skip[j]) la3 = neg_inf();
      nw[j] = (s < S) ? lse3(la1, la2, la3) + row[cls[j]] : neg_inf();
    }
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      al[j] = nw[j];
      const int s = j * 32 + lane;
      if (s < S) la_out[(long long)t * a.S + s] = al[j];
    }
  }
  __syncwarp();
  if (lane == 0) {
    const float* last = la_out + (long long)(Tb - 1) * a.S;
    float r;
    if (L == 0) {
      r = -last[0];
    } else {
      const float l1 = last[2 * L], l2 = last[2 * L - 1];
      float m = fmaxf(l1, l2);
      if (m == neg_inf()) m = 0.f;
      r = -(logf(expf(l1 - m) + expf(l2 - m)) + m);
    }
    a.nll[b] = r;
  }
}

struct CtcBwdArgs {
  CtcArgs f;
  const float* grad_out;  // (B) for reduction none, (1) otherwise
  int reduction;          // 0 none, 1 mean, 2 sum
  int zero_infinity;
  float* grad;            // (T, B_full, V) with strides gst_t, gst_b
  long long gst_t, gst_b;
};

template <int SPL>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) ctc_beta_grad_kernel(CtcBwdArgs g) {
  qeb_pdl_sync();
  const CtcArgs& a = g.f;
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerBlock + warp;
  if (b >= a.B) return;
  float* row = smem + warp * 3 * a.V;
  float* accm = row + a.V;
  float* accs = accm + a.V;
  const int L = a.target_lengths[b];
  int Tb = a.input_lengths[b];
  if (Tb < 0) Tb = 0;
  const int off = a.tgt_offsets[b];
  const int S = 2 * L + 1;
  const int col = a.batch_index ? a.batch_index[b] : b;
  const float* lp = a.lp + (long long)col * a.st_b;
  const float* la_in = a.log_alpha + (long long)b * a.T * a.S;
  float* gr = g.grad + (long long)col * g.gst_b;
  const float nll = a.nll[b];
  float go;
  if (g.reduction == 0) go = g.grad_out[b];
  else if (g.reduction == 1) go = g.grad_out[0] / (float)(L > 1 ? L : 1) / (float)a.B;
  else go = g.grad_out[0];
  const bool zero_all = g.zero_infinity && (nll == INFINITY);

  // frames past the input length carry zero gradient (with batch_index the caller's buffer is zero-filled and ADDED into:
  // one column may appear in several rows - the history depths of weighted_ctc_loss, tracking_utils.py:59-75)
  const bool add = a.batch_index != nullptr;
  if (!add)
    for (int t = Tb; t < a.T; ++t)
      for (int c = lane; c < a.V; c += 32) gr[(long long)t * g.gst_t + c] = 0.f;
  if (Tb == 0) return;

  int cls[SPL];
  bool skip[SPL];  // transition to s+2 allowed
  float be[SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = j * 32 + lane;
    cls[j] = a.blank;
    skip[j] = false;
    if (s < S && (s & 1)) {
      cls[j] = a.targets[off + (s >> 1)];
      if (s + 2 < S) skip[j] = (a.targets[off + (s >> 1) + 1] != cls[j]);
    }
    be[j] = neg_inf();
  }

  // as in ctc_alpha_kernel: the log-prob row and the alpha values of step t - 1 are loaded while step t is computed
  const bool pf = a.V <= 4 * 32;
  float nxt[4] = {0.f, 0.f, 0.f, 0.f}, la_nxt[SPL];
  {
    const float* lpt = lp + (long long)(Tb - 1) * a.st_t;
    if (pf) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lane + 32 * k < a.V) nxt[k] = lpt[lane + 32 * k];
    }
#pragma unroll
    for (int j = 0; j < SPL; ++j) la_nxt[j] = (j * 32 + lane < S) ? la_in[(long long)(Tb - 1) * a.S + j * 32 + lane] : neg_inf();
  }
  for (int t = Tb - 1; t >= 0; --t) {
    __syncwarp();
    const float* lpt = lp + (long long)t * a.st_t;
    float la_cur[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) la_cur[j] = la_nxt[j];
    if (pf) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (lane + 32 * k < a.V) row[lane + 32 * k] = nxt[k];
      if (t > 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (lane + 32 * k < a.V) nxt[k] = lpt[lane + 32 * k - a.st_t];
      }
    } else {
      for (int c = lane; c < a.V; c += 32) row[c] = lpt[c];
    }
    if (t > 0) {
#pragma unroll
      for (int j = 0; j < SPL; ++j)
        if (j * 32 + lane < S) la_nxt[j] = la_in[(long long)(t - 1) * a.S + j * 32 + lane];
    }
    for (int c = lane; c < a.V; c += 32) {
      accm[c] = neg_inf();
      accs[c] = 0.f;
    }
    __syncwarp();
    float nw[SPL];
    if (t == Tb - 1) {
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        const int s = j * 32 + lane;
        nw[j] = neg_inf();
        if (s == 2 * L) nw[j] = row[a.blank];
        else if (L > 0 && s == 2 * L - 1) nw[j] = row[cls[j]];
      }
    } else {
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        const int s = j * 32 + lane;
        const float dn1 = __shfl_down_sync(FULL_MASK, be[j], 1);
        const float dn2 = __shfl_down_sync(FULL_MASK, be[j], 2);
        float n0 = neg_inf(), n1 = neg_inf();
        if (j + 1 < SPL) {
          n0 = __shfl_sync(FULL_MASK, be[j + 1], 0);
          n1 = __shfl_sync(FULL_MASK, be[j + 1], 1);
        }
        const float lb1 = be[j];
        const float lb2 = (lane == 31) ? n0 : dn1;
        float lb3 = (lane == 31) ? n1 : (lane == 30 ? n0 : dn2);
        if (!skip[j]) lb3 = neg_inf();
        nw[j] = (s < S) ? lse3(lb1, lb2, lb3) + row[cls[j]] : neg_inf();
      }
    }
    float lab[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      be[j] = nw[j];
      const int s = j * 32 + lane;
      lab[j] = neg_inf();
      if (s < S) {
        lab[j] = la_cur[j] + be[j];
        if (lab[j] != neg_inf() && lab[j] == lab[j]) atomic_max_float(&accm[cls[j]], lab[j]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = j * 32 + lane;
      if (s < S && lab[j] != neg_inf() && lab[j] == lab[j]) atomicAdd(&accs[cls[j]], expf(lab[j] - accm[cls[j]]));
    }
    __syncwarp();
    for (int c = lane; c < a.V; c += 32) {
      const float res = (accm[c] == neg_inf()) ? neg_inf() : logf(accs[c]) + accm[c];
      const float l = row[c];
      float v = (expf(l) - expf(res + nll - l)) * go;
      if (zero_all) v = 0.f;
      if (add) atomicAdd(&gr[(long long)t * g.gst_t + c], v);
      else gr[(long long)t * g.gst_t + c] = v;
    }
  }
}

// loss reduction: out[0] = mean_b(nll_b / max(1, L_b)) (mean) or sum_b nll_b (sum); single warp, fixed order.
__global__ void ctc_reduce_kernel(const float* nll, const int* target_lengths, int B, int reduction, int zero_infinity,
                                  float* out) {
  qeb_pdl_sync();
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    float v = nll[b];
    if (zero_infinity && v == INFINITY) v = 0.f;
    if (reduction == 1) {
      const int L = target_lengths[b];
      v = v / (float)(L > 1 ? L : 1);
    }
    acc += v;
  }
  acc = warp_sum(acc);
  if (threadIdx.x == 0) out[0] = (reduction == 1) ? acc / (float)B : acc;
}

__global__ void ctc_zero_inf_kernel(float* nll, int B) {
  qeb_pdl_sync();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B && nll[b] == INFINITY) nll[b] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// log-softmax over the last (class) dimension, one warp per row
__global__ void log_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int V) {
  qeb_pdl_sync();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* xr = x + r * V;
  float m = neg_inf();
  for (int c = lane; c < V; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(xr[c] - m);
  s = warp_sum(s);
  const float ls = logf(s);
  for (int c = lane; c < V; c += 32) y[r * V + c] = xr[c] - m - ls;
}

// dx = dy - exp(y) * sum(dy)
__global__ void log_softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                       float* __restrict__ dx, long long rows, int V) {
  qeb_pdl_sync();
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += dy[r * V + c];
  s = warp_sum(s);
  for (int c = lane; c < V; c += 32) dx[r * V + c] = dy[r * V + c] - expf(y[r * V + c]) * s;
}


// ---------------------------------------------------------------------------------------------------------------------
// Block-per-sequence variants (the default whenever the panels fit in shared memory). The recursions above are chains of
// T dependent steps of ONE warp, and at B = 64 the whole loss is 64 such chains: what matters is the length of one step.
// Here a block of four warps first stages the sequence's whole (T, V) log-prob panel (and, in the backward kernel, its
// (T, S) log-alpha panel) with 4-byte cp.async copies - one global round trip instead of one per step - then warp 0 runs
// the chain out of shared memory. In the backward kernel the chain only stores log(alpha * beta); the per-class
// log-sum-exp and the gradient rows (the expensive part: shared-memory atomics, V exp/log per frame) are computed
// afterwards by all four warps, one frame per warp at a time, off the chain.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSeqThreads = 128;
constexpr size_t kSeqSmemMax = 160 * 1024;

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

template <int SPL>
__global__ void __launch_bounds__(kSeqThreads) ctc_alpha_seq_kernel(CtcArgs a) {
  qeb_pdl_sync();
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int L = a.target_lengths[b];
  const int Tb = min(a.input_lengths[b], a.T);
  const int off = a.tgt_offsets[b];
  const int S = 2 * L + 1;
  const int col = a.batch_index ? a.batch_index[b] : b;
  const float* lp = a.lp + (long long)col * a.st_b;
  float* la_out = a.log_alpha + (long long)b * a.T * a.S;
  if (Tb <= 0) {  // as in ctc_alpha_kernel
    if (threadIdx.x == 0) a.nll[b] = (L == 0) ? 0.f : INFINITY;
    return;
  }
  float* panel = smem;  // [Tb][V]
  for (int t = warp; t < Tb; t += kSeqThreads / 32)
    for (int c = lane; c < a.V; c += 32) cp_async4(panel + t * a.V + c, lp + (long long)t * a.st_t + c);
  int cls[SPL];
  bool skip[SPL];
  float al[SPL];
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = j * 32 + lane;
    cls[j] = a.blank;
    skip[j] = false;
    if (s < S && (s & 1)) {
      cls[j] = a.targets[off + (s >> 1)];
      if (s >= 3) skip[j] = (a.targets[off + (s >> 1) - 1] != cls[j]);
    }
    al[j] = neg_inf();
  }
  cp_async_wait_all();
  __syncthreads();
  if (warp != 0) return;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = j * 32 + lane;
    if (s == 0) al[j] = panel[a.blank];
    else if (s == 1 && S > 1) al[j] = panel[cls[j]];
    if (s < S) la_out[s] = al[j];
  }
  for (int t = 1; t < Tb; ++t) {
    const float* row = panel + t * a.V;
    float nw[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = j * 32 + lane;
      const float emit = row[cls[j]];
      const float up1 = __shfl_up_sync(FULL_MASK, al[j], 1);
      const float up2 = __shfl_up_sync(FULL_MASK, al[j], 2);
      float p31 = neg_inf(), p30 = neg_inf();
      if (j > 0) {
        p31 = __shfl_sync(FULL_MASK, al[j - 1], 31);
        p30 = __shfl_sync(FULL_MASK, al[j - 1], 30);
      }
      const float la1 = al[j];
      const float la2 = (lane == 0) ? p31 : up1;
      float la3 = (lane == 0) ? p30 : (lane == 1 ? p31 : up2);
      if (!skip[j]) la3 = neg_inf();
      nw[j] = (s < S) ? lse3(la1, la2, la3) + emit : neg_inf();
    }
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      al[j] = nw[j];
      const int s = j * 32 + lane;
      if (s < S) la_out[(long long)t * a.S + s] = al[j];
    }
  }
  // -log(alpha_T(2L) + alpha_T(2L - 1)) from the registers that hold the last frame
  float l1 = neg_inf(), l2 = neg_inf();
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const float v1 = __shfl_sync(FULL_MASK, al[j], (2 * L) & 31);
    const float v2 = __shfl_sync(FULL_MASK, al[j], (2 * L - 1) & 31);
    if (j == ((2 * L) >> 5)) l1 = v1;
    if (L > 0 && j == ((2 * L - 1) >> 5)) l2 = v2;
  }
  if (lane == 0) {
    float r;
    if (L == 0) {
      r = -l1;
    } else {
      float m = fmaxf(l1, l2);
      if (m == neg_inf()) m = 0.f;
      r = -(logf(expf(l1 - m) + expf(l2 - m)) + m);
    }
    a.nll[b] = r;
  }
}

template <int SPL>
__global__ void __launch_bounds__(kSeqThreads) ctc_beta_grad_seq_kernel(CtcBwdArgs g) {
  qeb_pdl_sync();
  const CtcArgs& a = g.f;
  extern __shared__ float smem[];
  constexpr int kWarps = kSeqThreads / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int L = a.target_lengths[b];
  const int Tb = max(0, min(a.input_lengths[b], a.T));
  const int off = a.tgt_offsets[b];
  const int S = 2 * L + 1;
  const int col = a.batch_index ? a.batch_index[b] : b;
  const float* lp = a.lp + (long long)col * a.st_b;
  const float* la_in = a.log_alpha + (long long)b * a.T * a.S;
  float* gr = g.grad + (long long)col * g.gst_b;
  const float nll = a.nll[b];
  float go;
  if (g.reduction == 0) go = g.grad_out[b];
  else if (g.reduction == 1) go = g.grad_out[0] / (float)(L > 1 ? L : 1) / (float)a.B;
  else go = g.grad_out[0];
  const bool zero_all = g.zero_infinity && (nll == INFINITY);

  float* panel = smem;                    // [T][V]   log-probs
  float* ab = panel + a.T * a.V;          // [T][a.S] log alpha, then log(alpha * beta)
  float* accm = ab + a.T * a.S + warp * 2 * a.V;   // per warp: running max / sum per class
  float* accs = accm + a.V;

  const bool add = a.batch_index != nullptr;   // see ctc_beta_grad_kernel
  if (!add)
    for (int t = Tb + warp; t < a.T; t += kWarps)   // frames past the input length carry zero gradient
      for (int c = lane; c < a.V; c += 32) gr[(long long)t * g.gst_t + c] = 0.f;
  if (Tb == 0) return;
  for (int t = warp; t < Tb; t += kWarps) {
    for (int c = lane; c < a.V; c += 32) cp_async4(panel + t * a.V + c, lp + (long long)t * a.st_t + c);
    for (int s = lane; s < S; s += 32) cp_async4(ab + t * a.S + s, la_in + (long long)t * a.S + s);
  }
  int cls[SPL];
  bool skip[SPL];  // transition to s+2 allowed
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    const int s = j * 32 + lane;
    cls[j] = a.blank;
    skip[j] = false;
    if (s < S && (s & 1)) {
      cls[j] = a.targets[off + (s >> 1)];
      if (s + 2 < S) skip[j] = (a.targets[off + (s >> 1) + 1] != cls[j]);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  if (warp == 0) {   // the chain: beta_t from beta_{t+1}, all operands in shared memory
    float be[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) be[j] = neg_inf();
    for (int t = Tb - 1; t >= 0; --t) {
      const float* row = panel + t * a.V;
      float nw[SPL];
      if (t == Tb - 1) {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          const int s = j * 32 + lane;
          nw[j] = neg_inf();
          if (s == 2 * L) nw[j] = row[a.blank];
          else if (L > 0 && s == 2 * L - 1) nw[j] = row[cls[j]];
        }
      } else {
#pragma unroll
        for (int j = 0; j < SPL; ++j) {
          const int s = j * 32 + lane;
          const float emit = row[cls[j]];
          const float dn1 = __shfl_down_sync(FULL_MASK, be[j], 1);
          const float dn2 = __shfl_down_sync(FULL_MASK, be[j], 2);
          float n0 = neg_inf(), n1 = neg_inf();
          if (j + 1 < SPL) {
            n0 = __shfl_sync(FULL_MASK, be[j + 1], 0);
            n1 = __shfl_sync(FULL_MASK, be[j + 1], 1);
          }
          const float lb1 = be[j];
          const float lb2 = (lane == 31) ? n0 : dn1;
          float lb3 = (lane == 31) ? n1 : (lane == 30 ? n0 : dn2);
          if (!skip[j]) lb3 = neg_inf();
          nw[j] = (s < S) ? lse3(lb1, lb2, lb3) + emit : neg_inf();
        }
      }
#pragma unroll
      for (int j = 0; j < SPL; ++j) {
        be[j] = nw[j];
        const int s = j * 32 + lane;
        if (s < S) ab[t * a.S + s] += be[j];
      }
    }
  }
  __syncthreads();

  for (int t = warp; t < Tb; t += kWarps) {   // gradient rows, one frame per warp at a time
    const float* row = panel + t * a.V;
    for (int c = lane; c < a.V; c += 32) {
      accm[c] = neg_inf();
      accs[c] = 0.f;
    }
    float lab[SPL];
    bool ok[SPL];
    float bm = neg_inf();
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int s = j * 32 + lane;
      lab[j] = (s < S) ? ab[t * a.S + s] : neg_inf();
      ok[j] = s < S && lab[j] != neg_inf() && lab[j] == lab[j];
      if (ok[j] && cls[j] == a.blank) bm = fmaxf(bm, lab[j]);
    }
    __syncwarp();
    // the blank class collects every other state: warp reduction instead of 32 atomics on one address
    bm = warp_max_f(bm);
    float bs = 0.f;
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      if (ok[j] && cls[j] == a.blank) bs += expf(lab[j] - bm);
      else if (ok[j]) atomic_max_float(&accm[cls[j]], lab[j]);
    }
    bs = warp_sum_f(bs);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SPL; ++j)
      if (ok[j] && cls[j] != a.blank) atomicAdd(&accs[cls[j]], expf(lab[j] - accm[cls[j]]));
    if (lane == 0 && bm != neg_inf()) {
      accm[a.blank] = bm;
      accs[a.blank] = bs;
    }
    __syncwarp();
    for (int c = lane; c < a.V; c += 32) {
      const float res = (accm[c] == neg_inf()) ? neg_inf() : logf(accs[c]) + accm[c];
      const float l = row[c];
      float v = (expf(l) - expf(res + nll - l)) * go;
      if (zero_all) v = 0.f;
      if (add) atomicAdd(&gr[(long long)t * g.gst_t + c], v);
      else gr[(long long)t * g.gst_t + c] = v;
    }
    __syncwarp();
  }
}

inline bool ctc_seq_enabled() {
  static const bool on = !(getenv("QEB_CTC_SEQ") && atoi(getenv("QEB_CTC_SEQ")) == 0);
  return on;
}
template <typename K>
int ctc_seq_smem_attr(K kernel) {
  QEB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeqSmemMax));
  return QEB_OK;
}

template <typename F>
int dispatch_spl(int S, F&& f) {
  if (S <= 64) return f(std::integral_constant<int, 2>());
  if (S <= 128) return f(std::integral_constant<int, 4>());
  return f(std::integral_constant<int, kMaxSPL>());
}

}  // namespace

QEB_API size_t qeb_ctc_workspace_bytes(int B, int T, int max_target_len) {
  return (size_t)B * T * (2 * (size_t)max_target_len + 1) * sizeof(float);
}

// Forward. log_alpha is the workspace kept for the backward call. loss_out may be NULL (reduction none).
QEB_API int qeb_ctc_fwd(const float* log_probs, long long st_t, long long st_b, const int* batch_index,
                        const int* targets, const int* tgt_offsets, const int* input_lengths,
                        const int* target_lengths, int B, int T, int V, int blank, int max_target_len,
                        int reduction, int zero_infinity, float* log_alpha, float* nll, float* loss_out,
                        void* stream) {
  QEB_REQUIRE(log_probs && input_lengths && target_lengths && tgt_offsets && log_alpha && nll, "ctc_fwd: null pointer");
  QEB_REQUIRE(B > 0 && T > 0 && V > 0 && blank >= 0 && blank < V, "ctc_fwd: bad sizes B=%d T=%d V=%d blank=%d", B, T, V, blank);
  QEB_REQUIRE(max_target_len >= 0 && 2 * max_target_len + 1 <= 32 * kMaxSPL, "ctc_fwd: target length %d > 127 unsupported", max_target_len);
  QEB_REQUIRE(max_target_len == 0 || targets, "ctc_fwd: null targets");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof("ctc_fwd", st, 0.0, 4.0 * B * T * V);
  CtcArgs a{log_probs, st_t, st_b, batch_index, targets, tgt_offsets, input_lengths, target_lengths,
            B, T, V, blank, 2 * max_target_len + 1, log_alpha, nll};
  const int grid = qeb_cdiv(B, kWarpsPerBlock);
  const size_t smem = (size_t)kWarpsPerBlock * V * sizeof(float);
  QEB_REQUIRE(smem <= 48 * 1024, "ctc_fwd: V=%d too large", V);
  const size_t seq_smem = (size_t)T * V * sizeof(float);
  int rc = dispatch_spl(a.S, [&](auto spl) {
    constexpr int kSpl = decltype(spl)::value;
    if (ctc_seq_enabled() && seq_smem <= kSeqSmemMax) {   // block per sequence, panels in shared memory
      static const int attr = ctc_seq_smem_attr(ctc_alpha_seq_kernel<kSpl>);
      if (attr) return attr;
      QEB_CUDA(qeb_launch(ctc_alpha_seq_kernel<kSpl>, B, kSeqThreads, seq_smem, st, a));
      return 0;
    }
    QEB_CUDA(qeb_launch(ctc_alpha_kernel<kSpl>, grid, kWarpsPerBlock * 32, smem, st, a));
    return 0;
  });
  if (rc) return rc;
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  if (loss_out && reduction != 0) {
    QEB_CUDA(qeb_launch(ctc_reduce_kernel, 1, 32, 0, st, nll, target_lengths, B, reduction, zero_infinity, loss_out));
    QEB_LAUNCH_CHECK();
    qeb_count_launch();
  }
  return QEB_OK;
}

// Backward: gradient w.r.t. log_probs (ATen convention, see header). grad has the full (T, Bfull, V) shape; with
// batch_index only the listed columns are written (caller zero-fills the rest).
QEB_API int qeb_ctc_bwd(const float* log_probs, long long st_t, long long st_b, const int* batch_index,
                        const int* targets, const int* tgt_offsets, const int* input_lengths,
                        const int* target_lengths, int B, int T, int V, int blank, int max_target_len,
                        int reduction, int zero_infinity, const float* log_alpha, const float* nll,
                        const float* grad_out, float* grad, long long gst_t, long long gst_b, void* stream) {
  QEB_REQUIRE(log_probs && input_lengths && target_lengths && tgt_offsets && log_alpha && nll && grad_out && grad,
              "ctc_bwd: null pointer");
  QEB_REQUIRE(B > 0 && T > 0 && V > 0, "ctc_bwd: bad sizes");
  QEB_REQUIRE(max_target_len >= 0 && 2 * max_target_len + 1 <= 32 * kMaxSPL, "ctc_bwd: target length %d > 127 unsupported", max_target_len);
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof("ctc_bwd", st, 0.0, 8.0 * B * T * V);
  CtcBwdArgs g;
  g.f = CtcArgs{log_probs, st_t, st_b, batch_index, targets, tgt_offsets, input_lengths, target_lengths,
                B, T, V, blank, 2 * max_target_len + 1, const_cast<float*>(log_alpha), const_cast<float*>(nll)};
  g.grad_out = grad_out;
  g.reduction = reduction;
  g.zero_infinity = zero_infinity;
  g.grad = grad;
  g.gst_t = gst_t;
  g.gst_b = gst_b;
  const int grid = qeb_cdiv(B, kWarpsPerBlock);
  const size_t smem = (size_t)kWarpsPerBlock * 3 * V * sizeof(float);
  QEB_REQUIRE(smem <= 48 * 1024, "ctc_bwd: V=%d too large", V);
  const size_t seq_smem = ((size_t)T * (V + g.f.S) + (size_t)(kSeqThreads / 32) * 2 * V) * sizeof(float);
  int rc = dispatch_spl(g.f.S, [&](auto spl) {
    constexpr int kSpl = decltype(spl)::value;
    if (ctc_seq_enabled() && seq_smem <= kSeqSmemMax) {
      static const int attr = ctc_seq_smem_attr(ctc_beta_grad_seq_kernel<kSpl>);
      if (attr) return attr;
      QEB_CUDA(qeb_launch(ctc_beta_grad_seq_kernel<kSpl>, B, kSeqThreads, seq_smem, st, g));
      return 0;
    }
    QEB_CUDA(qeb_launch(ctc_beta_grad_kernel<kSpl>, grid, kWarpsPerBlock * 32, smem, st, g));
    return 0;
  });
  if (rc) return rc;
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

QEB_API int qeb_log_softmax_fwd(const float* x, float* y, long long rows, int V, void* stream) {
  QEB_REQUIRE(x && y && rows > 0 && V > 0, "log_softmax_fwd: bad args");
  ProfScope prof("log_softmax", (cudaStream_t)stream, 0.0, 8.0 * rows * V);
  QEB_CUDA(qeb_launch(log_softmax_fwd_kernel, qeb_cdiv(rows, 8), 256, 0, (cudaStream_t)stream, x, y, rows, V));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

QEB_API int qeb_log_softmax_bwd(const float* y, const float* dy, float* dx, long long rows, int V, void* stream) {
  QEB_REQUIRE(y && dy && dx && rows > 0 && V > 0, "log_softmax_bwd: bad args");
  ProfScope prof("log_softmax", (cudaStream_t)stream, 0.0, 12.0 * rows * V);
  QEB_CUDA(qeb_launch(log_softmax_bwd_kernel, qeb_cdiv(rows, 8), 256, 0, (cudaStream_t)stream, y, dy, dx, rows, V));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}
