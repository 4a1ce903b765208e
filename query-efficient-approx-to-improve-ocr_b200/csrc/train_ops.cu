// Training-step glue kernels around the two networks:
//   - MSE of the preprocessor output against all-ones, the secondary loss of TrainNNPrep._get_loss
//     (train_nn_patch.py:177-185, train_nn_area.py:173-182: MSELoss()(img_preds, torch.ones(...)))
//   - Adam with L2-coupled weight decay over many tensors in one launch (torch.optim.Adam as constructed at
//     train_nn_patch.py:146-152 / train_nn_area.py:149-154; defaults betas (0.9, 0.999), eps 1e-8, amsgrad off).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) mse_ones_fwd_kernel(const float* __restrict__ x, long long n, float inv_n,
                                                                float* __restrict__ loss) {
  qeb_pdl_sync();
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    const float d = x[i] - 1.f;
    s = fmaf(d, d, s);
  }
  s = warp_sum(s);
  __shared__ float part[kThreads / 32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v * inv_n);
  }
}

__global__ void __launch_bounds__(kThreads) mse_ones_bwd_kernel(const float* __restrict__ x, long long n, float two_inv_n,
                                                                const float* __restrict__ gout, float* __restrict__ dx) {
  qeb_pdl_sync();
  const float g = __ldg(gout) * two_inv_n;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
    dx[i] = (x[i] - 1.f) * g;
}

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
  long long chunk0;  // first 1024-element chunk of this tensor in the global chunk numbering
};

__global__ void __launch_bounds__(kThreads) adam_multi_kernel(const AdamTensor* __restrict__ tab, int n_tensors,
                                                              long long n_chunks, float lr, float beta1, float beta2, float eps,
                                                              float weight_decay, float bc1, float bc2_sqrt) {
  qeb_pdl_sync();
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    int lo = 0, hi = n_tensors - 1;  // last tensor whose chunk0 <= c
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].chunk0 <= c) lo = mid; else hi = mid - 1;
    }
    const AdamTensor t = tab[lo];
    const long long base = (c - t.chunk0) * 1024;
#pragma unroll
    for (int k = 0; k < 1024 / kThreads; ++k) {
      const long long i = base + k * kThreads + threadIdx.x;
      if (i < t.n) {
        const float p = t.p[i];
        const float g = fmaf(weight_decay, p, t.g[i]);
        const float m = fmaf(beta1, t.m[i], (1.f - beta1) * g);
        const float v = fmaf(beta2, t.v[i], (1.f - beta2) * g * g);
        t.m[i] = m;
        t.v[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        t.p[i] = p - (lr / bc1) * (m / denom);
      }
    }
  }
}

}  // namespace

// loss (scalar, device) = mean((x - 1)^2) over n elements
QEB_API int qeb_mse_ones_fwd(const float* x, long long n, float* loss, void* stream) {
  QEB_REQUIRE(x && loss && n > 0, "mse_ones_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof("mse", st, 0.0, 4.0 * n);
  QEB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  QEB_CUDA(qeb_launch(mse_ones_fwd_kernel, qeb_grid(n, kThreads, 4), kThreads, 0, st, x, n, 1.f / (float)n, loss));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

// dx = 2 (x - 1) / n * grad_out[0]
QEB_API int qeb_mse_ones_bwd(const float* x, long long n, const float* grad_out, float* dx, void* stream) {
  QEB_REQUIRE(x && grad_out && dx && n > 0, "mse_ones_bwd: bad arguments");
  ProfScope prof("mse", (cudaStream_t)stream, 0.0, 8.0 * n);
  QEB_CUDA(qeb_launch(mse_ones_bwd_kernel, qeb_grid(n, kThreads), kThreads, 0, (cudaStream_t)stream, x, n, 2.f / (float)n, grad_out, dx));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

QEB_API int qeb_adam_table_entry_bytes(void) { return (int)sizeof(AdamTensor); }

// table (device): n_tensors entries {p, g, m, v, n, chunk0}, chunk0 = exclusive prefix sum of ceil(n/1024).
// step >= 1 is the step count AFTER this update (torch's bias corrections 1 - beta^step).
QEB_API int qeb_adam_multi(const void* table, int n_tensors, long long n_chunks, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int step, void* stream) {
  QEB_REQUIRE(table && n_tensors > 0 && n_chunks > 0 && step >= 1, "adam_multi: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  long long g = n_chunks;
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  ProfScope prof("adam", (cudaStream_t)stream, 0.0, 28.0 * 1024 * n_chunks);
  QEB_CUDA(qeb_launch(adam_multi_kernel, (int)g, kThreads, 0, (cudaStream_t)stream, static_cast<const AdamTensor*>(table), n_tensors, n_chunks, lr,
                                                                   beta1, beta2, eps, weight_decay, (float)bc1, (float)sqrt(bc2)));
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}
