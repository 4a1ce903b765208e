// Bidirectional LSTM recurrence (hidden size 256) as persistent thread-block-cluster kernels.
// Reference: nn.LSTM(512, 256, 2, bidirectional=True), models/model_crnn.py:9,19 (gate order i,f,g,o; h0 = c0 = 0).
//
// The input projections x*W_ih^T + b_ih + b_hh are tensor-core GEMMs (conv_tc.cu); this file runs the T dependent
// steps. One cluster of 8 CTAs serves one (direction, chunk of kBC batch rows); CTA r of the cluster owns hidden
// units [32r, 32r+32), i.e. 128 of the 1024 gate rows, whose W_hh slice (128 x 256 fp32 = 128 KB) stays in shared
// memory for the whole sequence. Per step every CTA computes its gate pre-activations for the chunk (k split over
// the 8 warps, reduced through shared memory), applies the gate non-linearities, and broadcasts its 32 new h values
// per batch row to the other 7 CTAs through distributed shared memory, double-buffered so that one cluster barrier
// per step is enough. The backward kernel mirrors it: dgates -> partial dh over the owned rows -> reduce-scatter of
// the partial sums over DSMEM.
#include "nn.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kH = 256;       // hidden size
constexpr int kCluster = 8;   // CTAs per cluster
constexpr int kUnits = kH / kCluster;  // 32 hidden units per CTA
constexpr int kRows = 4 * kUnits;      // 128 gate rows per CTA
constexpr int kBC = 8;        // batch rows per cluster
constexpr int kThreads = 256;

constexpr size_t kSmemW = (size_t)kRows * kH * sizeof(float);            // 128 KB
constexpr size_t kSmemH = 2ull * kH * kBC * sizeof(float);               // 16 KB  h double buffer [2][k][b]
constexpr size_t kSmemPart = 8ull * kBC * kRows * sizeof(float);         // 32 KB  partial sums [ks][b][row]
constexpr size_t kSmemFwd = kSmemW + kSmemH + kSmemPart;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct LstmArgs {
  float* gates;         // (T,B,2,1024)
  const float* w_hh[2]; // (1024,256) per direction
  float* cells;         // (T,B,2,256)
  float* y;             // (T,B,512)
  const float* dy;      // (T,B,512), backward only
  int T, B;
};

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_fwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* Wt = reinterpret_cast<float*>(smem_raw);                    // [k][row]  (256 x 128)
  float* hbuf = reinterpret_cast<float*>(smem_raw + kSmemW);         // [2][k][b]
  float* part = reinterpret_cast<float*>(smem_raw + kSmemW + kSmemH);  // [ks][b][row]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  const int dir = cid & 1, b0 = (cid >> 1) * kBC;
  const int tid = threadIdx.x;

  // W_hh slice, transposed: Wt[k][q*32 + jl] = W_hh[q*256 + 32*rank + jl][k]. Thread = (row, k quad): 16-byte global
  // loads, eight in flight, conflict-free transposed stores (consecutive threads -> consecutive rows -> banks).
  const float* W = a.w_hh[dir];
  {
    const int lr = tid % kRows;
    const int grow = (lr / kUnits) * kH + rank * kUnits + (lr % kUnits);
    const float4* src = reinterpret_cast<const float4*>(W + (long long)grow * kH);
#pragma unroll 1
    for (int kq0 = tid / kRows; kq0 < kH / 4; kq0 += 8 * (kThreads / kRows)) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(src + kq0 + u * (kThreads / kRows));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = (kq0 + u * (kThreads / kRows)) * 4;
        Wt[(k + 0) * kRows + lr] = v[u].x; Wt[(k + 1) * kRows + lr] = v[u].y;
        Wt[(k + 2) * kRows + lr] = v[u].z; Wt[(k + 3) * kRows + lr] = v[u].w;
      }
    }
  }
  for (int i = tid; i < 2 * kH * kBC; i += kThreads) hbuf[i] = 0.f;
  cluster.sync();

  // roles: mat-vec thread (row quad lq, k slice ks) and cell thread (unit jl, batch row bl)
  const int lq = tid & 31, ks = tid >> 5;
  const int jl = tid & 31, bl = tid >> 5;
  const int b = b0 + bl;
  const bool live = b < a.B;
  float c = 0.f;

  for (int s = 0; s < a.T; ++s) {
    const int t = dir ? a.T - 1 - s : s;
    const float* hcur = hbuf + (s & 1) * kH * kBC;
    float* hnext = hbuf + ((s + 1) & 1) * kH * kBC;
    // prefetch this thread's four x-projections
    float gx[4] = {0.f, 0.f, 0.f, 0.f};
    float* grow = a.gates + (((long long)t * a.B + b) * 2 + dir) * (4 * kH) + rank * kUnits + jl;
    if (live) {
#pragma unroll
      for (int q = 0; q < 4; ++q) gx[q] = grow[q * kH];
    }
    // partial mat-vec over k in [32 ks, 32 ks + 32): acc[b][4 rows]
    float acc[kBC][4];
#pragma unroll
    for (int i = 0; i < kBC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const int k = ks * 32 + kk;
      const float4 w4 = *reinterpret_cast<const float4*>(Wt + k * kRows + lq * 4);
      const float4 h0 = *reinterpret_cast<const float4*>(hcur + k * kBC);
      const float4 h1 = *reinterpret_cast<const float4*>(hcur + k * kBC + 4);
      const float hv[kBC] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int i = 0; i < kBC; ++i) {
        acc[i][0] = fmaf(w4.x, hv[i], acc[i][0]);
        acc[i][1] = fmaf(w4.y, hv[i], acc[i][1]);
        acc[i][2] = fmaf(w4.z, hv[i], acc[i][2]);
        acc[i][3] = fmaf(w4.w, hv[i], acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < kBC; ++i)
      *reinterpret_cast<float4*>(part + (ks * kBC + i) * kRows + lq * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    // cell update for (unit jl, batch row bl)
    float pre[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = gx[q];
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) v += part[(k8 * kBC + bl) * kRows + q * kUnits + jl];
      pre[q] = v;
    }
    const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
    c = fg * c + ig * gg;
    const float h = og * tanhf(c);
    hnext[(rank * kUnits + jl) * kBC + bl] = live ? h : 0.f;
    __syncthreads();
    // broadcast this CTA's 32 x kBC slice (256 contiguous floats) to the other CTAs
    {
      const float4* src = reinterpret_cast<const float4*>(hnext + rank * kUnits * kBC);
      for (int i = tid; i < (kCluster - 1) * (kUnits * kBC / 4); i += kThreads) {
        int dst_rank = i / (kUnits * kBC / 4);
        const int e = i % (kUnits * kBC / 4);
        dst_rank += (dst_rank >= rank);
        float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(hnext + rank * kUnits * kBC, dst_rank));
        dst[e] = src[e];
      }
    }
    // arrive (release) right after the DSMEM stores; the global stores below are issued after the fence, so the
    // barrier does not wait for them
    cluster.barrier_arrive();
    if (live) {
      grow[0] = ig; grow[kH] = fg; grow[2 * kH] = gg; grow[3 * kH] = og;
      a.cells[(((long long)t * a.B + b) * 2 + dir) * kH + rank * kUnits + jl] = c;
      a.y[((long long)t * a.B + b) * (2 * kH) + dir * kH + rank * kUnits + jl] = h;
    }
    cluster.barrier_wait();
  }
}

constexpr size_t kSmemDg = (size_t)kRows * kBC * sizeof(float);            // 4 KB   dgates [row][b]
constexpr size_t kSmemRecv = 2ull * kCluster * kBC * kUnits * sizeof(float);  // 16 KB  [2][src][b][unit]
constexpr size_t kSmemPartB = 4ull * kBC * kH * sizeof(float);            // 32 KB  [rs][b][k]
constexpr size_t kSmemBwd = kSmemW + kSmemDg + kSmemRecv + kSmemPartB;

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_bwd_kernel(LstmArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* Wr = reinterpret_cast<float*>(smem_raw);                              // [row][k]  (128 x 256)
  float* dgs = reinterpret_cast<float*>(smem_raw + kSmemW);                    // [row][b]
  float* recv = reinterpret_cast<float*>(smem_raw + kSmemW + kSmemDg);         // [2][src][b][unit]
  float* part = reinterpret_cast<float*>(smem_raw + kSmemW + kSmemDg + kSmemRecv);  // [rs][b][k]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  const int dir = cid & 1, b0 = (cid >> 1) * kBC;
  const int tid = threadIdx.x;

  const float* W = a.w_hh[dir];
#pragma unroll 4
  for (int i = tid; i < kRows * kH / 4; i += kThreads) {  // 16-byte copies, rows stay row-major
    const int lr = i / (kH / 4), kq = i % (kH / 4);
    const int grow = (lr / kUnits) * kH + rank * kUnits + (lr % kUnits);
    reinterpret_cast<float4*>(Wr)[lr * (kH / 4) + kq] = __ldg(reinterpret_cast<const float4*>(W + (long long)grow * kH) + kq);
  }
  for (int i = tid; i < 2 * kCluster * kBC * kUnits; i += kThreads) recv[i] = 0.f;
  cluster.sync();

  const int jl = tid & 31, bl = tid >> 5;        // cell thread
  const int kq = tid & 63, rs = tid >> 6;        // mat-vec thread: k quad, row slice (32 rows)
  const int b = b0 + bl;
  const bool live = b < a.B;
  float dc = 0.f;

  for (int s = 0; s < a.T; ++s) {
    const int t = dir ? s : a.T - 1 - s;          // reverse of the forward order
    const int tp = dir ? t + 1 : t - 1;           // step whose c is c_prev
    const float* rcv = recv + ((s + 1) & 1) * kCluster * kBC * kUnits;  // written during step s-1 (zeros at s = 0)
    float* gptr = a.gates + (((long long)t * a.B + b) * 2 + dir) * (4 * kH) + rank * kUnits + jl;
    float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f;
    if (live) {
      float dh = a.dy[((long long)t * a.B + b) * (2 * kH) + dir * kH + rank * kUnits + jl];
#pragma unroll
      for (int r = 0; r < kCluster; ++r) dh += rcv[(r * kBC + bl) * kUnits + jl];
      const float ig = gptr[0], fg = gptr[kH], gg = gptr[2 * kH], og = gptr[3 * kH];
      const float ct = a.cells[(((long long)t * a.B + b) * 2 + dir) * kH + rank * kUnits + jl];
      const float cp = (tp >= 0 && tp < a.T) ? a.cells[(((long long)tp * a.B + b) * 2 + dir) * kH + rank * kUnits + jl] : 0.f;
      const float tc = tanhf(ct);
      d_o = dh * tc * og * (1.f - og);
      const float dct = dc + dh * og * (1.f - tc * tc);
      d_i = dct * gg * ig * (1.f - ig);
      d_g = dct * ig * (1.f - gg * gg);
      d_f = dct * cp * fg * (1.f - fg);
      dc = dct * fg;
      gptr[0] = d_i; gptr[kH] = d_f; gptr[2 * kH] = d_g; gptr[3 * kH] = d_o;
    }
    dgs[(0 * kUnits + jl) * kBC + bl] = d_i;
    dgs[(1 * kUnits + jl) * kBC + bl] = d_f;
    dgs[(2 * kUnits + jl) * kBC + bl] = d_g;
    dgs[(3 * kUnits + jl) * kBC + bl] = d_o;
    __syncthreads();
    // partial dh_prev[b][k] over the 32 rows of slice rs
    float acc[kBC][4];
#pragma unroll
    for (int i = 0; i < kBC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const int lr = rs * 32 + rr;
      const float4 w4 = *reinterpret_cast<const float4*>(Wr + lr * kH + kq * 4);
      const float4 g0 = *reinterpret_cast<const float4*>(dgs + lr * kBC);
      const float4 g1 = *reinterpret_cast<const float4*>(dgs + lr * kBC + 4);
      const float gv[kBC] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < kBC; ++i) {
        acc[i][0] = fmaf(w4.x, gv[i], acc[i][0]);
        acc[i][1] = fmaf(w4.y, gv[i], acc[i][1]);
        acc[i][2] = fmaf(w4.z, gv[i], acc[i][2]);
        acc[i][3] = fmaf(w4.w, gv[i], acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < kBC; ++i)
      *reinterpret_cast<float4*>(part + (rs * kBC + i) * kH + kq * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    // reduce the 4 row slices and scatter: CTA dst gets the k range [32 dst, 32 dst + 32) of every batch row
    float* rnext = recv + (s & 1) * kCluster * kBC * kUnits;
    for (int i = tid; i < kBC * kH; i += kThreads) {
      const int bb = i / kH, k = i % kH;
      const float v = part[(0 * kBC + bb) * kH + k] + part[(1 * kBC + bb) * kH + k] + part[(2 * kBC + bb) * kH + k] +
                      part[(3 * kBC + bb) * kH + k];
      const int dst_rank = k / kUnits;
      float* dst = cluster.map_shared_rank(rnext, dst_rank);
      dst[(rank * kBC + bb) * kUnits + (k % kUnits)] = v;
    }
    cluster.barrier_arrive();
    cluster.barrier_wait();
  }
}

int launch_cluster(const void* fn, size_t smem, int n_clusters, LstmArgs& args, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  void* kargs[] = {&args};
  QEB_CUDA(cudaLaunchKernelExC(&cfg, fn, kargs));
  qeb_count_launch();
  return QEB_OK;
}

}  // namespace

int lstm_layer_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* cells, float* y, int T, int B,
                   cudaStream_t st) {
  QEB_REQUIRE(gates && w_hh_fwd && w_hh_rev && cells && y && T > 0 && B > 0, "lstm_layer_fwd: bad arguments");
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFwd));
    attr = true;
  }
  LstmArgs a;
  a.gates = gates; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_rev; a.cells = cells; a.y = y; a.dy = nullptr; a.T = T; a.B = B;
  ProfScope prof("lstm_fwd", st, 2.0 * T * B * 2 * 1024 * 256, 4.0 * T * B * (2 * 2048 + 512 + 512));
  return launch_cluster((const void*)lstm_fwd_kernel, kSmemFwd, 2 * qeb_cdiv(B, kBC), a, st);
}

int lstm_layer_bwd(float* gates, const float* cells, const float* dy, const float* w_hh_fwd, const float* w_hh_rev, int T,
                   int B, cudaStream_t st) {
  QEB_REQUIRE(gates && w_hh_fwd && w_hh_rev && cells && dy && T > 0 && B > 0, "lstm_layer_bwd: bad arguments");
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBwd));
    attr = true;
  }
  LstmArgs a;
  a.gates = gates; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_rev; a.cells = const_cast<float*>(cells); a.y = nullptr; a.dy = dy;
  a.T = T; a.B = B;
  ProfScope prof("lstm_bwd", st, 2.0 * T * B * 2 * 1024 * 256, 4.0 * T * B * (2 * 2048 + 512 + 512));
  return launch_cluster((const void*)lstm_bwd_kernel, kSmemBwd, 2 * qeb_cdiv(B, kBC), a, st);
}

// C ABI (tests): one bidirectional layer of the recurrence
QEB_API int qeb_lstm_layer_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* cells, float* y, int T, int B,
                               void* stream) {
  return lstm_layer_fwd(gates, w_hh_fwd, w_hh_rev, cells, y, T, B, (cudaStream_t)stream);
}
QEB_API int qeb_lstm_layer_bwd(float* gates, const float* cells, const float* dy, const float* w_hh_fwd, const float* w_hh_rev,
                               int T, int B, void* stream) {
  return lstm_layer_bwd(gates, cells, dy, w_hh_fwd, w_hh_rev, T, B, (cudaStream_t)stream);
}
