// Bidirectional LSTM recurrence (hidden size 256) as persistent thread-block-cluster kernels on tcgen05.
// Reference: nn.LSTM(512, 256, 2, bidirectional=True), models/model_crnn.py:9,19 (gate order i,f,g,o; h0 = c0 = 0).
//
// The input projections x*W_ih^T + b_ih + b_hh are tensor-core GEMMs (conv_tc.cu); this file runs the T dependent
// steps. One cluster of 8 CTAs serves one (direction, chunk of kBC = 16 batch rows); CTA r of the cluster owns hidden
// units [32r, 32r+32), i.e. 128 of the 1024 gate rows, whose W_hh slice (128 x 256 fp32 = 128 KB) stays in shared memory
// for the whole sequence as the K-major, 128-byte-swizzled A operand of a tcgen05.mma (kind::tf32, M = 128).
//
// Forward step: D[128 gate rows][16 batch] = W_slice (128 x 256) * h_{t-1}^T (256 x 16) is 32 MMAs (K = 8) issued by one
// thread, accumulator in TMEM. The operand copies of W_hh and h are rounded to tf32 with round-to-nearest (the tensor core
// itself would truncate); y, c and the saved gate activations stay fp32. Warp q of each half-block reads gate q of its 32
// units for 8 batch rows from TMEM (tcgen05.ld), adds the x-projection, applies the non-linearity and passes the activated
// gates through 8 KB of shared memory to the cell threads, which update c and write the 32 new h values per batch row into
// the B-operand layout; CTA r owns exactly K chunk r of that operand (2 KB) and copies it to the same place in the other 7
// CTAs of the cluster with 16-byte distributed-shared-memory stores, double-buffered. One cluster barrier per step.
//
// Backward step: the cell threads form d(gates) for their units and write them as the B operand [16 batch][128 own gate
// rows]; dh_{t-1}[k][b] partial = W_slice^T (256 x 128) * dgates (128 x 16) is 2 x 16 MMAs (two M = 128 halves of k); the
// partial sums are read from TMEM and reduce-scattered over DSMEM: CTA dst receives the k range [32 dst, 32 dst + 32) from
// every CTA and sums the eight contributions when it forms dh in the next step.
#include "nn.cuh"
#include "tc_common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

using namespace tc;

constexpr int kH = 256;       // hidden size
constexpr int kCluster = 8;   // CTAs per cluster
constexpr int kUnits = kH / kCluster;  // 32 hidden units per CTA
constexpr int kRows = 4 * kUnits;      // 128 gate rows per CTA
constexpr int kBC = 16;       // batch rows per cluster = MMA N
constexpr int kThreads = 256;

constexpr int kATile = 128 * 128;        // 16 KB: 128 rows x one 32-element (128-byte) K chunk of an A operand
constexpr int kBTile = kBC * 128;        // 2 KB: the same K chunk of the B operand
constexpr size_t kSmemW = 8 * kATile;    // 128 KB
constexpr int kHBuf = 8 * kBTile;        // 16 KB: one copy of h (hi or lo) for all 256 k
constexpr size_t kSmemFwd = kSmemW + 2 * kHBuf + 4 * kBC * kUnits * sizeof(float) + 64 + 1024;
constexpr size_t kSmemBwd = kSmemW + 4 * kBTile + 3 * kCluster * kBC * kUnits * sizeof(float) + 64 + 1024;
constexpr uint32_t kTmemCols = 32;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// byte offset of element (row, k) of a K-major SWIZZLE_128B operand whose 32-element K chunks are `tile_bytes` apart
__device__ __forceinline__ uint32_t sw128_off(int row, int k, int tile_bytes) {
  return (uint32_t)((k >> 5) * tile_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + ((k & 3) << 2));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct LstmArgs {
  float* gates;         // (T,B,2,1024)
  const float* w_hh[2]; // (1024,256) per direction
  float* cells;         // (T,B,2,256)
  float* y;             // (T,B,512)
  const float* dy;      // (T,B,512), backward only
  int T, B;
  long long* tl;        // debugging aid (qeb_debug_set_timeline): clock64 stamps of cluster 0 / CTA 0, 8 per step
};

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_fwd_kernel(LstmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Ws = smem;                                   // A operand: 8 K chunks x [128 gate rows][128 B]
  uint8_t* Hs = smem + kSmemW;                          // B operand, double-buffered: [2][8 K chunks][16 batch][128 B]
  float* act = reinterpret_cast<float*>(Hs + 2 * kHBuf);  // activated gates [q][b][unit]
  uint64_t* bar = reinterpret_cast<uint64_t*>(act + 4 * kBC * kUnits);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  const int dir = cid & 1, b0 = (cid >> 1) * kBC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // W_hh slice -> swizzled A operand, rounded to tf32. Thread = (row, k quad): a warp reads 512 contiguous bytes of one
  // row and writes four full 128-byte rows of four K chunks.
  {
    const float* W = a.w_hh[dir];
    for (int i = tid; i < kRows * (kH / 4); i += kThreads) {
      const int lr = i / (kH / 4), kq = i % (kH / 4);
      const int grow = (lr / kUnits) * kH + rank * kUnits + (lr % kUnits);
      float4 v = __ldg(reinterpret_cast<const float4*>(W + (long long)grow * kH) + kq);
      v.x = tf32_rn(v.x); v.y = tf32_rn(v.y); v.z = tf32_rn(v.z); v.w = tf32_rn(v.w);
      *reinterpret_cast<float4*>(Ws + sw128_off(lr, kq * 4, kATile)) = v;
    }
  }
  for (int i = tid; i < 2 * kHBuf / 16; i += kThreads) reinterpret_cast<float4*>(Hs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  fence_proxy_async_all();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // roles. gate thread: gate q of unit `lane` for batch rows [8 bh, 8 bh + 8); cell thread: unit `lane`, batch rows warp, warp + 8
  const int q = warp & 3, bh = warp >> 2;
  float c[2] = {0.f, 0.f};
  constexpr uint32_t idesc = instr_desc_tf32(128, kBC, 0, 0);
  const long long gstride = 2 * 4 * kH;  // floats between consecutive batch rows of `gates`

  // x-projections of this thread's gate row for its 8 batch rows, loaded one step ahead
  float gx[8];
  auto load_gx = [&](int s, float* dst) {
    const int t = dir ? a.T - 1 - s : s;
    const float* g = a.gates + (((long long)t * a.B + b0 + bh * 8) * 2 + dir) * (4 * kH) + q * kH + rank * kUnits + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = (b0 + bh * 8 + j < a.B) ? __ldg(g + j * gstride) : 0.f;
  };
  load_gx(0, gx);

  long long* tl = (a.tl && blockIdx.x == 0 && tid == 0) ? a.tl : nullptr;
  for (int s = 0; s < a.T; ++s) {
    const int t = dir ? a.T - 1 - s : s;
    uint8_t* hcur = Hs + (s & 1) * kHBuf;
    uint8_t* hnext = Hs + ((s + 1) & 1) * kHBuf;
    if (tl) tl[8 * s + 0] = clock64();
    if (tid == 0) {
      fence_proxy_async_all();
      tc_fence_after();
      const uint32_t wa = smem_u32(Ws), hb = smem_u32(hcur);
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        const uint64_t adesc = smem_desc_kmajor_sw128(wa + kc * kATile);
        const uint64_t bdesc = smem_desc_kmajor_sw128(hb + kc * kBTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
      }
      mma_commit(bar);
    }
    if (tl) tl[8 * s + 1] = clock64();
    float gx_next[8];
    if (s + 1 < a.T) load_gx(s + 1, gx_next);
    mbar_wait(bar, s & 1);
    tc_fence_after();
    if (tl) tl[8 * s + 2] = clock64();
    float pre[8];
    tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(bh * 8), pre);
    tmem_ld_wait();
    tc_fence_before();
    {
      float* g = a.gates + (((long long)t * a.B + b0 + bh * 8) * 2 + dir) * (4 * kH) + q * kH + rank * kUnits + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = pre[j] + gx[j];
        const float av = q == 2 ? tanhf(v) : sigmoidf_(v);
        act[(q * kBC + bh * 8 + j) * kUnits + lane] = av;
        if (b0 + bh * 8 + j < a.B) g[j * gstride] = av;   // saved for the backward pass
      }
    }
    __syncthreads();
    if (tl) tl[8 * s + 3] = clock64();
    float hv[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int bl = warp + 8 * p;
      const float ig = act[(0 * kBC + bl) * kUnits + lane], fg = act[(1 * kBC + bl) * kUnits + lane];
      const float gg = act[(2 * kBC + bl) * kUnits + lane], og = act[(3 * kBC + bl) * kUnits + lane];
      c[p] = fg * c[p] + ig * gg;
      const bool live = b0 + bl < a.B;
      hv[p] = live ? og * tanhf(c[p]) : 0.f;
      // the operand copy of h is rounded to tf32 here (round-to-nearest; the tensor core would truncate)
      *reinterpret_cast<float*>(hnext + sw128_off(bl, rank * kUnits + lane, kBTile)) = tf32_rn(hv[p]);
    }
    __syncthreads();
    if (tl) tl[8 * s + 4] = clock64();
    {  // this CTA's K chunk (16 batch rows x 128 B = 2 KB) -> the same place in the 7 peers, 16-byte stores
      const float4* src = reinterpret_cast<const float4*>(hnext + rank * kBTile);
      for (int i = tid; i < (kCluster - 1) * (kBTile / 16); i += kThreads) {
        int dst_rank = i / (kBTile / 16);
        const int e = i % (kBTile / 16);
        dst_rank += (dst_rank >= rank);
        float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(hnext + rank * kBTile, dst_rank));
        dst[e] = src[e];
      }
    }
    fence_proxy_async_all();
    if (tl) tl[8 * s + 5] = clock64();
    // arrive (release) right after the DSMEM stores; the global stores below are issued after it, so the barrier does
    // not wait for them
    cluster.barrier_arrive();
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int b = b0 + warp + 8 * p;
      if (b < a.B) {
        a.cells[(((long long)t * a.B + b) * 2 + dir) * kH + rank * kUnits + lane] = c[p];
        a.y[((long long)t * a.B + b) * (2 * kH) + dir * kH + rank * kUnits + lane] = hv[p];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) gx[j] = gx_next[j];
    if (tl) tl[8 * s + 6] = clock64();
    cluster.barrier_wait();
    if (tl) tl[8 * s + 7] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_bwd_kernel(LstmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Ws = smem;                    // A operand W_slice^T: [k half 2][gate 4] tiles of [128 k][32 gate rows = 128 B]
  uint8_t* Dg = smem + kSmemW;           // B operand d(gates): 4 K chunks (one per gate) x [16 batch][128 B]
  float* recv = reinterpret_cast<float*>(Dg + 4 * kBTile);  // [2][src CTA][batch][unit]
  float* stage = recv + 2 * kCluster * kBC * kUnits;          // [warp][batch][unit]: transposes the TMEM rows for the scatter
  uint64_t* bar = reinterpret_cast<uint64_t*>(stage + kCluster * kBC * kUnits);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  const int dir = cid & 1, b0 = (cid >> 1) * kBC;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // W_slice^T -> swizzled A operand (rows = k, K = own gate rows). Lanes = consecutive gate rows of one gate, so the four
  // scalar stores of a thread's float4 (four consecutive k = four A rows) are bank-conflict free.
  {
    const float* W = a.w_hh[dir];
    for (int i = tid; i < 4 * kUnits * (kH / 4); i += kThreads) {
      const int jl = i % kUnits, kq = (i / kUnits) % (kH / 4), gq = i / (kUnits * (kH / 4));
      const int grow = gq * kH + rank * kUnits + jl;
      const float4 v = __ldg(reinterpret_cast<const float4*>(W + (long long)grow * kH) + kq);
      const int k = kq * 4;
      uint8_t* tile = Ws + ((k >> 7) * 4 + gq) * kATile;
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) *reinterpret_cast<float*>(tile + sw128_off((k & 127) + e, jl, 0)) = tf32_rn(vv[e]);
    }
  }
  for (int i = tid; i < 2 * kCluster * kBC * kUnits; i += kThreads) recv[i] = 0.f;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  fence_proxy_async_all();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t idesc = instr_desc_tf32(128, kBC, 0, 0);
  float dc[2] = {0.f, 0.f};
  // per (unit = lane, batch row warp + 8p): saved gate activations, c_t, c_{t-1}, dy - loaded one step ahead
  struct Saved { float ig, fg, gg, og, ct, cp, dy; };
  auto load_saved = [&](int s, int p) {
    Saved v = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int t = dir ? s : a.T - 1 - s;          // reverse of the forward order
    const int tp = dir ? t + 1 : t - 1;           // step whose c is c_prev
    const int b = b0 + warp + 8 * p;
    if (b < a.B) {
      const float* g = a.gates + (((long long)t * a.B + b) * 2 + dir) * (4 * kH) + rank * kUnits + lane;
      v.ig = g[0]; v.fg = g[kH]; v.gg = g[2 * kH]; v.og = g[3 * kH];
      v.ct = a.cells[(((long long)t * a.B + b) * 2 + dir) * kH + rank * kUnits + lane];
      v.cp = (tp >= 0 && tp < a.T) ? a.cells[(((long long)tp * a.B + b) * 2 + dir) * kH + rank * kUnits + lane] : 0.f;
      v.dy = a.dy[((long long)t * a.B + b) * (2 * kH) + dir * kH + rank * kUnits + lane];
    }
    return v;
  };
  Saved sv[2] = {load_saved(0, 0), load_saved(0, 1)};

  long long* tl = (a.tl && blockIdx.x == 0 && tid == 0) ? a.tl : nullptr;
  for (int s = 0; s < a.T; ++s) {
    const int t = dir ? s : a.T - 1 - s;
    const float* rcv = recv + ((s + 1) & 1) * kCluster * kBC * kUnits;  // written during step s-1 (zeros at s = 0)
    if (tl) tl[8 * s + 0] = clock64();
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const int bl = warp + 8 * p, b = b0 + bl;
      float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f;
      if (b < a.B) {
        float dh = sv[p].dy;
#pragma unroll
        for (int r = 0; r < kCluster; ++r) dh += rcv[(r * kBC + bl) * kUnits + lane];
        const float ig = sv[p].ig, fg = sv[p].fg, gg = sv[p].gg, og = sv[p].og;
        const float tcv = tanhf(sv[p].ct);
        d_o = dh * tcv * og * (1.f - og);
        const float dct = dc[p] + dh * og * (1.f - tcv * tcv);
        d_i = dct * gg * ig * (1.f - ig);
        d_g = dct * ig * (1.f - gg * gg);
        d_f = dct * sv[p].cp * fg * (1.f - fg);
        dc[p] = dct * fg;
        float* g = a.gates + (((long long)t * a.B + b) * 2 + dir) * (4 * kH) + rank * kUnits + lane;
        g[0] = d_i; g[kH] = d_f; g[2 * kH] = d_g; g[3 * kH] = d_o;   // d(pre-activations) for the weight-gradient GEMMs
      }
      const uint32_t off = sw128_off(bl, lane, 0);
      // operand copies rounded to tf32 (round-to-nearest; the tensor core would truncate)
      *reinterpret_cast<float*>(Dg + 0 * kBTile + off) = tf32_rn(d_i);
      *reinterpret_cast<float*>(Dg + 1 * kBTile + off) = tf32_rn(d_f);
      *reinterpret_cast<float*>(Dg + 2 * kBTile + off) = tf32_rn(d_g);
      *reinterpret_cast<float*>(Dg + 3 * kBTile + off) = tf32_rn(d_o);
    }
    fence_proxy_async_all();
    __syncthreads();
    if (tl) tl[8 * s + 1] = clock64();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t wa = smem_u32(Ws), db = smem_u32(Dg);
#pragma unroll
      for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const uint64_t adesc = smem_desc_kmajor_sw128(wa + (mh * 4 + gq) * kATile);
          const uint64_t bdesc = smem_desc_kmajor_sw128(db + gq * kBTile);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_tf32_ss(tmem_base + mh * kBC, adesc + 2 * k, bdesc + 2 * k, idesc, (gq | k) != 0);
        }
      }
      mma_commit(bar);
    }
    if (s + 1 < a.T) {
      sv[0] = load_saved(s + 1, 0);
      sv[1] = load_saved(s + 1, 1);
    }
    if (tl) tl[8 * s + 2] = clock64();
    mbar_wait(bar, s & 1);
    tc_fence_after();
    if (tl) tl[8 * s + 3] = clock64();
    // warp w reads k = 128 (w / 4) + 32 (w % 4) + lane for the 16 batch rows: exactly the k range of CTA dst = w. A thread
    // owns one k (one TMEM lane); the block is transposed through shared memory so that the scatter is 16-byte stores of
    // four consecutive units.
    float v[kBC];
    tmem_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * kBC), v);
    tmem_ld_wait();
    tc_fence_before();
    {
      float* stg = stage + warp * kBC * kUnits;
#pragma unroll
      for (int j = 0; j < kBC; ++j) stg[j * kUnits + lane] = v[j];
      __syncwarp();
      float4* rnext = reinterpret_cast<float4*>(cluster.map_shared_rank(recv + (s & 1) * kCluster * kBC * kUnits, warp) + rank * kBC * kUnits);
#pragma unroll
      for (int i = 0; i < kBC * kUnits / 4 / 32; ++i) rnext[i * 32 + lane] = reinterpret_cast<const float4*>(stg)[i * 32 + lane];
    }
    if (tl) tl[8 * s + 4] = clock64();
    cluster.barrier_arrive();
    if (tl) tl[8 * s + 5] = clock64();
    cluster.barrier_wait();
    if (tl) tl[8 * s + 6] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

int launch_cluster(const void* fn, size_t smem, int n_clusters, LstmArgs& args, cudaStream_t st) {
  args.tl = qeb_debug_timeline();
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
