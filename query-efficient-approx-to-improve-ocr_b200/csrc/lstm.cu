// Bidirectional LSTM recurrence (hidden size 256) as persistent thread-block-cluster kernels on tcgen05.
// Reference: nn.LSTM(512, 256, 2, bidirectional=True), models/model_crnn.py:9,19 (gate order i,f,g,o; h0 = c0 = 0).
//
// The input projections x*W_ih^T + b_ih + b_hh are tensor-core GEMMs (conv_tc.cu); this file runs the T dependent
// steps. One cluster of 8 CTAs serves one (direction, chunk of kBC = 16 batch rows); CTA r of the cluster owns hidden
// units [32r, 32r+32), i.e. 128 of the 1024 gate rows. Its W_hh slice (128 x 256) is loaded ONCE into TENSOR MEMORY
// (128 lanes x 256 columns, rounded to tf32) and stays there for the whole sequence as the A operand of
// tcgen05.mma.kind::tf32 (A from TMEM, B from shared memory): with N = 16 an MMA whose A operand comes from shared memory
// is bound by reading the 4 KB A tile (measured 89 cycles per MMA, 2.9 k cycles per step), from TMEM it is not.
//
// Forward step: D[128 gate rows][16 batch] = W_slice * h_{t-1}^T is 32 MMAs (K = 8) issued by one thread. The operand
// copy of h is rounded to tf32 with round-to-nearest (the tensor core itself would truncate); y, c and the saved gate
// activations stay fp32. Warp q of each half-block reads gate q of its 32 units for 8 batch rows from TMEM (tcgen05.ld),
// adds the x-projection, applies the non-linearity and passes the activated gates through 8 KB of shared memory to the
// cell threads, which update c and write the 32 new h values per batch row into the B-operand layout; CTA r owns exactly
// K chunk r of that operand (2 KB) and pushes it into the other 7 CTAs with cp.async.bulk (shared::cta ->
// shared::cluster), which signals the receiver's mbarrier with the byte count. There is no cluster barrier in the loop:
// the operand is double-buffered and a CTA can only run one step ahead of its slowest peer because it needs that peer's
// chunk for its next MMA.
//
// Backward step: the cell threads form d(gates) for their units and write them as the B operand [16 batch][128 own gate
// rows]; dh_{t-1}[k][b] partial = W_slice^T (256 x 128, two M = 128 halves in TMEM) * dgates is 2 x 16 MMAs; warp w reads
// the k range that belongs to CTA w from TMEM, transposes it through shared memory and pushes the 2 KB block to CTA w with
// one bulk copy (a reduce-scatter): CTA dst receives its k range from all 8 CTAs and sums the contributions when it forms
// dh in the next step.
#include "nn.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
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

constexpr int kBTile = kBC * 128;        // 2 KB: one 32-element (128-byte) K chunk of a B operand, [16 batch][128 B]
constexpr int kHBuf = 8 * kBTile;        // 16 KB: h for all 256 k
constexpr int kBlk = kBC * kUnits;       // floats in one [batch][unit] block (2 KB)
// TMEM: columns [0,256) hold the A operand, the accumulator follows. 512 columns = the whole tensor memory of the SM; the
// shared-memory request is padded so that no other CTA can become resident next to a recurrence CTA and wait for TMEM.
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol = 256;
constexpr int kChains = 1;   // independent accumulators the K loop of a step can be spread over (measured: no gain, issue-bound before, 20 cycles per MMA now)
constexpr size_t kSmemPad = 200 * 1024;

__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return __fdividef(2.f, 1.f + __expf(-2.f * x)) - 1.f; }
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
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
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
// this thread's TMEM lane, 16 consecutive columns <- 16 registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem], kind::tf32, issued by one thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_saddr), "r"(rank));
  return r;
}
// bulk copy own shared memory -> shared memory of a CTA of the cluster; completes `bytes` on the destination's mbarrier
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}

struct LstmArgs {
  float* gates;         // (T,B,2,1024)
  const float* w_hh[2]; // (1024,256) per direction
  float* cells;         // (T,B,2,256)
  float* y;             // (T,B,512)
  __half* y16;          // optional fp16 shadow of y (operand of the next projection GEMM), forward only
  const float* dy;      // (T,B,512), backward only
  int T, B;
  int round_io;         // 1 (the network engines): y (forward) / d(pre-activations) (backward) are stored rounded to tf32 - they are
                        // operands of the projection / weight-gradient contractions that follow (common.cuh qeb_tf32r)
  long long* tl;        // debugging aid (qeb_debug_set_timeline): clock64 stamps of CTA 0, 8 per step
};

// RB: batch rows a cluster really serves (16, or 8: the MMA keeps N = kBC = 16 - M = 128 allows no narrower tile - with the
// upper 8 operand rows zero; the gate / cell / store work per thread and the pushed bytes halve and twice as many clusters
// run: 128 instead of 64 CTAs at B = 64. Chosen by the launcher when all clusters still fit one wave.)
template <int RB>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_fwd_kernel(LstmArgs a) {
  constexpr int RPT = RB / 2;    // batch rows per gate thread
  constexpr int CPT = RB / 8;    // batch rows per cell thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Hs = smem;                                     // B operand, double-buffered: [2][8 K chunks][16 batch][128 B]
  float* act = reinterpret_cast<float*>(Hs + 2 * kHBuf);  // activated gates [q][b][unit]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(act + 4 * kBlk);
  uint64_t* full_bar = mma_bar + 1;                       // [2]: the peers' chunks of h buffer b have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 2);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  const int dir = cid & 1, b0 = (cid >> 1) * RB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, bh = warp >> 2;

  for (int i = tid; i < 2 * kHBuf / 16; i += kThreads) reinterpret_cast<float4*>(Hs)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  qeb_pdl_sync();
  {
    // W_hh slice -> TMEM, rounded to tf32: lane = gate row (gate q, unit `lane`), column = k. Warps w and w + 4 share a
    // lane quarter and take 128 columns each; a thread streams 512 contiguous bytes of its row.
    const float* W = (dir ? a.w_hh[1] : a.w_hh[0]) + (long long)(q * kH + rank * kUnits + lane) * kH + bh * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + c0) + j);
        v[4 * j] = tf32_rn(w4.x); v[4 * j + 1] = tf32_rn(w4.y); v[4 * j + 2] = tf32_rn(w4.z); v[4 * j + 3] = tf32_rn(w4.w);
      }
      tmem_st16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(bh * 128 + c0), v);
    }
    tmem_st_wait();
  }
  fence_proxy_async_all();
  tc_fence_before();
  cluster.sync();
  tc_fence_after();

  // roles. gate thread: gate q of unit `lane` for batch rows [RPT bh, RPT bh + RPT); cell thread: unit `lane`, batch rows warp (, warp + 8)
  float c[2] = {0.f, 0.f};
  constexpr uint32_t idesc = instr_desc_tf32(128, kBC, 0, 0);
  const long long gstride = 2 * 4 * kH;  // floats between consecutive batch rows of `gates`

  // x-projections of this thread's gate row for its 8 batch rows, loaded one step ahead
  float gx[RPT];
  auto load_gx = [&](int s, float (&dst)[RPT]) {   // array by reference: a pointer parameter would force gx into local memory
    const int t = dir ? a.T - 1 - s : s;
    const float* g = a.gates + (((long long)t * a.B + b0 + bh * RPT) * 2 + dir) * (4 * kH) + q * kH + rank * kUnits + lane;
#pragma unroll
    for (int j = 0; j < RPT; ++j) dst[j] = (b0 + bh * RPT + j < a.B) ? g[j * gstride] : 0.f;
  };
  load_gx(0, gx);

  long long* tl = (a.tl && blockIdx.x == 0 && tid == 0) ? a.tl : nullptr;
  for (int s = 0; s < a.T; ++s) {
    const int t = dir ? a.T - 1 - s : s;
    uint8_t* hcur = Hs + (s & 1) * kHBuf;
    uint8_t* hnext = Hs + ((s + 1) & 1) * kHBuf;
    if (tl) tl[8 * s + 0] = clock64();
    if (warp == 0) {   // warp-uniform branch + elect.sync: descriptors stay in uniform registers, no per-MMA lane loop
      if (s > 0) mbar_wait(&full_bar[s & 1], ((s - 1) >> 1) & 1);   // the 7 remote chunks of h_s (own chunk: see below)
      if (tl) tl[8 * s + 1] = clock64();
      tc_fence_after();
      if (elect_one()) {
        const uint32_t hb = smem_u32(hcur);
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          const uint64_t bdesc = smem_desc_kmajor_sw128(hb + kc * kBTile);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // chain k accumulates K steps k, k + 4, ...: kChains independent accumulators
            mma_tf32_ts(tmem_base + kAccCol + (uint32_t)((k % kChains) * kBC), tmem_base + (uint32_t)(kc * 32 + k * 8), bdesc + 2 * k,
                        idesc, (kc * 4 + k) >= kChains);
        }
        mma_commit(mma_bar);
      }
      __syncwarp();
    }
    float gx_next[RPT];
    if (s + 1 < a.T) load_gx(s + 1, gx_next);
    mbar_wait(mma_bar, s & 1);
    tc_fence_after();
    if (tl) tl[8 * s + 2] = clock64();
    float pre[kChains][RPT];
#pragma unroll
    for (int ch = 0; ch < kChains; ++ch) {
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + kAccCol + (uint32_t)(ch * kBC + bh * RPT);
      if constexpr (RPT == 8) tmem_ld8(ta, pre[ch]);
      else tmem_ld4(ta, pre[ch]);
    }
    tmem_ld_wait();
    tc_fence_before();
    float av[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      float v = gx[j];
#pragma unroll
      for (int ch = 0; ch < kChains; ++ch) v += pre[ch][j];
      av[j] = q == 2 ? fast_tanh(v) : fast_sigmoid(v);
      act[(q * kBC + bh * RPT + j) * kUnits + lane] = av[j];
    }
    __syncthreads();
    if (tl) tl[8 * s + 3] = clock64();
    float hv[2] = {0.f, 0.f};
#pragma unroll
    for (int p = 0; p < CPT; ++p) {
      const int bl = warp + 8 * p;
      const float ig = act[(0 * kBC + bl) * kUnits + lane], fg = act[(1 * kBC + bl) * kUnits + lane];
      const float gg = act[(2 * kBC + bl) * kUnits + lane], og = act[(3 * kBC + bl) * kUnits + lane];
      c[p] = fg * c[p] + ig * gg;
      const bool live = b0 + bl < a.B;
      hv[p] = live ? og * fast_tanh(c[p]) : 0.f;
      // the operand copy of h is rounded to tf32 here (round-to-nearest; the tensor core would truncate)
      *reinterpret_cast<float*>(hnext + sw128_off(bl, rank * kUnits + lane, kBTile)) = tf32_rn(hv[p]);
    }
    fence_proxy_async_all();   // own chunk: generic-proxy writes -> visible to the bulk copy and to the next step's MMA
    __syncthreads();
    if (tl) tl[8 * s + 4] = clock64();
    if (warp == 0 && s + 1 < a.T) {   // warp-uniform branch + elected lane: the copy instructions take uniform operands
      if (elect_one()) {
        // push this CTA's K chunk (RB batch rows x 128 B: the first RB rows of the tile) into the same place of the 7 peers;
        // arm the own barrier for theirs
        constexpr uint32_t kPush = RB * 128;
        mbar_expect_tx(&full_bar[(s + 1) & 1], (kCluster - 1) * kPush);
        const uint32_t src = smem_u32(hnext + rank * kBTile), bar = smem_u32(&full_bar[(s + 1) & 1]);
#pragma unroll
        for (int r = 1; r < kCluster; ++r) {
          const uint32_t peer = (uint32_t)((rank + r) & (kCluster - 1));
          bulk_copy_to_cluster(mapa_u32(src, peer), src, kPush, mapa_u32(bar, peer));
        }
      }
      __syncwarp();
    }
    if (tl) tl[8 * s + 5] = clock64();
    // global stores last: a proxy fence waits for the thread's outstanding stores, these drain during the next step
    {
      float* g = a.gates + (((long long)t * a.B + b0 + bh * RPT) * 2 + dir) * (4 * kH) + q * kH + rank * kUnits + lane;
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        if (b0 + bh * RPT + j < a.B) g[j * gstride] = av[j];   // activated gates, saved for the backward pass
    }
#pragma unroll
    for (int p = 0; p < CPT; ++p) {
      const int b = b0 + warp + 8 * p;
      if (b < a.B) {
        a.cells[(((long long)t * a.B + b) * 2 + dir) * kH + rank * kUnits + lane] = c[p];
        a.y[((long long)t * a.B + b) * (2 * kH) + dir * kH + rank * kUnits + lane] = a.round_io ? tf32_rn(hv[p]) : hv[p];
        if (a.y16) a.y16[((long long)t * a.B + b) * (2 * kH) + dir * kH + rank * kUnits + lane] = __float2half_rn(hv[p]);
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) gx[j] = gx_next[j];
    if (tl) tl[8 * s + 6] = clock64();
  }
  tc_fence_before();
  cluster.sync();   // no CTA leaves while a peer may still read the chunk it pushed or write into this CTA
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

template <int RB>   // see lstm_fwd_kernel
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) lstm_bwd_kernel(LstmArgs a) {
  constexpr int CPT = RB / 8;    // batch rows per cell thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Dg = smem;                                        // B operand d(gates): 4 K chunks (one per gate) x [16 batch][128 B]
  float* recv = reinterpret_cast<float*>(Dg + 4 * kBTile);   // [2][src CTA][batch][unit]
  float* stage = recv + 2 * kCluster * kBlk;                 // [2][dst CTA = warp][batch][unit]: sources of the bulk copies
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(stage + 2 * kCluster * kBlk);
  uint64_t* full_bar = mma_bar + 1;                          // [2]: all 8 blocks of recv buffer b have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 2);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  const int dir = cid & 1, b0 = (cid >> 1) * RB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (RB < kBC)   // operand rows [RB, 16) are never written: they must read as zero
    for (int i = tid; i < 4 * kBTile / 16; i += kThreads) reinterpret_cast<float4*>(Dg)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_init(&full_bar[0], 1);
    mbar_init(&full_bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  qeb_pdl_sync();
  {
    // W_slice^T -> TMEM, rounded to tf32: half mh = warp / 4 of k lives in columns [128 mh, 128 mh + 128); lane = k % 128,
    // column = own gate row (gate gq, unit jl). For a fixed gate row the 32 lanes of a warp read 32 consecutive k.
    const int mh = warp >> 2;
    const float* W = (dir ? a.w_hh[1] : a.w_hh[0]) + (long long)(rank * kUnits) * kH + mh * 128 + (warp & 3) * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 16) {   // c0 = gq * 32 + jl
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tf32_rn(__ldg(W + (long long)((c0 >> 5) * kH + (c0 & 31) + j) * kH));
      tmem_st16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mh * 128 + c0), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();

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
  Saved sv[2] = {load_saved(0, 0), CPT > 1 ? load_saved(0, 1) : Saved{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}};

  long long* tl = (a.tl && blockIdx.x == 0 && tid == 0) ? a.tl : nullptr;
  for (int s = 0; s < a.T; ++s) {
    const int t = dir ? s : a.T - 1 - s;
    const float* rcv = recv + ((s + 1) & 1) * kCluster * kBlk;  // pushed by the 8 CTAs during step s-1
    if (tl) tl[8 * s + 0] = clock64();
    if (s > 0) mbar_wait(&full_bar[(s + 1) & 1], ((s - 1) >> 1) & 1);
    if (tl) tl[8 * s + 1] = clock64();
    float dgv[2][4];
#pragma unroll
    for (int p = 0; p < CPT; ++p) {
      const int bl = warp + 8 * p, b = b0 + bl;
      float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f;
      if (b < a.B) {
        float dh = sv[p].dy;
        if (s > 0) {
#pragma unroll
          for (int r = 0; r < kCluster; ++r) dh += rcv[(r * kBC + bl) * kUnits + lane];
        }
        const float ig = sv[p].ig, fg = sv[p].fg, gg = sv[p].gg, og = sv[p].og;
        const float tcv = fast_tanh(sv[p].ct);
        d_o = dh * tcv * og * (1.f - og);
        const float dct = dc[p] + dh * og * (1.f - tcv * tcv);
        d_i = dct * gg * ig * (1.f - ig);
        d_g = dct * ig * (1.f - gg * gg);
        d_f = dct * sv[p].cp * fg * (1.f - fg);
        dc[p] = dct * fg;
      }
      dgv[p][0] = d_i; dgv[p][1] = d_f; dgv[p][2] = d_g; dgv[p][3] = d_o;
      const uint32_t off = sw128_off(bl, lane, 0);
      // operand copies rounded to tf32 (round-to-nearest; the tensor core would truncate)
      *reinterpret_cast<float*>(Dg + 0 * kBTile + off) = tf32_rn(d_i);
      *reinterpret_cast<float*>(Dg + 1 * kBTile + off) = tf32_rn(d_f);
      *reinterpret_cast<float*>(Dg + 2 * kBTile + off) = tf32_rn(d_g);
      *reinterpret_cast<float*>(Dg + 3 * kBTile + off) = tf32_rn(d_o);
    }
    auto store_dgates = [&]() {   // d(pre-activations) for the weight-gradient GEMMs; after the proxy fence (see forward)
#pragma unroll
      for (int p = 0; p < CPT; ++p) {
        const int b = b0 + warp + 8 * p;
        if (b < a.B) {
          float* g = a.gates + (((long long)t * a.B + b) * 2 + dir) * (4 * kH) + rank * kUnits + lane;
          // operands of the W_ih / W_hh weight-gradient and the d(input) contractions: rounded to tf32 here
          if (a.round_io) {
            g[0] = tf32_rn(dgv[p][0]); g[kH] = tf32_rn(dgv[p][1]); g[2 * kH] = tf32_rn(dgv[p][2]); g[3 * kH] = tf32_rn(dgv[p][3]);
          } else {
            g[0] = dgv[p][0]; g[kH] = dgv[p][1]; g[2 * kH] = dgv[p][2]; g[3 * kH] = dgv[p][3];
          }
        }
      }
    };
    if (s + 1 == a.T) {   // dh before the first step is not needed
      store_dgates();
      break;
    }
    fence_proxy_async_all();
    __syncthreads();
    if (tl) tl[8 * s + 2] = clock64();
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(&full_bar[s & 1], kCluster * RB * kUnits * sizeof(float));   // the 8 blocks the cluster pushes in this step
      tc_fence_after();
      if (elect_one()) {
        const uint32_t db = smem_u32(Dg);
#pragma unroll
        for (int mh = 0; mh < 2; ++mh) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const uint64_t bdesc = smem_desc_kmajor_sw128(db + gq * kBTile);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_tf32_ts(tmem_base + kAccCol + (uint32_t)((mh * kChains + k % kChains) * kBC), tmem_base + (uint32_t)(mh * 128 + gq * 32 + k * 8),
                          bdesc + 2 * k, idesc, (gq * 4 + k) >= kChains);
          }
        }
        mma_commit(mma_bar);
      }
      __syncwarp();
    }
    store_dgates();
    sv[0] = load_saved(s + 1, 0);
    if (CPT > 1) sv[1] = load_saved(s + 1, 1);
    mbar_wait(mma_bar, s & 1);
    tc_fence_after();
    if (tl) tl[8 * s + 3] = clock64();
    // warp w reads k = 128 (w / 4) + 32 (w % 4) + lane for the 16 batch rows: exactly the k range of CTA dst = w. A thread
    // owns one k (one TMEM lane); the block is transposed through shared memory into the receiver's [batch][unit] layout and
    // pushed with one 2 KB bulk copy. The staging buffer is double-buffered: the copy of step s-2 has been consumed by the
    // time step s was allowed to start, the one of step s-1 possibly not.
    float v[RB];
    {
      float vc[kChains][RB];
#pragma unroll
      for (int ch = 0; ch < kChains; ++ch) {
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kAccCol + (uint32_t)(((warp >> 2) * kChains + ch) * kBC);
        if constexpr (RB == 16) tmem_ld16(ta, vc[ch]);
        else tmem_ld8(ta, vc[ch]);
      }
      tmem_ld_wait();
      tc_fence_before();
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        v[j] = vc[0][j];
#pragma unroll
        for (int ch = 1; ch < kChains; ++ch) v[j] += vc[ch][j];
      }
    }
    {
      float* stg = stage + ((s & 1) * kCluster + warp) * kBlk;
#pragma unroll
      for (int j = 0; j < RB; ++j) stg[j * kUnits + lane] = v[j];
      fence_proxy_async_all();
      __syncwarp();
      if (elect_one()) {
        const uint32_t dst = smem_u32(recv + ((s & 1) * kCluster + rank) * kBlk);
        bulk_copy_to_cluster(mapa_u32(dst, warp), smem_u32(stg), RB * kUnits * sizeof(float), mapa_u32(smem_u32(&full_bar[s & 1]), warp));
      }
    }
    if (tl) tl[8 * s + 4] = clock64();
  }
  tc_fence_before();
  cluster.sync();   // no CTA leaves while a peer may still read the block it pushed or write into this CTA
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// Batch rows per cluster: 16. The 8-row form (twice the clusters: 128 instead of 64 CTAs at B = 64, half the gate / cell /
// store work per thread and step) is correct (tests) but SLOWER in the step - 3.49 -> 3.66 ms (phase B), 12.4 -> 13.8 ms (jitter
// step): a step of the recurrence is a latency chain (MMA -> tcgen05.ld -> gates -> barrier -> cell -> proxy fence -> barrier ->
// push -> the peers' wait), not per-thread arithmetic, and twice as many whole-SM CTAs take the SMs away from the weight-
// gradient kernels that run beside it. Kept behind QEB_LSTM_ROWS=8 for measurements.
int rows_per_cluster(int B) {
  static const int forced = getenv("QEB_LSTM_ROWS") ? atoi(getenv("QEB_LSTM_ROWS")) : 0;
  (void)B;
  return forced == 8 ? 8 : 16;
}

int launch_cluster(const void* fn, size_t smem, int n_clusters, LstmArgs& args, cudaStream_t st) {
  args.tl = qeb_debug_timeline();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_clusters * kCluster);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = qeb_pdl_enabled() ? 1 : 0;
  void* kargs[] = {&args};
  if (qeb_dbg_skip_active() && qeb_skip_flag()) return QEB_OK;
  QEB_CUDA(cudaLaunchKernelExC(&cfg, fn, kargs));
  qeb_count_launch();
  return QEB_OK;
}

}  // namespace

int lstm_layer_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* cells, float* y, int T, int B,
                   cudaStream_t st, void* y16, int round_io) {
  QEB_REQUIRE(gates && w_hh_fwd && w_hh_rev && cells && y && T > 0 && B > 0, "lstm_layer_fwd: bad arguments");
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPad));
    QEB_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPad));
    attr = true;
  }
  LstmArgs a;
  a.gates = gates; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_rev; a.cells = cells; a.y = y; a.dy = nullptr; a.T = T; a.B = B;
  a.y16 = static_cast<__half*>(y16);
  a.round_io = round_io;
  ProfScope prof("lstm_fwd", st, 2.0 * T * B * 2 * 1024 * 256, 4.0 * T * B * (2 * 2048 + 512 + 512));
  if (rows_per_cluster(B) == 8) return launch_cluster((const void*)lstm_fwd_kernel<8>, kSmemPad, 2 * qeb_cdiv(B, 8), a, st);
  return launch_cluster((const void*)lstm_fwd_kernel<16>, kSmemPad, 2 * qeb_cdiv(B, kBC), a, st);
}

int lstm_layer_bwd(float* gates, const float* cells, const float* dy, const float* w_hh_fwd, const float* w_hh_rev, int T,
                   int B, cudaStream_t st, int round_io) {
  QEB_REQUIRE(gates && w_hh_fwd && w_hh_rev && cells && dy && T > 0 && B > 0, "lstm_layer_bwd: bad arguments");
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPad));
    QEB_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPad));
    attr = true;
  }
  LstmArgs a;
  a.gates = gates; a.w_hh[0] = w_hh_fwd; a.w_hh[1] = w_hh_rev; a.cells = const_cast<float*>(cells); a.y = nullptr; a.dy = dy;
  a.y16 = nullptr;
  a.T = T; a.B = B;
  a.round_io = round_io;
  ProfScope prof("lstm_bwd", st, 2.0 * T * B * 2 * 1024 * 256, 4.0 * T * B * (2 * 2048 + 512 + 512));
  if (rows_per_cluster(B) == 8) return launch_cluster((const void*)lstm_bwd_kernel<8>, kSmemPad, 2 * qeb_cdiv(B, 8), a, st);
  return launch_cluster((const void*)lstm_bwd_kernel<16>, kSmemPad, 2 * qeb_cdiv(B, kBC), a, st);
}

// C ABI (tests): one bidirectional layer of the recurrence
QEB_API int qeb_lstm_layer_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* cells, float* y, int T, int B,
                               void* stream) {
  return lstm_layer_fwd(gates, w_hh_fwd, w_hh_rev, cells, y, T, B, (cudaStream_t)stream, nullptr, 0);
}
QEB_API int qeb_lstm_layer_bwd(float* gates, const float* cells, const float* dy, const float* w_hh_fwd, const float* w_hh_rev,
                               int T, int B, void* stream) {
  return lstm_layer_bwd(gates, cells, dy, w_hh_fwd, w_hh_rev, T, B, (cudaStream_t)stream, 0);
}
