// Implicit-GEMM convolution / GEMM on tcgen05 (kind::tf32, fp32 accumulate in TMEM), operands staged by TMA.
//
// One kernel family serves every dense contraction of the hot path whose operands are K-major:
//   conv fprop    models/model_crnn.py:47-56 (conv2..conv7), models/model_unet.py:_block (all but the 1-channel conv)
//   conv dgrad    the same kernel on dY with the spatially flipped, channel-transposed weights
//   GEMM          LSTM input projections (nn.LSTM, models/model_crnn.py:9,19), Linear (models/model_crnn.py:10,20),
//                 ConvTranspose2d 2x2 s2 (models/model_unet.py:25-44) as a per-pixel GEMM with a pixel-shuffle store
//
// Data layout: activations NHWC fp32 in HBM (channels contiguous, arbitrary channel stride so that concat buffers
// are read/written in place); weights packed [Cout][tap][Cin] (K-major). A CTA computes a 128-pixel x BLOCK_N
// output tile: for every filter tap and 32-channel slice one 4-D TMA box {32 ch, Wt, Ht, Nt} (Wt*Ht*Nt = 128
// output pixels, shifted by the tap, out-of-bounds = zero padding) lands in shared memory as the K-major
// 128B-swizzled A tile, one 2-D box {32, BLOCK_N} as the B tile; four tcgen05.mma (K=8 each) consume the stage.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue (tcgen05.ld -> bias/ReLU ->
// global). Pipelines: smem full/empty mbarriers (kStages deep), one TMEM-full barrier.
#include "tc_common.cuh"

namespace tc {

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_32b_atom) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    QEB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      qeb_set_error("cuTensorMapEncodeTiled not available from the driver");
      return QEB_ERR_CUDA;
    }
    encode = (EncodeFn)fn;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[4];
  cuuint32_t b[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  if (((uintptr_t)base & 15) != 0) {
    qeb_set_error("tensor map base %p not 16-byte aligned", base);
    return QEB_ERR_INVALID;
  }
  CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_32b_atom ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    qeb_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
                  (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), (unsigned long long)(rank > 2 ? d[2] : 0),
                  (unsigned long long)(rank > 3 ? d[3] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
    return QEB_ERR_CUDA;
  }
  return QEB_OK;
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                     // fp32 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 4;  // 16 KB
constexpr int kThreads = 192;

struct FpropParams {
  int n_img, h_out, w_out;
  int wt, ht, nt;  // tile shape in output pixels, wt*ht*nt == 128
  int tiles_w, tiles_h;
  int kh, kw, ph, pw;
  int cin, kchunks;
  int n_total;        // GEMM N (all output channels)
  const float* bias;  // per GEMM-N column, nullable
  int relu;
  float* out;
  long long out_pix_stride;  // elements between consecutive output pixels
  int mode;                  // 0: plain NHWC store; 1: ConvTranspose 2x2 s2 pixel shuffle (N index = (dh*2+dw)*up_c + co)
  int up_c;
  int accumulate;            // 1: out += result (used when a gradient already holds a partial sum)
};

template <int BLOCK_N>
struct FpropCfg {
  static constexpr int kBBytes = BLOCK_N * kBlockK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const FpropParams p) {
  using Cfg = FpropCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
  const int w0 = tw * p.wt, h0 = th * p.ht, n0 = tn * p.nt;
  const int num_kb = p.kh * p.kw * p.kchunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tap = 0; tap < p.kh * p.kw; ++tap) {
        const int dy = tap / p.kw - p.ph, dx = tap % p.kw - p.pw;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_4d(sa, &tmap_a, &full_bar[stage], kc * kBlockK, w0 + dx, h0 + dy, n0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], tap * p.cin + kc * kBlockK, tile_n * BLOCK_N);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_tf32(kBlockM, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint64_t adesc = smem_desc_kmajor_sw128(sa);
        const uint64_t bdesc = smem_desc_kmajor_sw128(sa + kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 8; ++k) {
          // advance 8 tf32 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          mma_tf32_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        }
        mma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      mma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = TMEM lane
    const int ww = r % p.wt, hh = (r / p.wt) % p.ht, nn = r / (p.wt * p.ht);
    const int w = w0 + ww, h = h0 + hh, n = n0 + nn;
    const bool valid = (w < p.w_out) && (h < p.h_out) && (n < p.n_img);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      const int ncol = tile_n * BLOCK_N + c0;  // first GEMM-N column of this chunk
      if (valid && ncol < p.n_total) {
        float* dst;
        if (p.mode == 0) {
          dst = p.out + ((long long)(n * p.h_out + h) * p.w_out + w) * p.out_pix_stride + ncol;
        } else {
          const int sub = ncol / p.up_c, co = ncol - sub * p.up_c;
          const int oh = 2 * h + (sub >> 1), ow = 2 * w + (sub & 1);
          dst = p.out + ((long long)(n * 2 * p.h_out + oh) * (2 * p.w_out) + ow) * p.out_pix_stride + co;
        }
        const int lim = min(32, p.n_total - ncol);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (j < lim) ? __ldg(p.bias + ncol + j) : 0.f;
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (lim == 32) {
          float4* d4 = reinterpret_cast<float4*>(dst);
          if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o = d4[j];
              d4[j] = make_float4(o.x + v[4 * j], o.y + v[4 * j + 1], o.z + v[4 * j + 2], o.w + v[4 * j + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          for (int j = 0; j < lim; ++j) dst[j] = p.accumulate ? dst[j] + v[j] : v[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

template <int BLOCK_N>
int launch_fprop(const CUtensorMap& ta, const CUtensorMap& tb, const FpropParams& p, int m_tiles, int n_tiles,
                 cudaStream_t st) {
  using Cfg = FpropCfg<BLOCK_N>;
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr = true;
  }
  conv_fprop_tc_kernel<BLOCK_N><<<dim3(m_tiles, n_tiles), kThreads, Cfg::kSmemBytes, st>>>(ta, tb, p);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace

// Convolution (stride 1) / GEMM, K-major operands.
//   x: NHWC fp32 activations, n_img x h_in x w_in pixels, `cin` channels read at channel stride x_cstride
//      (x already points at the first channel of the slice). cin % 32 == 0.
//   wpacked: [n_total][kh*kw*cin] fp32 (K-major rows; a torch Linear weight is already in this form for kh=kw=1).
//   out: NHWC fp32, n_img x h_out x w_out pixels, channel stride out_cstride, n_total channels written
//        (mode 1: ConvTranspose2d 2x2 s2 pixel shuffle, n_total = 4*up_c, output image 2h_out x 2w_out, up_c channels).
//   A plain GEMM C[M,N] = A[M,K] B[N,K]^T is n_img=1, h=1, w=M, cin=K, kh=kw=1.
//   block_n: 0 = choose; else one of 32/64/128/256.
QEB_API int qeb_conv_fprop_tc(const float* x, int n_img, int h_in, int w_in, int cin, int x_cstride, const float* wpacked,
                              int n_total, int kh, int kw, int ph, int pw, int h_out, int w_out, const float* bias,
                              int relu, float* out, int out_cstride, int mode, int up_c, int accumulate, int block_n,
                              void* stream) {
  QEB_REQUIRE(x && wpacked && out, "conv_fprop_tc: null pointer");
  QEB_REQUIRE(n_img > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0, "conv_fprop_tc: bad spatial dims");
  QEB_REQUIRE(cin > 0 && cin % 32 == 0, "conv_fprop_tc: cin=%d must be a positive multiple of 32", cin);
  QEB_REQUIRE(x_cstride >= cin && x_cstride % 4 == 0 && out_cstride % 4 == 0, "conv_fprop_tc: channel strides must be multiples of 4");
  QEB_REQUIRE(n_total > 0 && kh > 0 && kw > 0, "conv_fprop_tc: bad filter dims");
  QEB_REQUIRE(((uintptr_t)out & 15) == 0, "conv_fprop_tc: out must be 16-byte aligned");
  QEB_REQUIRE(mode == 0 || (mode == 1 && up_c > 0 && n_total == 4 * up_c && up_c % 32 == 0), "conv_fprop_tc: bad mode/up_c");
  cudaStream_t st = (cudaStream_t)stream;

  FpropParams p;
  p.n_img = n_img; p.h_out = h_out; p.w_out = w_out;
  p.wt = min(pow2_ceil(w_out), kBlockM);
  p.ht = min(pow2_ceil(h_out), kBlockM / p.wt);
  p.nt = kBlockM / (p.wt * p.ht);
  p.tiles_w = qeb_cdiv(w_out, p.wt);
  p.tiles_h = qeb_cdiv(h_out, p.ht);
  const int tiles_n_img = qeb_cdiv(n_img, p.nt);
  const int m_tiles = p.tiles_w * p.tiles_h * tiles_n_img;
  p.kh = kh; p.kw = kw; p.ph = ph; p.pw = pw;
  p.cin = cin; p.kchunks = cin / kBlockK;
  p.n_total = n_total;
  p.bias = bias; p.relu = relu;
  p.out = out; p.out_pix_stride = out_cstride;
  p.mode = mode; p.up_c = up_c > 0 ? up_c : 1; p.accumulate = accumulate;

  int bn = block_n;
  if (bn == 0) {
    // largest tile that still yields >= ~1 wave of CTAs; never wider than the (padded) problem
    const int nmax = min(256, max(32, pow2_ceil(n_total)));
    bn = nmax;
    while (bn > 32 && (long long)m_tiles * qeb_cdiv(n_total, bn) < kNumSMs) bn >>= 1;
    if (mode == 1) while (bn > p.up_c) bn >>= 1;
  }
  QEB_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "conv_fprop_tc: block_n %d", bn);
  QEB_REQUIRE(mode == 0 || p.up_c % bn == 0, "conv_fprop_tc: block_n must divide up_c in pixel-shuffle mode");
  const int n_tiles = qeb_cdiv(n_total, bn);

  CUtensorMap ta, tb;
  {
    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)n_img};
    const uint64_t str[3] = {(uint64_t)x_cstride * 4, (uint64_t)w_in * x_cstride * 4, (uint64_t)h_in * w_in * x_cstride * 4};
    const uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)p.wt, (uint32_t)p.ht, (uint32_t)p.nt};
    int rc = make_tmap_f32(&ta, x, 4, dims, str, box);
    if (rc) return rc;
  }
  {
    const uint64_t ktot = (uint64_t)kh * kw * cin;
    const uint64_t dims[2] = {ktot, (uint64_t)n_total};
    const uint64_t str[1] = {ktot * 4};
    const uint32_t box[2] = {(uint32_t)kBlockK, (uint32_t)bn};
    int rc = make_tmap_f32(&tb, wpacked, 2, dims, str, box);
    if (rc) return rc;
  }
  switch (bn) {
    case 32: return launch_fprop<32>(ta, tb, p, m_tiles, n_tiles, st);
    case 64: return launch_fprop<64>(ta, tb, p, m_tiles, n_tiles, st);
    case 128: return launch_fprop<128>(ta, tb, p, m_tiles, n_tiles, st);
    default: return launch_fprop<256>(ta, tb, p, m_tiles, n_tiles, st);
  }
}

// =================================================================================================================
// Weight gradient: dW = sum over pixels of (shifted input) x (output gradient). Both operands are read straight from
// the NHWC activations, i.e. M/N (channel)-major with the pixel index as GEMM-K, so the MMA uses MN-major descriptors in the
// SWIZZLE_128B_BASE32B layout: a TMA box {32 ch, wt, ht, nt} of 32 pixels is one [32 px][32 ch] block = eight 4x32 atoms.
// The M side of a CTA is four such blocks ("row blocks" = (filter tap, 32-channel group) pairs of the shifted input),
// the N side BLOCK_N/32 blocks of dY channels. The pixel range is split across gridDim.z (split-K); partial results are
// combined with fp32 red.global.add into the (zero-initialised or accumulating) gradient in torch's own weight layout.
namespace {

constexpr int kWgPix = 32;                 // pixels (GEMM-K) per stage
constexpr int kBoxBytes = kWgPix * 32 * 4;  // 4 KB: [32 px][32 ch]

struct WgradParams {
  int n_img, h_out, w_out;
  int wt, ht, nt;  // pixel box shape, wt*ht*nt == 32
  int tiles_w, tiles_h, tiles_total;
  int kh, kw, ph, pw;
  int a_groups;       // 32-channel groups on the A side per tap
  int row_blocks;     // taps * a_groups
  int n_total;        // B-side channels
  int per_split;      // pixel tiles per gridDim.z slice
  int a_map_per_tap;  // 1: tap selects the A tensor map (ConvTranspose sub-lattices), no coordinate shift
  float* out;
  long long s_rowc, s_kh, s_kw, s_col;  // element strides of the gradient tensor
};

struct TmapArray4 {
  CUtensorMap m[4];
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_tc_kernel(const __grid_constant__ TmapArray4 tmaps_a, const __grid_constant__ CUtensorMap tmap_b,
                     const WgradParams p) {
  using Cfg = FpropCfg<BLOCK_N>;  // same stage geometry: 16 KB A side + BLOCK_N*128 B side
  constexpr int kBBoxes = BLOCK_N / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int t_begin = blockIdx.z * p.per_split;
  const int t_end = min(t_begin + p.per_split, p.tiles_total);
  const int rb0 = tile_m * 4;
  const int n_rb = min(4, p.row_blocks - rb0);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < (p.a_map_per_tap ? 4 : 1); ++i) prefetch_tmap(&tmaps_a.m[i]);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (t_begin < t_end) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t bytes = (uint32_t)(n_rb + kBBoxes) * kBoxBytes;
        for (int t = t_begin; t < t_end; ++t) {
          const int tw = t % p.tiles_w, th = (t / p.tiles_w) % p.tiles_h, tn = t / (p.tiles_w * p.tiles_h);
          const int w0 = tw * p.wt, h0 = th * p.ht, n0 = tn * p.nt;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(&full_bar[stage], bytes);
          for (int i = 0; i < n_rb; ++i) {
            const int rb = rb0 + i;
            const int tap = rb / p.a_groups, cg = rb - tap * p.a_groups;
            if (p.a_map_per_tap) {
              tma_load_4d(sa + i * kBoxBytes, &tmaps_a.m[tap], &full_bar[stage], cg * 32, w0, h0, n0);
            } else {
              const int dy = tap / p.kw - p.ph, dx = tap % p.kw - p.pw;
              tma_load_4d(sa + i * kBoxBytes, &tmaps_a.m[0], &full_bar[stage], cg * 32, w0 + dx, h0 + dy, n0);
            }
          }
#pragma unroll
          for (int j = 0; j < kBBoxes; ++j)
            tma_load_4d(sb + j * kBoxBytes, &tmap_b, &full_bar[stage], tile_n * BLOCK_N + j * 32, w0, h0, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = instr_desc_tf32(kBlockM, BLOCK_N, 1, 1);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kWgPix / 8; ++k) {
            // 8 pixels (one swizzle atom of K) per MMA: +1024 B inside every [32 px][32 ch] block
            const uint64_t adesc = smem_desc_mnmajor_sw128_32b(sa + k * 1024, kBoxBytes, 512);
            const uint64_t bdesc = smem_desc_mnmajor_sw128_32b(sb + k * 1024, kBoxBytes, 512);
            mma_tf32_ss(tmem_base, adesc, bdesc, idesc, (t > t_begin) || (k != 0));
          }
          mma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        mma_commit(tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int rb = rb0 + q;  // row block = TMEM lane quarter
      const bool valid = rb < p.row_blocks;
      const int tap = valid ? rb / p.a_groups : 0, cg = valid ? rb - tap * p.a_groups : 0;
      float* row_out = p.out + (long long)(cg * 32 + (r & 31)) * p.s_rowc + (long long)(tap / p.kw) * p.s_kh +
                       (long long)(tap % p.kw) * p.s_kw;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        const int ncol = tile_n * BLOCK_N + c0;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ncol + j < p.n_total) atomicAdd(row_out + (long long)(ncol + j) * p.s_col, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

template <int BLOCK_N>
int launch_wgrad(const TmapArray4& ta, const CUtensorMap& tb, const WgradParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = FpropCfg<BLOCK_N>;
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr = true;
  }
  conv_wgrad_tc_kernel<BLOCK_N><<<grid, kThreads, Cfg::kSmemBytes, st>>>(ta, tb, p);
  QEB_LAUNCH_CHECK();
  qeb_count_launch();
  return QEB_OK;
}

}  // namespace

// Weight gradient of a stride-1 convolution / GEMM (mode 0) or of ConvTranspose2d 2x2 s2 (mode 1), accumulated with
// atomics into `dw` (caller zero-fills it unless it wants accumulation).
//  mode 0: a = layer input x (n_img,h_in,w_in,a_c @ a_cstride), b = dY (n_img,h_out,w_out,b_c @ b_cstride);
//          dw[b_ch][a_ch][kh][kw] (torch Conv2d / Linear layout), dw[(co*a_c + ci)*kh*kw + tap].
//  mode 1: a = dY of the transposed conv (n_img, 2h_out, 2w_out, a_c @ a_cstride), b = layer input x (n_img,h_out,w_out,b_c);
//          dw[b_ch (Cin)][a_ch (Cout)][2][2] (torch ConvTranspose2d layout).
QEB_API int qeb_conv_wgrad_tc(const float* a, int a_c, int a_cstride, int h_in, int w_in, const float* b, int b_c,
                              int b_cstride, int n_img, int h_out, int w_out, int kh, int kw, int ph, int pw, float* dw,
                              int mode, int block_n, void* stream) {
  QEB_REQUIRE(a && b && dw, "conv_wgrad_tc: null pointer");
  QEB_REQUIRE(a_c > 0 && a_c % 32 == 0 && b_c > 0 && b_c % 32 == 0, "conv_wgrad_tc: channels must be multiples of 32 (%d, %d)", a_c, b_c);
  QEB_REQUIRE(a_cstride % 4 == 0 && b_cstride % 4 == 0, "conv_wgrad_tc: channel strides must be multiples of 4");
  QEB_REQUIRE(mode == 0 || (mode == 1 && kh == 2 && kw == 2), "conv_wgrad_tc: bad mode");
  cudaStream_t st = (cudaStream_t)stream;
  WgradParams p;
  p.n_img = n_img; p.h_out = h_out; p.w_out = w_out;
  p.wt = min(pow2_ceil(w_out), kWgPix);
  p.ht = min(pow2_ceil(h_out), kWgPix / p.wt);
  p.nt = kWgPix / (p.wt * p.ht);
  p.tiles_w = qeb_cdiv(w_out, p.wt);
  p.tiles_h = qeb_cdiv(h_out, p.ht);
  p.tiles_total = p.tiles_w * p.tiles_h * qeb_cdiv(n_img, p.nt);
  p.kh = kh; p.kw = kw; p.ph = ph; p.pw = pw;
  p.a_groups = a_c / 32;
  p.row_blocks = kh * kw * p.a_groups;
  p.n_total = b_c;
  p.a_map_per_tap = mode;
  p.out = dw;
  if (mode == 0) {
    p.s_rowc = (long long)kh * kw; p.s_kh = kw; p.s_kw = 1; p.s_col = (long long)a_c * kh * kw;
  } else {
    p.s_rowc = 4; p.s_kh = 2; p.s_kw = 1; p.s_col = (long long)a_c * 4;
  }
  int bn = block_n;
  if (bn == 0) bn = min(256, max(32, pow2_ceil(b_c)));
  QEB_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "conv_wgrad_tc: block_n %d", bn);
  const int m_tiles = qeb_cdiv(p.row_blocks, 4), n_tiles = qeb_cdiv(b_c, bn);
  // split-K so that the grid covers the SMs a few times over, at least 8 pixel tiles per CTA
  int splits = qeb_cdiv(2 * kNumSMs, m_tiles * n_tiles);
  splits = max(1, min(splits, qeb_cdiv(p.tiles_total, 8)));
  p.per_split = qeb_cdiv(p.tiles_total, splits);
  splits = qeb_cdiv(p.tiles_total, p.per_split);

  TmapArray4 ta;
  CUtensorMap tb;
  const uint32_t box[4] = {32u, (uint32_t)p.wt, (uint32_t)p.ht, (uint32_t)p.nt};
  if (mode == 0) {
    const uint64_t dims[4] = {(uint64_t)a_c, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)n_img};
    const uint64_t str[3] = {(uint64_t)a_cstride * 4, (uint64_t)w_in * a_cstride * 4, (uint64_t)h_in * w_in * a_cstride * 4};
    int rc = make_tmap_f32(&ta.m[0], a, 4, dims, str, box, 1);
    if (rc) return rc;
    ta.m[1] = ta.m[2] = ta.m[3] = ta.m[0];
  } else {
    // four sub-lattices (dh,dw) of the 2h_out x 2w_out gradient image, each viewed as an h_out x w_out image
    const int H2 = 2 * h_out, W2 = 2 * w_out;
    for (int t = 0; t < 4; ++t) {
      const int dh = t >> 1, dw_ = t & 1;
      const uint64_t dims[4] = {(uint64_t)a_c, (uint64_t)w_out, (uint64_t)h_out, (uint64_t)n_img};
      const uint64_t str[3] = {(uint64_t)2 * a_cstride * 4, (uint64_t)2 * W2 * a_cstride * 4, (uint64_t)H2 * W2 * a_cstride * 4};
      int rc = make_tmap_f32(&ta.m[t], a + ((long long)dh * W2 + dw_) * a_cstride, 4, dims, str, box, 1);
      if (rc) return rc;
    }
  }
  {
    const uint64_t dims[4] = {(uint64_t)b_c, (uint64_t)w_out, (uint64_t)h_out, (uint64_t)n_img};
    const uint64_t str[3] = {(uint64_t)b_cstride * 4, (uint64_t)w_out * b_cstride * 4, (uint64_t)h_out * w_out * b_cstride * 4};
    int rc = make_tmap_f32(&tb, b, 4, dims, str, box, 1);
    if (rc) return rc;
  }
  const dim3 grid(m_tiles, n_tiles, splits);
  switch (bn) {
    case 32: return launch_wgrad<32>(ta, tb, p, grid, st);
    case 64: return launch_wgrad<64>(ta, tb, p, grid, st);
    case 128: return launch_wgrad<128>(ta, tb, p, grid, st);
    default: return launch_wgrad<256>(ta, tb, p, grid, st);
  }
}
