// Implicit-GEMM convolution / GEMM on tcgen05 (fp32 accumulate in TMEM), operands staged by TMA.
//
// One kernel family serves every dense contraction of the hot path whose operands are K-major:
//   conv fprop    models/model_crnn.py:47-56 (conv2..conv7), models/model_unet.py:_block (all but the 1-channel conv)
//   conv dgrad    the same kernel on dY with the spatially flipped, channel-transposed weights
//   GEMM          LSTM input projections (nn.LSTM, models/model_crnn.py:9,19), Linear (models/model_crnn.py:10,20),
//                 ConvTranspose2d 2x2 s2 (models/model_unet.py:25-44) as a per-pixel GEMM with a pixel-shuffle store
//
// Data layout: activations NHWC fp32 in HBM (channels contiguous, arbitrary channel stride so that concat buffers are
// read/written in place), optionally shadowed by fp16 copies with the same element layout; weights packed
// [Cout][tap][Cin] (K-major). The kernel is persistent: a CTA walks 128-pixel x BLOCK_N output tiles; for every filter tap
// and K block (32 tf32 channels, or 64 / 32 fp16 channels = one 128- / 64-byte swizzled row) one 4-D TMA box
// {channels, Wt, Ht, Nt} (Wt*Ht*Nt = 128 output pixels, shifted by the tap, out-of-bounds = zero padding) lands in shared
// memory as the K-major A tile, one 2-D box {channels, BLOCK_N} as the B tile; tcgen05.mma (kind::tf32 K = 8 or kind::f16
// K = 16 per instruction) consume the stage into one of two TMEM accumulators.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue (tcgen05.ld -> scale / bias / ReLU /
// mask / BatchNorm sums -> transposed through shared memory -> global, fp32 and optional fp16 shadow). Pipelines: smem
// full/empty mbarriers (2-8 stages, running across tiles), TMEM full/empty barriers per accumulator.
// The weight-gradient kernel (conv_wgrad_tc_kernel, below) reads both operands MN-major straight from the activations.
#include "tc_common.cuh"
#include "nn.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <unordered_map>
#include <string.h>

namespace tc {

static int make_tmap_any(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box, CUtensorMapDataType dtype, CUtensorMapSwizzle swz) {
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
  // A step encodes ~250 tensor maps (operands and, with the TMA-store epilogue, outputs) and the workspace layout is a pure
  // function of (B, H, W): the caching allocator hands the same addresses back step after step, so the encoded descriptors are
  // kept per host thread, keyed on everything that goes into them. An eagerly launched step (the unmodified trainers) then
  // pays a hash lookup instead of a driver call per map; a map is a pure function of its key, so a hit can never be stale.
  struct Key {
    uintptr_t base;
    uint64_t d[5], s[4];
    uint32_t b[5];
    int rank, dtype, swz;
  };
  Key key;
  memset(&key, 0, sizeof(key));
  key.base = (uintptr_t)base; key.rank = rank; key.dtype = (int)dtype; key.swz = (int)swz;
  for (int i = 0; i < rank; ++i) { key.d[i] = d[i]; key.b[i] = b[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.s[i] = s[i];
  struct KeyHash {
    size_t operator()(const Key& k) const {
      const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
      uint64_t h = 1469598103934665603ull;
      for (size_t i = 0; i < sizeof(Key); ++i) { h ^= p[i]; h *= 1099511628211ull; }
      return (size_t)h;
    }
  };
  struct KeyEq {
    bool operator()(const Key& a, const Key& c) const { return memcmp(&a, &c, sizeof(Key)) == 0; }
  };
  static thread_local std::unordered_map<Key, CUtensorMap, KeyHash, KeyEq> cache;
  static const bool use_cache = !(getenv("QEB_TMAP_CACHE") && atoi(getenv("QEB_TMAP_CACHE")) == 0);
  if (use_cache) {
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return QEB_OK;
    }
  }
  CUresult r = encode(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS && use_cache) {
    if (cache.size() >= 8192) cache.clear();
    cache.emplace(key, *out);
  }
  if (r != CUDA_SUCCESS) {
    qeb_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
                  (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0), (unsigned long long)(rank > 2 ? d[2] : 0),
                  (unsigned long long)(rank > 3 ? d[3] : 0), b[0], rank > 1 ? b[1] : 0, rank > 2 ? b[2] : 0, rank > 3 ? b[3] : 0);
    return QEB_ERR_CUDA;
  }
  return QEB_OK;
}

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_32b_atom) {
  return make_tmap_any(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32,
                       swizzle_32b_atom ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

// output tensor maps of the TMA-store epilogue: plain fp32 (128-byte rows of 32 channels) / fp16 (64-byte rows)
int make_tmap_store(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool half) {
  return make_tmap_any(out, base, rank, dims, strides_bytes, box, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                       half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box, int row_bytes) {
  return make_tmap_any(out, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                       row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace tc

namespace {

using namespace tc;

long long* g_timeline = nullptr;  // qeb_debug_set_timeline

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                     // fp32 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 4;  // 16 KB
constexpr int kThreads = 192;
constexpr int kBarBytes = 512;   // mbarrier area: full / empty per stage, full / empty per accumulator, weights, TMEM slot

struct FpropParams {
  int n_img, h_out, w_out;
  int wt, ht, nt;  // tile shape in output pixels, wt*ht*nt == 128
  int tiles_w, tiles_h;
  int kh, kw, ph, pw;
  int cin, kchunks;
  int n_total;        // GEMM N (all output channels)
  const float* bias;   // per GEMM-N column, nullable
  const float* scale;  // per GEMM-N column, nullable: v = acc*scale + bias
  int relu;
  float* out;
  long long osn, osh, osw;   // output strides (image, row, pixel) in elements
  const float* mask;         // nullable: v = mask(n,h,w,c) > 0 ? v : 0
  long long msn, msh, msw;
  int mode;                  // 0: plain NHWC store; 1: ConvTranspose 2x2 s2 pixel shuffle (N index = (dh*2+dw)*up_c + co)
  int up_c;
  int accumulate;            // 1: out += result (used when a gradient already holds a partial sum)
  int stages;                // depth of the shared-memory ring
  int m_tiles, n_tiles, splits;  // tile grid: tile t -> (K slice t / (m*n), N tile (t / m) % n, pixel tile t % m)
  int kb_per_split;          // K blocks (tap x 32-channel slice) per gridDim.z slice; split-K partial sums are combined with
                             // 16-byte vector reductions into a zero-filled output (plain epilogue only)
  int vec_ok;                // 16-byte aligned rows: float4 stores allowed
  int a_map_per_tap;         // 1: tap selects the A tensor map (ConvTranspose dgrad sub-lattices), no coordinate shift
  long long* timeline;       // debugging aid (qeb_debug_set_timeline): per-CTA clock64 stamps, NULL in production
  double* stats;             // fused BatchNorm statistics (TcEpilogue::bn_stats) or backward reductions (bn_red), NULL = off
  const float* bn_scsh;      // non-NULL: `mask` holds z of the layer below; mask = z*scale + shift > 0, second sum = g*xhat
  int kblk;                  // operand elements per K block: 32 (tf32, or fp16 in 64-byte rows) or 64 (fp16 in 128-byte rows)
  __half* out16;             // optional fp16 shadow of the output (same element layout as `out`): the next layer's operand
  int round_out;             // 1: fp32 stores are rounded to tf32 (the tensor is a tf32 operand of a later contraction)
  int* amax;                 // LSM kernels only (fused log-softmax head): arg-max class per output row, nullable
  int tma_out;               // 1: the epilogue stages each 32 x 32 chunk in shared memory and a TMA store writes it (tmap_out /
                             // tmap_out16 kernel parameters); bw / bh: the per-warp box {32 ch, bw, bh, 32 / (bw * bh)}
  int bw, bh;
  int win_bytes, w_bytes;    // WIN kernels: bytes of one window stage (180 operand rows, 1 KB aligned) / of the resident weight area
  int epi_bytes;             // bytes of the epilogue staging area (FpropCfg::kEpiBytes, + kEpi16Bytes with an fp16 TMA store)
  const float* alpha;        // device scalar or NULL: the accumulator is multiplied by *alpha first (1 / scale of a scaled fp16 operand)
  const float* out16_scale;  // device scalar or NULL: the fp16 shadow holds v * *out16_scale (nn.cuh GradShadow)
  unsigned* gamax;           // device or NULL: running maximum of |v| over the stored values (fp32 bit pattern)
  int out16_cols;            // the fp16 shadow (and the running maximum) cover output channels [0, out16_cols) only
  int mc_dim;                // > 0: A-tile multicast - the kernel runs as clusters of TWO CTAs that work on the two N tiles 2j, 2j + 1 of the
                             // same pixel tile; each loads HALF of the shared A tile (tensor map m[1]: the box halved along box dimension
                             // mc_dim, mc_half = half its extent) and multicasts it into both CTAs' rings, so the A operand crosses
                             // L2 -> SM once per pair. The kernels are bound by that path (~43 B per clock and SM with all SMs streaming).
  int mc_half;
};

struct TmapOut {
  CUtensorMap f32, f16;
};

struct TmapArray4 {
  CUtensorMap m[4];
};

// ROWB: bytes per operand row in shared memory = one swizzle span: 128 (32 tf32 or 64 fp16) or 64 (32 fp16, layers with
// 32 input channels)
template <int BLOCK_N, int ROWB = 128>
struct FpropCfg {
  static constexpr int kABytes = kBlockM * ROWB;
  static constexpr int kBBytes = BLOCK_N * ROWB;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBytes = 4 * 4096;              // per epilogue warp: 32 x 32 floats (store transposition / TMA-store staging)
  static constexpr int kEpi16Bytes = 4 * 2048;            // + 32 x 32 halves when the kernel also writes an fp16 shadow through TMA
  // The kernel is persistent: a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... with the TMA ring running
  // across tile boundaries and two TMEM accumulators, so that the epilogue of one tile overlaps the main loop of the
  // next. Narrow tiles (short main loops, epilogue-heavy) still run two CTAs per SM to double the epilogue warps.
  static constexpr int kCtasPerSm = (BLOCK_N >= 256) ? 1 : 2;
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kMaxStages = 8;
  static constexpr int kMaxSmem = 227 * 1024;
  static constexpr int kSmBudget = 224 * 1024;   // what resident CTAs share: 228 KB per SM minus 1 KB reserved per CTA and slack
  static constexpr int kStatBytes = 2 * BLOCK_N * 4;   // per-CTA (sum, sum of squares) accumulators of the fused BN statistics
  static constexpr int smem_bytes(int stages, int epi_bytes = kEpiBytes) {
    return stages * kStageBytes + epi_bytes + kStatBytes + 1024 /*align slack*/ + kBarBytes;
  }
  static int resident(long long n_tiles_total) {
    int r = (int)((n_tiles_total + kNumSMs - 1) / kNumSMs);
    return r > kCtasPerSm ? kCtasPerSm : (r < 1 ? 1 : r);
  }
  // every byte of shared memory the resident CTAs leave goes into ring stages
  static int pick_stages(long long n_tiles_total, int epi_bytes = kEpiBytes) {
    int stages = (kSmBudget / resident(n_tiles_total) - 1280 - epi_bytes - kStatBytes) / kStageBytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) stages = 2;
    return stages;
  }
};

// Fused epilogue of one 32-column chunk of one output pixel: per-channel scale / bias, ReLU, ReLU-backward mask, then a
// plain, accumulating, pixel-shuffled or split-K (vector reduction) store.
__device__ __forceinline__ void fprop_epilogue_store(const FpropParams& p, float (&v)[32], int n, int h, int w, int ncol,
                                                     bool split) {
  float* dst;
  if (p.mode == 0) {
    dst = p.out + n * p.osn + h * p.osh + w * p.osw + ncol;
  } else {
    const int sub = ncol / p.up_c, co = ncol - sub * p.up_c;
    dst = p.out + n * p.osn + (2 * h + (sub >> 1)) * p.osh + (2 * w + (sub & 1)) * p.osw + co;
  }
  const int lim = min(32, p.n_total - ncol);
  const int pcol = p.mode == 1 ? ncol % p.up_c : ncol;  // per-channel vectors are indexed by the output channel
  if (p.alpha) {
    const float al = __ldg(p.alpha);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= al;
  }
  if (p.scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= (j < lim) ? __ldg(p.scale + pcol + j) : 0.f;
  }
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += (j < lim) ? __ldg(p.bias + pcol + j) : 0.f;
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.mask) {
    const float* mk = p.mask + n * p.msn + h * p.msh + w * p.msw + ncol;
    if (lim == 32 && p.vec_ok) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(mk) + j);
        v[4 * j] = q4.x > 0.f ? v[4 * j] : 0.f;
        v[4 * j + 1] = q4.y > 0.f ? v[4 * j + 1] : 0.f;
        v[4 * j + 2] = q4.z > 0.f ? v[4 * j + 2] : 0.f;
        v[4 * j + 3] = q4.w > 0.f ? v[4 * j + 3] : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)   // static indices + predicate: a run-time trip count would put v[] into local memory
        if (j < lim) v[j] = mk[j] > 0.f ? v[j] : 0.f;
    }
  }
  if (split) {  // partial sum of this K slice: 16-byte vector reductions (plain epilogue, lim == 32 guaranteed)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                   "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                   : "memory");
  } else if (lim == 32 && p.vec_ok) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    if (p.round_out && !p.accumulate) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = qeb_tf32r(v[j]);
    }
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
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < lim) dst[j] = p.accumulate ? dst[j] + v[j] : (p.round_out ? qeb_tf32r(v[j]) : v[j]);
  }
}

// Warp-level epilogue of one 32-row x 32-column chunk. A thread owns one accumulator ROW (TMEM lane) - storing that row
// directly makes every warp store touch 32 different 128-byte lines (measured: 1.3-4 us of the CTA's ~6-13 us). The chunk is
// therefore transposed through 4 KB of shared memory (XOR-swizzled 16-byte slots, conflict-free both ways) so that a warp
// instruction writes 4 complete 128-byte row segments. stg: this warp's staging area (the pipeline ring is idle by then).
__device__ __forceinline__ void fprop_epilogue_warp(const FpropParams& p, float (&v)[32], bool valid, int n, int h, int w,
                                                    int ncol, bool split, float* stg, int lane, long long* tle = nullptr,
                                                    float* cta_stats = nullptr, int c0_local = 0, float* gmax_run = nullptr) {
  const int lim = min(32, p.n_total - ncol);
  if (lim < 32 || !p.vec_ok) {  // ragged / unaligned rows (warp-uniform): per-thread path
    if (valid) fprop_epilogue_store(p, v, n, h, w, ncol, split);
    return;
  }
  const int pcol = p.mode == 1 ? ncol % p.up_c : ncol;
  if (p.alpha) {
    const float al = __ldg(p.alpha);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= al;
  }
  if (p.scale) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= __ldg(p.scale + pcol + j);
  }
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + pcol + j);
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  float* dst = nullptr;
  const float* mk = nullptr;
  if (valid) {
    if (p.mode == 0) {
      dst = p.out + n * p.osn + h * p.osh + w * p.osw + ncol;
    } else {
      const int sub = ncol / p.up_c, co = ncol - sub * p.up_c;
      dst = p.out + n * p.osn + (2 * h + (sub >> 1)) * p.osh + (2 * w + (sub & 1)) * p.osw + co;
    }
    if (p.mask) mk = p.mask + n * p.msn + h * p.msh + w * p.msw + ncol;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  if (tle) tle[9] = clock64();
  const int rsub = lane >> 3, c16 = lane & 7;
  float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = make_float4(0.f, 0.f, 0.f, 0.f);
  const float s16 = p.out16_scale ? __ldg(p.out16_scale) : 1.f;
  float gmax = 0.f;
  float4 bsc, bsh, bmu, bis;   // BatchNorm constants of this lane's four channels (fused backward reductions)
  if (p.bn_scsh) {
    const float* c4 = p.bn_scsh + ncol + c16 * 4;
    bsc = __ldg(reinterpret_cast<const float4*>(c4)); bsh = __ldg(reinterpret_cast<const float4*>(c4 + p.n_total));
    bmu = __ldg(reinterpret_cast<const float4*>(c4 + 2 * p.n_total)); bis = __ldg(reinterpret_cast<const float4*>(c4 + 3 * p.n_total));
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + rsub;
    float4 o = *reinterpret_cast<const float4*>(stg + row * 32 + ((c16 ^ (row & 7)) << 2));
    float* d = reinterpret_cast<float*>(__shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(dst), row));
    const float* m = reinterpret_cast<const float*>(__shfl_sync(FULL_MASK, reinterpret_cast<unsigned long long>(mk), row));
    if (d) {
      d += c16 * 4;
      float4 xh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m) {
        float4 q4 = __ldg(reinterpret_cast<const float4*>(m) + c16);
        if (p.bn_scsh) {   // q4 = z of the layer below: xhat for the reduction, relu(bn(z)) > 0 as the mask
          xh = make_float4((q4.x - bmu.x) * bis.x, (q4.y - bmu.y) * bis.y, (q4.z - bmu.z) * bis.z, (q4.w - bmu.w) * bis.w);
          q4 = make_float4(fmaf(q4.x, bsc.x, bsh.x), fmaf(q4.y, bsc.y, bsh.y), fmaf(q4.z, bsc.z, bsh.z), fmaf(q4.w, bsc.w, bsh.w));
        }
        o.x = q4.x > 0.f ? o.x : 0.f; o.y = q4.y > 0.f ? o.y : 0.f; o.z = q4.z > 0.f ? o.z : 0.f; o.w = q4.w > 0.f ? o.w : 0.f;
      }
      if (split) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
      } else {
        if (p.accumulate) {
          const float4 a = *reinterpret_cast<const float4*>(d);
          o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
        }
        *reinterpret_cast<float4*>(d) = p.round_out ? qeb_tf32r4(o) : o;
        if (ncol < p.out16_cols) gmax = fmaxf(fmaxf(gmax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        if (p.out16 && ncol < p.out16_cols) {   // fp16 shadow, same element offset
          // (saturating: a value beyond fp16's range becomes +-65504, not inf; the fp32 output keeps the exact value)
          const float4 os = make_float4(o.x * s16, o.y * s16, o.z * s16, o.w * s16);
          const __half2 h0 = __floats2half2_rn(fminf(fmaxf(os.x, -65504.f), 65504.f), fminf(fmaxf(os.y, -65504.f), 65504.f));
          const __half2 h1 = __floats2half2_rn(fminf(fmaxf(os.z, -65504.f), 65504.f), fminf(fmaxf(os.w, -65504.f), 65504.f));
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&h0);
          pk.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(p.out16 + (d - p.out)) = pk;
        }
      }
      ssum.x += o.x; ssum.y += o.y; ssum.z += o.z; ssum.w += o.w;
      if (!p.bn_scsh) xh = o;   // forward statistics: second sum = sum of squares
      ssq.x = fmaf(o.x, xh.x, ssq.x); ssq.y = fmaf(o.y, xh.y, ssq.y); ssq.z = fmaf(o.z, xh.z, ssq.z); ssq.w = fmaf(o.w, xh.w, ssq.w);
    }
  }
  if (cta_stats) {
    // fused BatchNorm statistics: this lane holds the sums of channels [4 c16, 4 c16 + 4) over its 8 rows; fold the four
    // row groups of the warp (lanes c16, c16 + 8, ...) and add into the CTA's shared-memory accumulators
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      ssum.x += __shfl_xor_sync(FULL_MASK, ssum.x, off); ssum.y += __shfl_xor_sync(FULL_MASK, ssum.y, off);
      ssum.z += __shfl_xor_sync(FULL_MASK, ssum.z, off); ssum.w += __shfl_xor_sync(FULL_MASK, ssum.w, off);
      ssq.x += __shfl_xor_sync(FULL_MASK, ssq.x, off); ssq.y += __shfl_xor_sync(FULL_MASK, ssq.y, off);
      ssq.z += __shfl_xor_sync(FULL_MASK, ssq.z, off); ssq.w += __shfl_xor_sync(FULL_MASK, ssq.w, off);
    }
    if (rsub == 0) {
      float* a = cta_stats + 2 * (c0_local + c16 * 4);
      atomicAdd(a + 0, ssum.x); atomicAdd(a + 1, ssq.x); atomicAdd(a + 2, ssum.y); atomicAdd(a + 3, ssq.y);
      atomicAdd(a + 4, ssum.z); atomicAdd(a + 5, ssq.z); atomicAdd(a + 6, ssum.w); atomicAdd(a + 7, ssq.w);
    }
  }
  if (gmax_run) *gmax_run = fmaxf(*gmax_run, gmax);   // folded over the CTA's tiles, one reduction per warp at the end of the kernel
  __syncwarp();
}

// ---- TMA-store epilogue ----------------------------------------------------------------------------------------------------
// A thread owns one accumulator row = 32 consecutive channels of one output pixel = one 128-byte row of the output tensor. The
// warp's 32 x 32 chunk is written to shared memory exactly in the SWIZZLE_128B layout of a TMA box {32 ch, bw, bh, bn} (the
// warp's 32 rows are such a box of the tile: rows run w-fastest) and ONE cp.async.bulk.tensor store per chunk writes it -
// full lines, image edges clipped by the hardware. The fp16 operand shadow leaves the same way (64-byte rows, SWIZZLE_64B).
// Against the transposing epilogue (fprop_epilogue_warp: 8 x {ld.shared, 4 shuffles, st.global} per chunk, measured
// 3.0-3.4 k cycles per chunk) this is 8 + 4 st.shared, one fence and one instruction of one lane.
__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}

// stg32 / stg16: this warp's staging areas (shared-window addresses, 4 KB / 2 KB, 1024-byte aligned). (w0, h0, n0): output
// coordinates of the warp's first row. Lane 0 owns the warp's bulk-store group: it waits for the previous chunk's store to
// have READ the staging area before anybody overwrites it.
__device__ __forceinline__ void fprop_epilogue_tma(const FpropParams& p, const TmapOut& to, float (&v)[32], bool valid, int n, int h,
                                                   int w, int w0, int h0, int n0, int ncol, uint32_t stg32, uint32_t stg16, int lane,
                                                   float* cta_stats, int c0_local, float* gmax_run) {
  if (p.alpha) {
    const float al = __ldg(p.alpha);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= al;
  }
  if (p.scale) {
    const float sv = __ldg(p.scale + ncol + lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= __shfl_sync(FULL_MASK, sv, j);
  }
  if (p.bias) {
    const float bv = __ldg(p.bias + ncol + lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(FULL_MASK, bv, j);
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (p.mask) {
    if (valid) {   // ReLU-backward mask of the layer below: this pixel's 32 channels are one 128-byte row of the mask tensor
      const float4* mk = reinterpret_cast<const float4*>(p.mask + n * p.msn + h * p.msh + w * p.msw + ncol);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q4 = __ldg(mk + j);
        v[4 * j] = q4.x > 0.f ? v[4 * j] : 0.f;
        v[4 * j + 1] = q4.y > 0.f ? v[4 * j + 1] : 0.f;
        v[4 * j + 2] = q4.z > 0.f ? v[4 * j + 2] : 0.f;
        v[4 * j + 3] = q4.w > 0.f ? v[4 * j + 3] : 0.f;
      }
    }
  }
  if (p.round_out) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = qeb_tf32r(v[j]);
  }
  if (cta_stats && !valid) {   // rows outside the image are clipped by the store but must not reach the statistics
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (lane == 0) tma_store_wait_read();
  __syncwarp();
  const uint32_t row32 = stg32 + (uint32_t)lane * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j) st_shared_v4(row32 + (uint32_t)((j ^ (lane & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  const bool shadow = p.out16 && ncol < p.out16_cols;
  if (p.gamax && valid && ncol < p.out16_cols) {   // running maximum of the stored gradient (rows outside the image are clipped by the store: not counted);
    float m = *gmax_run;     // folded over the CTA's tiles, one reduction per warp at the end of the kernel
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, fabsf(v[j]));
    *gmax_run = m;
  }
  if (shadow) {
    const uint32_t row16 = stg16 + (uint32_t)lane * 64u;
    const float s16 = p.out16_scale ? __ldg(p.out16_scale) : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st_shared_v4u(row16 + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4), pack_half2(v[8 * j] * s16, v[8 * j + 1] * s16),
                    pack_half2(v[8 * j + 2] * s16, v[8 * j + 3] * s16), pack_half2(v[8 * j + 4] * s16, v[8 * j + 5] * s16),
                    pack_half2(v[8 * j + 6] * s16, v[8 * j + 7] * s16));
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(&to.f32, stg32, ncol, w0, h0, n0);
    if (shadow) tma_store_4d(&to.f16, stg16, ncol, w0, h0, n0);
    tma_store_commit();
  }
  if (cta_stats) {
    // fused BatchNorm statistics: lane c sums column c of the staged chunk (conflict-free: a row's 32 words are distinct banks)
    float s1 = 0.f, s2 = 0.f;
    const uint32_t cw = (uint32_t)(lane & 3) << 2, cs = (uint32_t)(lane >> 2);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float x = ld_shared_f32(stg32 + (uint32_t)i * 128u + ((cs ^ (uint32_t)(i & 7)) << 4) + cw);
      s1 += x;
      s2 = fmaf(x, x, s2);
    }
    float* a = cta_stats + 2 * (c0_local + lane);
    atomicAdd(a, s1);
    atomicAdd(a + 1, s2);
  }
}

// torch.argmax ordering: NaN counts as the largest value, the first maximal index wins (cer.cu greedy_decode_kernel)
__device__ __forceinline__ bool lsm_gt(float a, float b) { return (a > b) || ((a != a) && !(b != b)); }

// Fused head of the CRNN (models/model_crnn.py:20, fn.log_softmax(self.linear(x), 2); utils.py:78-89, the per-frame arg-max
// of pred_to_string): all n_total <= BLOCK_N classes of an output row are columns of ONE accumulator row, i.e. they sit in
// one TMEM lane = one epilogue thread, so the row's max, log-sum-exp and arg-max need no cross-thread reduction: the row is
// read from TMEM once into registers, lp = (logit - max) - log(sum exp) and its first arg-max are computed in place, and the
// tile's log-probs - one contiguous block of the (T*B, V) output - leave through a shared-memory staging tile with full-width
// coalesced stores (a thread storing its own row would touch 32 different lines per warp store). The logits never reach HBM.
constexpr int kLsmPitch = 129;                                     // odd row pitch of the staging tile: conflict-free both ways
constexpr int kLsmStageBytes = (kBlockM * kLsmPitch * 4 + 1023) & ~1023;
// sbias: BLOCK_N floats of shared memory (the per-CTA statistics area, unused by this variant)
template <int BLOCK_N>
__device__ __forceinline__ void lsm_epilogue(const FpropParams& p, uint32_t taddr, int row, int row0, float* stage, float* sbias) {
  static_assert(BLOCK_N == 128, "the fused head keeps a whole row of <= 128 classes in registers");
  const int nt = p.n_total;
  // bias staged once per tile (-inf beyond the last class: those columns drop out of max, sum and arg-max by themselves);
  // per-element `if (j < nt) v += __ldg(bias + j)` compiles to a branch and an exposed global load per class (measured:
  // 48 k cycles per tile)
  sbias[row] = row < nt ? (p.bias ? __ldg(p.bias + row) : 0.f) : -INFINITY;
  float v[BLOCK_N];
#pragma unroll
  for (int c = 0; c < BLOCK_N / 32; ++c) tmem_ld32(taddr + (uint32_t)(32 * c), v + 32 * c);
  tmem_ld_wait();
  asm volatile("bar.sync 1, 128;" ::: "memory");
  // four independent chains per reduction (a single running max / sum / arg-max is 128 dependent steps for the one warp
  // per scheduler that runs here)
  float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < BLOCK_N; ++j) {
    v[j] += sbias[j];
    m4[j & 3] = fmaxf(m4[j & 3], v[j]);
  }
  const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < BLOCK_N; ++j) s4[j & 3] += __expf(v[j] - m);    // exp(-inf) = 0 for the padding columns
  const float ls = logf((s4[0] + s4[1]) + (s4[2] + s4[3]));
  float* srow = stage + row * kLsmPitch;
  // arg-max of the stored log-probs, first maximal index: chain q scans the contiguous quarter [32q, 32q + 32), the
  // quarters are merged in index order with a strict comparison
  float b4[4];
  int i4[4];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    b4[q4] = v[32 * q4] - m - ls;
    i4[q4] = 32 * q4;
    srow[32 * q4] = b4[q4];
  }
#pragma unroll
  for (int jj = 1; jj < 32; ++jj) {
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int j = 32 * q4 + jj;
      const float lp = v[j] - m - ls;
      const bool gt = lsm_gt(lp, b4[q4]);                         // padding columns are -inf: they never win
      b4[q4] = gt ? lp : b4[q4];
      i4[q4] = gt ? j : i4[q4];
      srow[j] = lp;
    }
  }
  float best = b4[0];
  int bi = i4[0];
#pragma unroll
  for (int q4 = 1; q4 < 4; ++q4) {
    const bool gt = lsm_gt(b4[q4], best);
    best = gt ? b4[q4] : best;
    bi = gt ? i4[q4] : bi;
  }
  const int rows_here = min(kBlockM, p.w_out - row0);
  if (row < rows_here && p.amax) p.amax[row0 + row] = min(bi, nt - 1);
  // the tile's rows are one contiguous block of the (T*B, V) output: each warp copies whole rows with coalesced stores
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const int wq = (threadIdx.x >> 5) - 2, lane = threadIdx.x & 31;
  float* dst = p.out + (long long)row0 * nt;
#pragma unroll 4
  for (int r = wq; r < rows_here; r += 4) {
    const float* sr = stage + r * kLsmPitch;
    float* dr = dst + (long long)r * nt;
#pragma unroll
    for (int c = 0; c < BLOCK_N / 32; ++c)
      if (lane + 32 * c < nt) dr[lane + 32 * c] = sr[lane + 32 * c];
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");   // staging tile and bias are free for the CTA's next tile
}

// EPI: which epilogue this instance carries (one per instance keeps the kernels small: the transposing epilogue alone is
// ~1000 instructions per variant branch) - kEpiLegacy: transposition through shared memory, every epilogue feature;
// kEpiTma: TMA-store epilogue; kEpiLsm: fused log-softmax head
enum { kEpiLegacy = 0, kEpiTma = 1, kEpiLsm = 2 };
// WIN ("window" variant, 3x3 pad-1 convolutions on images of at least 16 rows): the kernel above loads the tap-shifted
// 128-pixel tile once per filter tap, i.e. 9 x 128 operand rows per tile and channel slice - and the TMA unit retires a box
// at ~3.4 cycles per ROW whatever its width (measured: 32 -> 32 and 64 -> 32 channels at 32 x 128 take the same 29 us of
// main loop with half / all of the bytes), so the few-channel full-resolution layers are bound by TMA rows, not by bytes,
// L2 or the tensor pipe. Here a tile is 8 x 16 output pixels and ONE box {channels, 8 + 2, 16 + 2} - the tile with its halo,
// zero padding by TMA's out-of-bounds fill - lands per channel slice: 180 rows instead of 1152. A tap (dy, dx) is the same
// window read from row (dy + 1) * 10 + (dx + 1) on: tile row hh (8 pixels = one 8-row core-matrix group) starts 10 window
// rows after tile row hh - 1, which is exactly what the shared-memory descriptor's stride-byte-offset expresses (SBO =
// 10 rows instead of 8), and the swizzle is a function of absolute shared-memory address bits, so a descriptor may start at
// any row. The nine weight tiles of every channel slice stay RESIDENT in shared memory for all tiles of a CTA.
constexpr int kWinW = 8, kWinH = 16, kWinRows = (kWinW + 2) * (kWinH + 2);

// WIN kernels run one CTA per SM and are epilogue-bound on the few-channel layers (measured 1.3-1.8 k cycles of epilogue per
// 128 x 32 tile against ~1 k of everything else): they carry TWO epilogue warpgroups, group g draining accumulator g, i.e.
// every other tile, with its own staging and statistics areas.
template <bool WIN>
constexpr int fprop_threads() { return WIN ? 64 + 2 * 128 : kThreads; }
// Accumulators of a WIN kernel: kAccs of BLOCK_N columns, the MMA warp works on kAccs / 2 tiles at a time, tap by tap across
// the batch (independent accumulators back to back, one shared weight descriptor). Measured with 8 accumulators / batches
// of 4: no gain (32 -> 32 at 32 x 128: 21.7 -> 22.8 us) - the narrow tiles are bound by the tensor core's shared-memory
// operand reads (every 128 x 32 x 16 MMA re-reads 4 KB of the window for 16 cycles of tensor work), not by the latency of
// dependent accumulations - so the default is the plain double buffer (2 accumulators, batches of 1).
template <int BLOCK_N>
constexpr int win_accs() { return 2; }

template <int BLOCK_N, int ROWB, bool F16, int EPI = kEpiLegacy, bool WIN = false>
__global__ void __launch_bounds__(fprop_threads<WIN>(), (EPI == kEpiLsm || WIN) ? 1 : FpropCfg<BLOCK_N, ROWB>::kCtasPerSm)
conv_fprop_tc_kernel(const __grid_constant__ TmapArray4 tmaps_a, const __grid_constant__ CUtensorMap tmap_b,
                     const __grid_constant__ TmapOut tmaps_o, const FpropParams p) {
  using Cfg = FpropCfg<BLOCK_N, ROWB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_all = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem = smem_all + (WIN ? p.w_bytes : 0);   // the operand ring; WIN: the resident weight tiles come first
  const int stage_bytes = WIN ? p.win_bytes : Cfg::kStageBytes;
  constexpr int kGroups = WIN ? 2 : 1;              // epilogue warpgroups
  float* epi_stage = reinterpret_cast<float*>(smem + p.stages * stage_bytes);   // kGroups areas of p.epi_bytes
  float* cta_stats = epi_stage + kGroups * p.epi_bytes / 4;   // kGroups x [BLOCK_N][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes + kGroups * (p.epi_bytes + Cfg::kStatBytes));
  uint64_t* empty_bar = full_bar + p.stages;
  constexpr int kAccs = WIN ? win_accs<BLOCK_N>() : 2;   // TMEM accumulators of BLOCK_N columns
  constexpr int kBatch = kAccs / 2;                      // WIN: tiles the MMA warp interleaves
  uint64_t* tmem_full_bar = empty_bar + p.stages;        // [kAccs]: accumulator a holds a finished tile
  uint64_t* tmem_empty_bar = tmem_full_bar + kAccs;      // [kAccs]: the epilogue has read accumulator a
  uint64_t* w_full_bar = tmem_empty_bar + kAccs;         // WIN: the weight tiles of the current N tile have landed
  uint64_t* w_empty_bar = w_full_bar + 1;           // WIN: every MMA that reads the current weight tiles has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = p.kh * p.kw * p.kchunks;
  const int mn_tiles = p.m_tiles * p.n_tiles, total_tiles = mn_tiles * p.splits;
  const bool split = p.splits > 1;
  // A-tile multicast (FpropParams::mc_dim): cluster c of gridDim.x / 2 walks the tile PAIRS c, c + gridDim.x / 2, ...; pair i =
  // pixel tile i % m_tiles, N tiles 2 (i / m_tiles) + {0, 1}. it_* describe the walk of this CTA, tile_of() maps a walk index to
  // the tile number t the code below decomposes.
  const bool mc = !WIN && p.mc_dim > 0;
  const uint32_t crank = mc ? cluster_ctarank() : 0u;
  const int it_first = mc ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int it_step = mc ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int it_count = mc ? mn_tiles / 2 : total_tiles;
  auto tile_of = [&](int it) { return mc ? ((it / p.m_tiles) * 2 + (int)crank) * p.m_tiles + it % p.m_tiles : it; };
  long long* tl = p.timeline ? p.timeline + 16 * blockIdx.x : nullptr;
  if (tl && threadIdx.x == 0) { tl[0] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); tl[7] = sm; }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < (p.a_map_per_tap ? 4 : 1); ++i) prefetch_tmap(&tmaps_a.m[i]);
    prefetch_tmap(&tmap_b);
    if (EPI == kEpiTma) {
      prefetch_tmap(&tmaps_o.f32);
      if (p.out16) prefetch_tmap(&tmaps_o.f16);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (!WIN && p.mc_dim > 0) ? 2 : 1);   // multicast: a stage is free when BOTH CTAs' MMAs have read it
    }
    for (int a = 0; a < kAccs; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 128);   // every epilogue thread of the group that drains it arrives
    }
    mbar_init(w_full_bar, 1);
    mbar_init(w_empty_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kAccs * BLOCK_N);
  for (int i = threadIdx.x; i < (WIN ? 4 : 2) * BLOCK_N; i += fprop_threads<WIN>()) cta_stats[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (mc) cluster_sync_all();   // the peer's barriers are initialised before anything of ours can arrive on them
  qeb_pdl_sync();   // everything above (tensor-map prefetch, barrier init, TMEM allocation) overlaps the previous grid's tail
  if (tl && threadIdx.x == 0) tl[1] = clock64();

  if (warp == 0) {
    // ===== TMA producer: the whole warp runs the loop (warp-uniform control flow), one elected lane issues. With a
    // lane-divergent `if (lane == 0)` the compiler cannot keep descriptors and coordinates in uniform registers and wraps
    // every TMA / MMA instruction in an ELECT + R2UR + BRA.U.ANY lane loop (~60 cycles per instruction) =====
    int stage = 0;
    uint32_t phase = 0;
    if constexpr (WIN) {
      int cur_n = -1;
      uint32_t w_phase = 0;
      const int G = min(kBatch, p.stages);
      int t = blockIdx.x;
      while (t < total_tiles) {
        // a batch: up to G consecutive tiles of this CTA that share the N tile (= the resident weights)
        const int tile_n = t / p.m_tiles;
        int nb = 1;
        while (nb < G && t + nb * (int)gridDim.x < total_tiles && (t + nb * (int)gridDim.x) / p.m_tiles == tile_n) ++nb;
        if (tile_n != cur_n) {   // (re)load the 9 x kchunks weight tiles of this N tile
          if (cur_n >= 0) { mbar_wait(w_empty_bar, w_phase); w_phase ^= 1; }
          if (elect_one()) {
            mbar_expect_tx(w_full_bar, (uint32_t)(9 * p.kchunks * BLOCK_N * ROWB));
            for (int i = 0; i < 9 * p.kchunks; ++i) {
              const int kc = i / 9, tap = i - kc * 9;
              tma_load_2d(smem_all + (size_t)i * BLOCK_N * ROWB, &tmap_b, w_full_bar, tap * p.cin + kc * p.kblk, tile_n * BLOCK_N);
            }
          }
          __syncwarp();
          cur_n = tile_n;
        }
        for (int kc = 0; kc < p.kchunks; ++kc) {      // channel slice by channel slice across the batch (the MMA warp's order)
          for (int g = 0; g < nb; ++g) {
            const int tile_m = t + g * (int)gridDim.x - tile_n * p.m_tiles;
            const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full_bar[stage], (uint32_t)(kWinRows * ROWB));
              tma_load_4d(smem + stage * stage_bytes, &tmaps_a.m[0], &full_bar[stage], kc * p.kblk, tw * kWinW - 1, th * kWinH - 1, tn);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        t += nb * (int)gridDim.x;
      }
    } else
    for (int it = it_first; it < it_count; it += it_step) {
      const int t = tile_of(it);
      const int z = t / mn_tiles, r = t - z * mn_tiles;
      const int tile_n = r / p.m_tiles, tile_m = r - tile_n * p.m_tiles;
      const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
      const int w0 = tw * p.wt, h0 = th * p.ht, n0 = tn * p.nt;
      // multicast: this CTA's half of the A box starts `crank * mc_half` further along box dimension mc_dim
      const int mw = (mc && p.mc_dim == 1) ? (int)crank * p.mc_half : 0, mh = (mc && p.mc_dim == 2) ? (int)crank * p.mc_half : 0,
                mn = (mc && p.mc_dim == 3) ? (int)crank * p.mc_half : 0;
      const int kb_begin = z * p.kb_per_split, kb_end = min(kb_begin + p.kb_per_split, num_kb);
      int tap = kb_begin / p.kchunks, kc = kb_begin - tap * p.kchunks;
      int dy = tap / p.kw - p.ph, dx = tap % p.kw - p.pw;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (mc) tma_load_4d_mc(sa + crank * (Cfg::kABytes / 2), &tmaps_a.m[1], &full_bar[stage], kc * p.kblk, w0 + dx + mw, h0 + dy + mh, n0 + mn, (uint16_t)3);
          else if (p.a_map_per_tap) tma_load_4d(sa, &tmaps_a.m[tap], &full_bar[stage], kc * p.kblk, w0, h0, n0);
          else tma_load_4d(sa, &tmaps_a.m[0], &full_bar[stage], kc * p.kblk, w0 + dx, h0 + dy, n0);
          tma_load_2d(sb, &tmap_b, &full_bar[stage], tap * p.cin + kc * p.kblk, tile_n * BLOCK_N);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++kc == p.kchunks) {
          kc = 0;
          ++tap;
          dy = tap / p.kw - p.ph;
          dx = tap % p.kw - p.pw;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform loop, elected lane issues - see the producer) =====
    constexpr uint32_t idesc = F16 ? instr_desc_f16(kBlockM, BLOCK_N, 0, 0) : instr_desc_tf32(kBlockM, BLOCK_N, 0, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    if constexpr (WIN) {
      int cur_n = -1;
      uint32_t w_phase = 0;
      constexpr uint64_t kSwz = ROWB == 128 ? (2ull << 61) : (4ull << 61);
      // window operand: 8-row groups (one tile row) 10 operand rows apart; weight tiles: dense 8-row groups
      constexpr uint64_t kDescA = (1ull << 16) | ((uint64_t)((10 * ROWB) >> 4) << 32) | (1ull << 46) | kSwz;
      constexpr uint64_t kDescB = (1ull << 16) | ((uint64_t)((8 * ROWB) >> 4) << 32) | (1ull << 46) | kSwz;
      const int G = min(kBatch, p.stages);
      const uint32_t wbase = smem_u32(smem_all);
      int t = blockIdx.x, i_tile = 0;   // i_tile: running tile index of this CTA -> accumulator i_tile % kAccs
      while (t < total_tiles) {
        const int tile_n = t / p.m_tiles;
        int nb = 1;
        while (nb < G && t + nb * (int)gridDim.x < total_tiles && (t + nb * (int)gridDim.x) / p.m_tiles == tile_n) ++nb;
        if (tile_n != cur_n) { mbar_wait(w_full_bar, w_phase); w_phase ^= 1; cur_n = tile_n; }
        long long c0 = tl ? clock64() : 0;
        uint32_t td[kBatch];
#pragma unroll
        for (int g = 0; g < kBatch; ++g) {
          const int it = i_tile + g, a = it % kAccs;
          td[g] = tmem_base + (uint32_t)(a * BLOCK_N);
          if (g < nb) mbar_wait(&tmem_empty_bar[a], (uint32_t)(((it / kAccs) & 1) ^ 1));   // drained kAccs tiles ago
        }
        if (tl && lane == 0) tl[14] += clock64() - c0;     // waiting for the epilogue to drain accumulators
        tc_fence_after();
        for (int kc = 0; kc < p.kchunks; ++kc) {
          uint32_t sa[kBatch];
          int st_idx[kBatch];
          c0 = tl ? clock64() : 0;
#pragma unroll
          for (int g = 0; g < kBatch; ++g) {
            st_idx[g] = stage;
            sa[g] = smem_u32(smem + stage * stage_bytes);
            if (g < nb) {
              mbar_wait(&full_bar[stage], phase);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
          if (tl && lane == 0) tl[13] += clock64() - c0;   // waiting for windows to land
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t aoff = (uint32_t)((tap / 3) * (kWinW + 2) + tap % 3) * ROWB;
              const uint64_t bdesc = kDescB | (uint64_t)(((wbase + (uint32_t)((kc * 9 + tap) * BLOCK_N * ROWB)) >> 4) & 0x3FFF);
#pragma unroll
              for (int k = 0; k < ROWB / 32; ++k) {
#pragma unroll
                for (int g = 0; g < kBatch; ++g) {
                  if (g < nb) {
                    const uint64_t adesc = kDescA | (uint64_t)(((sa[g] + aoff) >> 4) & 0x3FFF);
                    if (F16) mma_f16_ss(td[g], adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0);
                    else mma_tf32_ss(td[g], adesc + 2 * k, bdesc + 2 * k, idesc, (kc | tap | k) != 0);
                  }
                }
              }
            }
#pragma unroll
            for (int g = 0; g < kBatch; ++g)
              if (g < nb) mma_commit(&empty_bar[st_idx[g]]);
          }
          __syncwarp();
        }
        const int t_next = t + nb * (int)gridDim.x;
        if (elect_one()) {
#pragma unroll
          for (int g = 0; g < kBatch; ++g)
            if (g < nb) mma_commit(&tmem_full_bar[(i_tile + g) % kAccs]);
          // the weight tiles may be replaced once these MMAs have retired
          if (t_next < total_tiles && t_next / p.m_tiles != tile_n) mma_commit(w_empty_bar);
        }
        __syncwarp();
        i_tile += nb;
        t = t_next;
      }
    } else
    for (int it = it_first; it < it_count; it += it_step) {
      const int t = tile_of(it);
      const int z = t / mn_tiles;
      const int kb_begin = z * p.kb_per_split, kb_end = min(kb_begin + p.kb_per_split, num_kb);
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);   // the epilogue has drained this accumulator (two tiles ago)
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (tl && lane == 0 && t == (int)blockIdx.x && kb == kb_begin) tl[2] = clock64();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = ROWB == 128 ? smem_desc_kmajor_sw128(sa) : smem_desc_kmajor_sw64(sa);
          const uint64_t bdesc = ROWB == 128 ? smem_desc_kmajor_sw128(sa + Cfg::kABytes) : smem_desc_kmajor_sw64(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < ROWB / 32; ++k) {
            // one MMA consumes 32 bytes of K (8 tf32 or 16 fp16) of every row: +2 in the (addr >> 4) field per step
            if (F16) mma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != kb_begin) || (k != 0));
            else mma_tf32_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != kb_begin) || (k != 0));
          }
          if (mc) mma_commit_mc(&empty_bar[stage], (uint16_t)3);   // ... in both CTAs: either may refill the stage's A half
          else mma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) mma_commit(&tmem_full_bar[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===== epilogue: warps 2..5 (WIN: and 6..9, group 1), TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int grp = WIN ? (warp - 2) >> 2 : 0;   // group g drains accumulator g = the tiles with (index within the CTA) % 2 == g
    float* const epi_stage_g = epi_stage + grp * (p.epi_bytes / 4);
    float* const cta_stats_g = cta_stats + grp * 2 * BLOCK_N;
    const int gtid = (int)threadIdx.x - 64 - grp * 128;   // 0..127 inside the group
    const int r = q * 32 + lane;  // row of the tile = TMEM lane
    const int ww = r % p.wt, hh = (r / p.wt) % p.ht, nn = r / (p.wt * p.ht);
    int acc = 0;
    uint32_t acc_phase = 0;
    int n_done = 0, n_seen = 0;
    float gmax_run = 0.f;   // running maximum of |stored value| (FpropParams::gamax)
    // fused BatchNorm statistics: the four epilogue warps add into shared-memory accumulators; they are flushed to global
    // memory (one double atomic per channel and CTA) when the CTA moves to another N tile and at the end
    int stat_n = -1;
    auto group_sync = [&]() {
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };
    auto flush_stats = [&](int tn) {
      group_sync();
      for (int c = gtid; c < BLOCK_N; c += 128) {
        const int col = tn * BLOCK_N + c;
        if (col < p.n_total) {
          atomicAdd(p.stats + col, (double)cta_stats_g[2 * c]);
          atomicAdd(p.stats + p.n_total + col, (double)cta_stats_g[2 * c + 1]);
        }
        cta_stats_g[2 * c] = 0.f;
        cta_stats_g[2 * c + 1] = 0.f;
      }
      group_sync();
    };
    for (int it = it_first; it < it_count; it += it_step, ++n_seen) {
      const int t = tile_of(it);
      if (WIN && (n_seen & 1) != grp) continue;   // the other group's tile
      if (WIN) { acc = n_seen % kAccs; acc_phase = (uint32_t)((n_seen / kAccs) & 1); }
      const int z = t / mn_tiles, rr = t - z * mn_tiles;
      const int tile_n = rr / p.m_tiles, tile_m = rr - tile_n * p.m_tiles;
      if (p.stats && tile_n != stat_n) {
        if (stat_n >= 0) flush_stats(stat_n);
        stat_n = tile_n;
      }
      const int tw = tile_m % p.tiles_w, th = (tile_m / p.tiles_w) % p.tiles_h, tn = tile_m / (p.tiles_w * p.tiles_h);
      const int w = tw * p.wt + ww, h = th * p.ht + hh, n = tn * p.nt + nn;
      const bool valid = (w < p.w_out) && (h < p.h_out) && (n < p.n_img);
      const long long e0 = tl ? clock64() : 0;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      const long long e1 = tl ? clock64() : 0;
      if (tl && threadIdx.x == 64) tl[11] += e1 - e0;       // epilogue warp waiting for a finished accumulator
      if (tl && threadIdx.x == 64 && n_done == 0) tl[4] = clock64();
      tc_fence_after();
      if constexpr (EPI == kEpiLsm) {
        lsm_epilogue<BLOCK_N>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N), r, tile_m * kBlockM,
                              reinterpret_cast<float*>(smem + p.stages * Cfg::kStageBytes + p.epi_bytes + Cfg::kStatBytes + kBarBytes), cta_stats);
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0), v);
          tmem_ld_wait();
          long long* tle = (tl && threadIdx.x == 64 && n_done == 0 && c0 == 0) ? tl : nullptr;
          if (tle) tle[8] = clock64();
          const int ncol = tile_n * BLOCK_N + c0;  // first GEMM-N column of this chunk
          if (ncol < p.n_total) {
            if constexpr (EPI == kEpiTma) {
              // first row of this warp's 32-row box: tile row 32 q
              const int r0 = q * 32;
              fprop_epilogue_tma(p, tmaps_o, v, valid, n, h, w, tw * p.wt + r0 % p.wt, th * p.ht + (r0 / p.wt) % p.ht,
                                 tn * p.nt + r0 / (p.wt * p.ht), ncol, smem_u32(epi_stage_g) + (uint32_t)q * 4096u,
                                 smem_u32(epi_stage_g) + 16384u + (uint32_t)q * 2048u, lane, p.stats ? cta_stats_g : nullptr, c0, &gmax_run);
            } else {
              fprop_epilogue_warp(p, v, valid, n, h, w, ncol, split, epi_stage_g + q * 1024, lane, tle, p.stats ? cta_stats_g : nullptr, c0,
                                  p.gamax ? &gmax_run : nullptr);
            }
          }
          if (tle) tle[10] = clock64();
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (tl && threadIdx.x == 64) tl[12] += clock64() - e1;   // epilogue body
      if (tl && threadIdx.x == 64 && n_done == 0) tl[5] = clock64();
      ++n_done;
      if (!WIN) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    if (p.stats && stat_n >= 0) flush_stats(stat_n);
    if (EPI != kEpiLsm && p.gamax) {
      gmax_run = warp_max(gmax_run);
      if (lane == 0 && gmax_run > 0.f) atomicMax(p.gamax, __float_as_uint(gmax_run));
    }
    if (EPI == kEpiTma && lane == 0) tma_store_wait_all();   // the staging areas must outlive the stores that read them
    if (tl && threadIdx.x == 64) tl[3] = n_done;
  }
  tc_fence_before();
  __syncthreads();
  if (mc) cluster_sync_all();   // the peer may multicast into this CTA's ring / arrive on its barriers until its own last tile is done
  if (warp == 1) tmem_dealloc(tmem_base, kAccs * BLOCK_N);
  if (tl && threadIdx.x == 32) tl[6] = clock64();
}

template <int BLOCK_N, int ROWB, bool F16, int EPI = kEpiLegacy, bool WIN = false>
int launch_fprop(const TmapArray4& ta, const CUtensorMap& tb, const TmapOut& to, const FpropParams& p_in, int m_tiles, int n_tiles,
                 int splits, cudaStream_t st) {
  using Cfg = FpropCfg<BLOCK_N, ROWB>;
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(conv_fprop_tc_kernel<BLOCK_N, ROWB, F16, EPI, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmem));
    attr = true;
  }
  FpropParams p = p_in;
  const long long total = (long long)m_tiles * n_tiles * splits;
  p.epi_bytes = Cfg::kEpiBytes + ((p.tma_out && p.out16) ? Cfg::kEpi16Bytes : 0);
  p.stages = Cfg::pick_stages(total, p.epi_bytes);
  constexpr bool LSM = EPI == kEpiLsm;
  if (LSM) p.stages = min(p.stages, (Cfg::kMaxSmem - Cfg::smem_bytes(0) - kLsmStageBytes) / Cfg::kStageBytes);
  int smem_total = Cfg::smem_bytes(p.stages, p.epi_bytes) + (LSM ? kLsmStageBytes : 0);
  int resident = Cfg::resident(total);
  if (WIN) {   // one CTA per SM: resident weights + as many window stages as fit (<= 8)
    p.win_bytes = (kWinRows * ROWB + 1023) & ~1023;
    p.w_bytes = 9 * p.kchunks * BLOCK_N * ROWB;
    const int fixed = p.w_bytes + 2 * (p.epi_bytes + Cfg::kStatBytes) + 1024 + kBarBytes;   // two epilogue groups
    p.stages = min(8, (Cfg::kMaxSmem - fixed) / p.win_bytes);
    QEB_REQUIRE(p.stages >= 2, "tc fprop (window): weights of %d bytes leave no room for two window stages", p.w_bytes);
    smem_total = fixed + p.stages * p.win_bytes;
    resident = 1;
  }
  p.m_tiles = m_tiles; p.n_tiles = n_tiles; p.splits = splits;
  int grid = (int)(total < (long long)kNumSMs * resident ? total : (long long)kNumSMs * resident);
  if (p.mc_dim > 0) {
    QEB_REQUIRE(!WIN && !LSM && splits == 1 && n_tiles % 2 == 0, "tc fprop: A-tile multicast needs an even number of N tiles and no split-K");
    grid &= ~1;   // whole clusters of two
  }
  // ".f16" / ".tf32": the operand kind, so that bench.py can rate each against its own measured dense peak
  ProfScope prof(LSM ? "tc_head_logsoftmax" : p.a_map_per_tap ? "tc_convT_dgrad" : (p.mode == 1 ? "tc_convT_fprop" : (F16 ? "tc_conv_fprop.f16" : "tc_conv_fprop.tf32")), st,
                 2.0 * p.n_img * p.h_out * p.w_out * (double)p.n_total * p.kh * p.kw * p.cin,
                 4.0 * ((double)p.n_img * p.h_out * p.w_out * (p.cin + p.n_total) + (double)p.n_total * p.kh * p.kw * p.cin));
  if (p.mc_dim > 0)
    QEB_CUDA(qeb_launch_cluster(conv_fprop_tc_kernel<BLOCK_N, ROWB, F16, EPI, WIN>, grid, fprop_threads<WIN>(), smem_total, st, 2, ta, tb, to, p));
  else
    QEB_CUDA(qeb_launch(conv_fprop_tc_kernel<BLOCK_N, ROWB, F16, EPI, WIN>, grid, fprop_threads<WIN>(), smem_total, st, ta, tb, to, p));
  qeb_count_launch();
  return QEB_OK;
}

int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

}  // namespace

namespace {

int tmap_img(CUtensorMap* out, const Img& a, const float* base, int c, long long sn, long long sh, long long sw, int w,
             int h, const uint32_t* box, int swz32) {
  const uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)a.n};
  const uint64_t str[3] = {(uint64_t)sw * 4, (uint64_t)sh * 4, (uint64_t)sn * 4};
  return make_tmap_f32(out, base, 4, dims, str, box, swz32);
}
// the fp16 shadow of an image: same element strides, 2-byte elements; box[0] = elements per K block (64 or 32)
int tmap_img16(CUtensorMap* out, const Img& a, const void* base16, const uint32_t* box) {
  const uint64_t dims[4] = {(uint64_t)a.c, (uint64_t)a.w, (uint64_t)a.h, (uint64_t)a.n};
  const uint64_t str[3] = {(uint64_t)a.sw * 2, (uint64_t)a.sh * 2, (uint64_t)a.sn * 2};
  return make_tmap_16(out, base16, 4, dims, str, box, (int)box[0] * 2);
}
// K-block width of the fp16 path: 64 elements (128-byte rows) when the channel count allows, else 32 (64-byte rows)
inline int kblk16(int cin) { return cin % 64 == 0 ? 64 : 32; }
inline bool strides_ok16(const Img& a) { return a.sn % 8 == 0 && a.sh % 8 == 0 && a.sw % 8 == 0; }

// window variant (conv_fprop_tc_kernel<..., WIN>): bytes the resident weight tiles may take so that three window stages, the
// epilogue staging and the barriers still fit into one CTA's shared memory
inline int win_weight_budget(int rowb) {
  const int win_bytes = (kWinRows * rowb + 1023) & ~1023;
  return 227 * 1024 - 4 * win_bytes - 2 * (4 * 4096 + 4 * 2048) - 2 * 2 * 128 * 4 - 1024 - kBarBytes;
}
// 3x3 pad-1 convolution on an image of >= 16 rows whose weights fit (at the narrowest N tile)
// ... and with enough 8 x 16 tiles for the one-CTA-per-SM grid to balance (>= 7 per CTA). Measured in the step (per launch,
// window vs per-tap loads): 32 x 128 images 36.0 -> 28.8 us (32 -> 32 channels), 40.7 -> 34.1 us (64 -> 32); 16 x 64 images (512
// tiles, 3.5 per CTA) 30.2 -> 34.3 us: those stay on the per-tap kernel. QEB_WIN=0 switches the variant off, =2 drops the
// tile-count condition.
inline bool win_eligible(int kh, int kw, int ph, int pw, int n_img, int h_out, int w_out, int cin, int kblk, int rowb) {
  static const int allow = getenv("QEB_WIN") ? atoi(getenv("QEB_WIN")) : 1;
  const long long tiles = (long long)n_img * qeb_cdiv(h_out, kWinH) * qeb_cdiv(w_out, kWinW);
  return allow && kh == 3 && kw == 3 && ph == 1 && pw == 1 && h_out >= kWinH && w_out >= kWinW &&
         (allow == 2 || tiles >= 7 * kNumSMs) && 9 * (cin / kblk) * 32 * rowb <= win_weight_budget(rowb);
}

// shared driver of the three K-major entry points. a_maps: number of A tensor maps already encoded in ta (1 or 4).
// f16: the A maps in ta_in describe the fp16 shadow (K block = kblk16(cin) elements) and ep.w16 holds the fp16 weights
int fprop_common(const TmapArray4& ta_in, bool per_tap, const Img& x_geom, const float* wpacked, int n_total, int kh,
                 int kw, int ph, int pw, int cin, const Img& out, int h_out, int w_out, const TcEpilogue& ep, int mode,
                 int up_c, const float* bias, cudaStream_t st, bool f16 = false, bool win = false, int mc_dim = 0, int mc_half = 0) {
  FpropParams p;
  p.n_img = x_geom.n; p.h_out = h_out; p.w_out = w_out;
  p.wt = min(pow2_ceil(w_out), kBlockM);
  p.ht = min(pow2_ceil(h_out), kBlockM / p.wt);
  p.nt = kBlockM / (p.wt * p.ht);
  if (win) { p.wt = kWinW; p.ht = kWinH; p.nt = 1; }   // window variant: 8 x 16 pixel tiles (the A map's box is the tile + halo)
  p.win_bytes = p.w_bytes = 0;
  p.tiles_w = qeb_cdiv(w_out, p.wt);
  p.tiles_h = qeb_cdiv(h_out, p.ht);
  const int m_tiles = p.tiles_w * p.tiles_h * qeb_cdiv(x_geom.n, p.nt);
  p.kh = kh; p.kw = kw; p.ph = ph; p.pw = pw;
  const int kblk = f16 ? kblk16(cin) : kBlockK;
  p.kblk = kblk;
  p.cin = cin; p.kchunks = cin / kblk;
  p.n_total = n_total;
  p.bias = bias; p.scale = ep.scale; p.relu = ep.relu;
  p.out = out.p; p.osn = out.sn; p.osh = out.sh; p.osw = out.sw;
  p.mask = nullptr; p.msn = p.msh = p.msw = 0;
  bool vec = ((uintptr_t)out.p & 15) == 0 && out.sn % 4 == 0 && out.sh % 4 == 0 && out.sw % 4 == 0;
  if (ep.mask) {
    p.mask = ep.mask->p; p.msn = ep.mask->sn; p.msh = ep.mask->sh; p.msw = ep.mask->sw;
    vec = vec && ((uintptr_t)p.mask & 15) == 0 && p.msn % 4 == 0 && p.msh % 4 == 0 && p.msw % 4 == 0;
  }
  p.vec_ok = vec;
  p.mode = mode; p.up_c = up_c > 0 ? up_c : 1; p.accumulate = ep.accumulate;
  p.a_map_per_tap = per_tap;
  p.timeline = g_timeline;
  p.stats = nullptr;
  p.bn_scsh = nullptr;
  QEB_REQUIRE(!ep.out16 || (p.vec_ok && n_total % 32 == 0 && ((uintptr_t)ep.out16 & 7) == 0),
              "tc fprop: an fp16 output shadow needs 16-byte aligned output rows and a multiple of 32 channels");
  p.out16 = static_cast<__half*>(ep.out16);
  p.alpha = ep.alpha;
  p.out16_scale = nullptr;
  p.gamax = nullptr;
  p.out16_cols = (ep.gs_cols > 0 && ep.gs_cols % 32 == 0) ? ep.gs_cols : n_total;
  if (ep.gs.amax || ep.gs.out16) {
    // gradient shadow / running maximum: the vector epilogues only (aligned full rows); the caller learns through gs_done
    const bool ok = p.vec_ok && n_total % 32 == 0 && mode == 0 && !ep.accumulate && !ep.out16 &&
                    (!ep.gs.out16 || (((uintptr_t)ep.gs.out16 & 7) == 0 && ep.gs.scale));
    if (ok) {
      p.gamax = ep.gs.amax;
      if (ep.gs.out16) { p.out16 = static_cast<__half*>(ep.gs.out16); p.out16_scale = ep.gs.scale; }
    }
    if (ep.gs_done) *ep.gs_done = ok ? 1 : 0;
  }
  p.round_out = ep.round_out;
  p.amax = ep.argmax;
  if (ep.log_softmax) {
    QEB_REQUIRE(n_total <= 128 && mode == 0 && !ep.scale && !ep.relu && !ep.mask && !ep.accumulate && !ep.out16 && !ep.bn_stats && !ep.bn_red,
                "tc fprop: the fused log-softmax head needs <= 128 classes and a plain bias epilogue");
    QEB_REQUIRE(!f16 || kblk == 64, "tc fprop: the fused log-softmax head needs a multiple of 64 input channels in fp16 mode");
    QEB_REQUIRE(x_geom.n == 1 && h_out == 1 && out.c == n_total && out.sw == n_total && ((uintptr_t)out.p & 15) == 0,
                "tc fprop: the fused log-softmax head writes a dense (rows, classes) matrix");
  }

  // widest tile that still yields about one wave of CTAs; never wider than the (padded) problem
  // 128, not the SM count: a layer with exactly 128 tiles at the wide N tile stays there instead of being narrowed to 256 tiles
  // (same-box A/B with fp16 operands: 112 / 120 / 128 -> 2.993-2.999 ms per step, 136 / 144 / 148 -> 3.04-3.05 ms)
  static const int min_ctas = getenv("QEB_TC_MIN_CTAS") ? atoi(getenv("QEB_TC_MIN_CTAS")) : 128;
  // Split-K with global reductions was the round-1 answer for the deep layers (tf32 operands: a narrowed N tile made every CTA stream
  // the whole fp32 A operand for a sliver of MMA work). With fp16 operands - forward AND backward since round 2 - the narrowed tile
  // is never slower per layer (scripts/exp/cluster_splitk_time.py, QEB_TC_CLUSTER=0 columns) and keeps the fused statistics, the
  // TMA-store epilogue and no zero-fill: same-box A/B of the step 3.206 -> 3.088 ms with split-K off. Off by default
  // (QEB_TC_SPLITK=1 brings it back); forward and input gradients are then bit-reproducible run to run.
  static const int allow_split = getenv("QEB_TC_SPLITK") ? atoi(getenv("QEB_TC_SPLITK")) : 0;
  const int num_kb = kh * kw * (cin / kblk);
  static const int bn_cap = getenv("QEB_TC_BN_MAX") ? atoi(getenv("QEB_TC_BN_MAX")) : 256;
  const int bn_max = min(bn_cap, max(32, pow2_ceil(n_total)));
  int bn = bn_max, splits = 1;
  // Few pixels, long K (the deep UNet levels and their input gradients): narrowing the N tile to fill the SMs makes every
  // CTA stream the whole A operand for a sliver of MMA work and the per-SM L2 read rate becomes the limit. With a plain
  // epilogue the K range is split instead and the partial sums are reduced into a zero-filled output.
  const bool plain = !ep.scale && !bias && !ep.relu && !ep.mask && !ep.accumulate && !ep.round_out && !ep.gs.out16 && !ep.gs.amax && mode == 0 && p.vec_ok &&
                     n_total % 32 == 0 &&
                     out.c == n_total && out.sw == n_total && img_flat(out);
  if (allow_split && !win && plain && (long long)m_tiles * qeb_cdiv(n_total, bn_max) * 2 <= min_ctas && num_kb >= 16) {
    // round DOWN: tiles beyond one per SM would make a few CTAs walk two tiles while the rest idle
    splits = min(num_kb / 8, min_ctas / (m_tiles * qeb_cdiv(n_total, bn_max)));
    if (splits < 1) splits = 1;
  }
  if (win) {
    // widest N tile whose nine weight tiles per channel slice stay resident beside >= 3 window stages, narrowed to fill the SMs
    splits = 1;
    const int rowb = kblk * (f16 ? 2 : 4);
    bn = min(128, bn_max);
    while (bn > 32 && (9 * (cin / kblk) * bn * rowb > win_weight_budget(rowb) || (long long)m_tiles * qeb_cdiv(n_total, bn) < min_ctas)) bn >>= 1;
  } else if (ep.log_softmax) {
    bn = 128; splits = 1;   // every class of a row in ONE accumulator row
  } else if (splits == 1) {
    while (bn > 32 && (long long)m_tiles * qeb_cdiv(n_total, bn) < min_ctas) bn >>= 1;
  }
  p.kb_per_split = qeb_cdiv(num_kb, splits);
  splits = qeb_cdiv(num_kb, p.kb_per_split);
  if (splits > 1) QEB_CUDA(cudaMemsetAsync(out.p, 0, (size_t)img_pixels(out) * n_total * sizeof(float), st));
  // BatchNorm statistics of the output: fused into the epilogue unless the K range is split (partial sums) or rows are ragged
  const bool stats_fused = ep.bn_stats && splits == 1 && p.vec_ok && n_total % 32 == 0 && mode == 0;
  if (stats_fused) p.stats = ep.bn_stats;
  // BatchNorm-backward reductions of the layer below: same accumulators, the mask source is that layer's z
  if (ep.bn_red && ep.bn_z && ep.bn_scsh && !ep.bn_stats && !ep.mask && splits == 1 && p.vec_ok && n_total % 32 == 0 && mode == 0 &&
      ep.bn_z->c == n_total && ((uintptr_t)ep.bn_z->p & 15) == 0 && ep.bn_z->sn % 4 == 0 && ep.bn_z->sh % 4 == 0 && ep.bn_z->sw % 4 == 0 &&
      ((uintptr_t)ep.bn_scsh & 15) == 0) {
    p.mask = ep.bn_z->p; p.msn = ep.bn_z->sn; p.msh = ep.bn_z->sh; p.msw = ep.bn_z->sw;
    p.bn_scsh = ep.bn_scsh;
    p.stats = ep.bn_red;
    if (ep.bn_red_fused) *ep.bn_red_fused = 1;
  }
  if (mode == 1) while (bn > p.up_c) bn >>= 1;
  QEB_REQUIRE(mode == 0 || p.up_c % bn == 0, "tc fprop: tile width %d must divide the up-conv channels %d", bn, p.up_c);
  const int n_tiles = qeb_cdiv(n_total, bn);
  // A-tile multicast between the two CTAs of a cluster (FpropParams::mc_dim): every layer with an even number of N tiles
  static const int allow_mc = getenv("QEB_TC_MCAST") ? atoi(getenv("QEB_TC_MCAST")) : 1;
  p.mc_dim = (allow_mc && mc_dim > 0 && !win && !ep.log_softmax && mode == 0 && !per_tap && splits == 1 && n_tiles % 2 == 0) ? mc_dim : 0;
  p.mc_half = mc_half;

  CUtensorMap tb;
  {
    const uint64_t ktot = (uint64_t)kh * kw * cin;
    const uint64_t dims[2] = {ktot, (uint64_t)n_total};
    const uint64_t str[1] = {ktot * (f16 ? 2 : 4)};
    const uint32_t box[2] = {(uint32_t)kblk, (uint32_t)bn};
    int rc = f16 ? make_tmap_16(&tb, ep.w16, 2, dims, str, box, kblk * 2) : make_tmap_f32(&tb, wpacked, 2, dims, str, box);
    if (rc) return rc;
  }
  // TMA-store epilogue for plain NHWC outputs with full 32-channel chunks (not: pixel shuffle, split-K reductions, +=, the
  // fused BatchNorm-backward reductions, ragged / unaligned rows, the log-softmax head)
  static const int allow_tma_out = getenv("QEB_TMA_STORE") ? atoi(getenv("QEB_TMA_STORE")) : 1;
  TmapOut to;
  memset(&to, 0, sizeof(to));
  p.tma_out = 0; p.bw = p.bh = 1;
  if (allow_tma_out && mode == 0 && splits == 1 && !ep.accumulate && !p.bn_scsh && p.vec_ok && n_total % 32 == 0 && !ep.log_softmax &&
      out.sw % 4 == 0 && (!p.out16 || out.sw % 8 == 0)) {
    p.bw = min(p.wt, 32);
    p.bh = min(p.ht, 32 / p.bw);
    const int bnn = 32 / (p.bw * p.bh);
    // a dimension of extent 1 may carry any stride in the Img (e.g. sh = 0 of the sequence-major conv7 output): give the
    // tensor map a legal one
    const uint64_t sw = (uint64_t)out.sw, sh = h_out > 1 ? (uint64_t)out.sh : sw * (uint64_t)w_out,
                   sn = x_geom.n > 1 ? (uint64_t)out.sn : sh * (uint64_t)h_out;
    const uint64_t dims[4] = {(uint64_t)n_total, (uint64_t)w_out, (uint64_t)h_out, (uint64_t)x_geom.n};
    const uint32_t box[4] = {32u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)bnn};
    const uint64_t str32[3] = {sw * 4, sh * 4, sn * 4};
    const uint64_t str16[3] = {sw * 2, sh * 2, sn * 2};
    const bool ok = sh % 4 == 0 && sn % 4 == 0 && (!p.out16 || (sh % 8 == 0 && sn % 8 == 0));
    if (ok && make_tmap_store(&to.f32, out.p, 4, dims, str32, box, false) == QEB_OK &&
        (!p.out16 || make_tmap_store(&to.f16, p.out16, 4, dims, str16, box, true) == QEB_OK))
      p.tma_out = 1;
  }
  int rc;
#define QEB_FPROP_CASE(BN, RB, F)                                                                                                      \
  rc = win ? (p.tma_out ? launch_fprop<BN, RB, F, kEpiTma, true>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st)                        \
                        : launch_fprop<BN, RB, F, kEpiLegacy, true>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st))                    \
           : (p.tma_out ? launch_fprop<BN, RB, F, kEpiTma>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st)                              \
                        : launch_fprop<BN, RB, F>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st))
#define QEB_FPROP_CASE_NOWIN(BN, RB, F)                                                                                                \
  rc = p.tma_out ? launch_fprop<BN, RB, F, kEpiTma>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st)                                     \
                 : launch_fprop<BN, RB, F>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st)
  if (ep.log_softmax) {
    rc = f16 ? launch_fprop<128, 128, true, kEpiLsm>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st)
             : launch_fprop<128, 128, false, kEpiLsm>(ta_in, tb, to, p, m_tiles, n_tiles, splits, st);
  } else if (!f16) {
    switch (bn) {
      case 32: QEB_FPROP_CASE(32, 128, false); break;
      case 64: QEB_FPROP_CASE(64, 128, false); break;
      case 128: QEB_FPROP_CASE(128, 128, false); break;
      default: QEB_FPROP_CASE_NOWIN(256, 128, false); break;
    }
  } else if (kblk == 64) {
    switch (bn) {
      case 32: QEB_FPROP_CASE(32, 128, true); break;
      case 64: QEB_FPROP_CASE(64, 128, true); break;
      case 128: QEB_FPROP_CASE(128, 128, true); break;
      default: QEB_FPROP_CASE_NOWIN(256, 128, true); break;
    }
  } else {
    switch (bn) {
      case 32: QEB_FPROP_CASE(32, 64, true); break;
      case 64: QEB_FPROP_CASE(64, 64, true); break;
      case 128: QEB_FPROP_CASE(128, 64, true); break;
      default: QEB_FPROP_CASE_NOWIN(256, 64, true); break;
    }
  }
#undef QEB_FPROP_CASE
#undef QEB_FPROP_CASE_NOWIN
  if (rc == QEB_OK && ep.bn_stats && !stats_fused) rc = bn_train_stats(out, ep.bn_stats, st);   // separate pass over the output
  return rc;
}

void fprop_box(int h_out, int w_out, uint32_t* box) {
  const int wt = min(pow2_ceil(w_out), kBlockM);
  const int ht = min(pow2_ceil(h_out), kBlockM / wt);
  box[0] = kBlockK; box[1] = wt; box[2] = ht; box[3] = kBlockM / (wt * ht);
}

bool strides_ok(const Img& a) { return ((uintptr_t)a.p & 15) == 0 && a.sn % 4 == 0 && a.sh % 4 == 0 && a.sw % 4 == 0; }

}  // namespace

int tc_conv_fprop(const Img& x, const float* wpacked, int n_total, int kh, int kw, int ph, int pw, const Img& out,
                  const TcEpilogue& ep, cudaStream_t st) {
  QEB_REQUIRE(x.p && wpacked && out.p, "tc_conv_fprop: null pointer");
  QEB_REQUIRE(x.c > 0 && x.c % 32 == 0, "tc_conv_fprop: cin=%d must be a positive multiple of 32", x.c);
  QEB_REQUIRE(strides_ok(x), "tc_conv_fprop: input must be 16-byte aligned with strides that are multiples of 4");
  QEB_REQUIRE(out.n == x.n && out.h == x.h + 2 * ph - kh + 1 && out.w == x.w + 2 * pw - kw + 1,
              "tc_conv_fprop: output geometry %dx%dx%d does not match input %dx%dx%d k%dx%d p%d,%d", out.n, out.h, out.w, x.n,
              x.h, x.w, kh, kw, ph, pw);
  QEB_REQUIRE(n_total > 0 && n_total <= out.c, "tc_conv_fprop: n_total %d vs out.c %d", n_total, out.c);
  TmapArray4 ta;
  uint32_t box[4];
  fprop_box(out.h, out.w, box);
  const bool f16 = ep.in16 && ep.w16 && strides_ok16(x);
  if (f16) box[0] = kblk16(x.c);
  const bool win = !ep.log_softmax && win_eligible(kh, kw, ph, pw, x.n, out.h, out.w, x.c, (int)box[0], (int)box[0] * (f16 ? 2 : 4));
  if (win) { box[1] = kWinW + 2; box[2] = kWinH + 2; box[3] = 1; }   // the tile with its halo
  int rc = f16 ? tmap_img16(&ta.m[0], x, ep.in16, box) : tmap_img(&ta.m[0], x, x.p, x.c, x.sn, x.sh, x.sw, x.w, x.h, box, 0);
  if (rc) return rc;
  // the A box halved along its outermost dimension of extent > 1 (rows 0-63 / 64-127 of the tile): the operand of the multicast form
  int mc_dim = 0, mc_half = 0;
  if (!win) {
    uint32_t hb[4] = {box[0], box[1], box[2], box[3]};
    mc_dim = box[3] > 1 ? 3 : (box[2] > 1 ? 2 : 1);
    hb[mc_dim] /= 2;
    mc_half = (int)hb[mc_dim];
    rc = f16 ? tmap_img16(&ta.m[1], x, ep.in16, hb) : tmap_img(&ta.m[1], x, x.p, x.c, x.sn, x.sh, x.sw, x.w, x.h, hb, 0);
    if (rc) return rc;
  }
  return fprop_common(ta, false, x, wpacked, n_total, kh, kw, ph, pw, x.c, out, out.h, out.w, ep, 0, 0, ep.bias, st, f16, win, mc_dim, mc_half);
}

int tc_convT_fprop(const Img& x, const float* wpacked, const float* bias, const Img& out, cudaStream_t st,
                   const TcEpilogue* shadows) {
  QEB_REQUIRE(x.p && wpacked && out.p, "tc_convT_fprop: null pointer");
  QEB_REQUIRE(x.c % 32 == 0 && out.c % 32 == 0, "tc_convT_fprop: channels must be multiples of 32");
  QEB_REQUIRE(strides_ok(x), "tc_convT_fprop: input alignment");
  QEB_REQUIRE(out.n == x.n && out.h == 2 * x.h && out.w == 2 * x.w, "tc_convT_fprop: output must be 2x the input");
  TmapArray4 ta;
  uint32_t box[4];
  fprop_box(x.h, x.w, box);
  TcEpilogue ep;
  if (shadows) { ep.in16 = shadows->in16; ep.w16 = shadows->w16; ep.out16 = shadows->out16; ep.round_out = shadows->round_out; }
  const bool f16 = ep.in16 && ep.w16 && strides_ok16(x);
  if (f16) box[0] = kblk16(x.c);
  int rc = f16 ? tmap_img16(&ta.m[0], x, ep.in16, box) : tmap_img(&ta.m[0], x, x.p, x.c, x.sn, x.sh, x.sw, x.w, x.h, box, 0);
  if (rc) return rc;
  return fprop_common(ta, false, x, wpacked, 4 * out.c, 1, 1, 0, 0, x.c, out, x.h, x.w, ep, 1, out.c, bias, st, f16);
}

int tc_convT_dgrad(const Img& dy, const float* wpacked, const Img& dx, const TcEpilogue& ep, cudaStream_t st) {
  QEB_REQUIRE(dy.p && wpacked && dx.p, "tc_convT_dgrad: null pointer");
  QEB_REQUIRE(dy.c % 32 == 0, "tc_convT_dgrad: channels must be multiples of 32");
  QEB_REQUIRE(strides_ok(dy), "tc_convT_dgrad: input alignment");
  QEB_REQUIRE(dy.n == dx.n && dy.h == 2 * dx.h && dy.w == 2 * dx.w, "tc_convT_dgrad: dy must be 2x dx");
  TmapArray4 ta;
  uint32_t box[4];
  fprop_box(dx.h, dx.w, box);
  const bool f16 = ep.in16 && ep.w16 && strides_ok16(dy) && ((uintptr_t)ep.in16 & 15) == 0;
  if (f16) box[0] = kblk16(dy.c);
  for (int t = 0; t < 4; ++t) {  // sub-lattice (dh,dw) of dy viewed as a dx-sized image
    const int dh = t >> 1, dw = t & 1;
    int rc;
    if (f16) {
      Img sub = dy;   // geometry of the sub-lattice: every second row / pixel
      sub.w = dx.w; sub.h = dx.h; sub.sh = 2 * dy.sh; sub.sw = 2 * dy.sw;
      rc = tmap_img16(&ta.m[t], sub, static_cast<const __half*>(ep.in16) + dh * dy.sh + dw * dy.sw, box);
    } else {
      rc = tmap_img(&ta.m[t], dy, dy.p + dh * dy.sh + dw * dy.sw, dy.c, dy.sn, 2 * dy.sh, 2 * dy.sw, dx.w, dx.h, box, 0);
    }
    if (rc) return rc;
  }
  return fprop_common(ta, true, dx, wpacked, dx.c, 2, 2, 0, 0, dy.c, dx, dx.h, dx.w, ep, 0, 0, ep.bias, st, f16);
}

// C ABI (tests and external callers): contiguous NHWC with channel strides.
QEB_API int qeb_conv_fprop_tc(const float* x, int n_img, int h_in, int w_in, int cin, int x_cstride, const float* wpacked,
                              int n_total, int kh, int kw, int ph, int pw, const float* bias, const float* scale, int relu,
                              float* out, int out_cstride, int accumulate, void* stream) {
  QEB_REQUIRE(h_in + 2 * ph - kh + 1 > 0 && w_in + 2 * pw - kw + 1 > 0, "conv_fprop_tc: empty output");
  Img xi = img_nhwc(const_cast<float*>(x), n_img, h_in, w_in, cin, x_cstride);
  Img oi = img_nhwc(out, n_img, h_in + 2 * ph - kh + 1, w_in + 2 * pw - kw + 1, n_total, out_cstride);
  TcEpilogue ep;
  ep.bias = bias; ep.scale = scale; ep.relu = relu; ep.accumulate = accumulate;
  return tc_conv_fprop(xi, wpacked, n_total, kh, kw, ph, pw, oi, ep, (cudaStream_t)stream);
}

QEB_API int qeb_convT2x2_fprop_tc(const float* x, int n_img, int h, int w, int cin, int x_cstride, const float* wpacked,
                                  const float* bias, int cout, float* out, int out_cstride, void* stream) {
  Img xi = img_nhwc(const_cast<float*>(x), n_img, h, w, cin, x_cstride);
  Img oi = img_nhwc(out, n_img, 2 * h, 2 * w, cout, out_cstride);
  return tc_convT_fprop(xi, wpacked, bias, oi, (cudaStream_t)stream);
}

QEB_API int qeb_convT2x2_dgrad_tc(const float* dy, int n_img, int h, int w, int cout, int dy_cstride, const float* wpacked,
                                  int cin, float* dx, int dx_cstride, void* stream) {
  Img di = img_nhwc(const_cast<float*>(dy), n_img, 2 * h, 2 * w, cout, dy_cstride);
  Img xi = img_nhwc(dx, n_img, h, w, cin, dx_cstride);
  TcEpilogue ep;
  return tc_convT_dgrad(di, wpacked, xi, ep, (cudaStream_t)stream);
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
constexpr int kBoxBytes = kWgPix * 32 * 4;  // 4 KB: [32 px][32 ch] tf32

// ROWB = 0: tf32 operands (fp32 tensors, boxes of 32 channels, SWIZZLE_128B_BASE32B). ROWB = 128 / 64: fp16 operand shadows
// in MN-major SWIZZLE_128B / SWIZZLE_64B tiles - a box is [32 px][64 or 32 ch] and an MMA consumes 16 pixels. With 64-channel
// rows the same 32 pixels of a 128 x BLOCK_N tile are (2 + BLOCK_N / 64) x 32 operand rows instead of (4 + BLOCK_N / 32) x 32: the
// kernel is bound by TMA rows (~3.4 cycles each), so the fp16 form does the same contraction in about half the time.
template <int ROWB>
struct WgCfg {
  static constexpr bool kF16 = ROWB != 0;
  static constexpr int kRowBytes = kF16 ? ROWB : 128;
  static constexpr int kCh = kF16 ? ROWB / 2 : 32;            // channels per box = per operand row
  static constexpr int kBox = kWgPix * kRowBytes;             // bytes of one box
  static constexpr int kABoxes = kBlockM / kCh;
  static constexpr int kASide = kABoxes * kBox;
};

struct WgradParams {
  int n_img, h_out, w_out;
  int wt, ht, nt;  // pixel box shape, wt*ht*nt == 32
  int tiles_w, tiles_h, tiles_total;
  int kh, kw, ph, pw;
  int a_groups;       // channel groups (one box wide) on the A side per tap
  int row_blocks;     // taps * a_groups
  int n_total;        // B-side channels
  int per_split;      // pixel tiles per gridDim.z slice
  int stages;         // depth of the shared-memory ring
  int a_map_per_tap;  // 1: tap selects the A tensor map (ConvTranspose sub-lattices), no coordinate shift
  float* out;
  long long s_rowc, s_kh, s_kw, s_col;  // element strides of the gradient tensor
  long long* timeline;  // debugging aid (qeb_debug_set_timeline), NULL in production
  const float* alpha;   // device scalar or NULL: the accumulator is multiplied by *alpha (1 / scale of a scaled fp16 operand)
};

template <int BLOCK_N, int ROWB>
__global__ void __launch_bounds__(kThreads, FpropCfg<BLOCK_N>::kCtasPerSm)
conv_wgrad_tc_kernel(const __grid_constant__ TmapArray4 tmaps_a, const __grid_constant__ CUtensorMap tmap_b,
                     const WgradParams p) {
  using Cfg = FpropCfg<BLOCK_N>;  // stage geometry of the tf32 form (16 KB A side + BLOCK_N*128 B side); the fp16 forms use half
  using W = WgCfg<ROWB>;
  constexpr int kBBoxes = BLOCK_N / W::kCh;
  static_assert(kBBoxes >= 1, "tile narrower than a box");
  constexpr int kStage = W::kF16 ? W::kASide + kBBoxes * W::kBox : Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * kStage);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full_bar = empty_bar + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_m = blockIdx.x, tile_n = blockIdx.y;
  const int t_begin = blockIdx.z * p.per_split;
  const int t_end = min(t_begin + p.per_split, p.tiles_total);
  const int rb0 = tile_m * W::kABoxes;
  const int n_rb = min(W::kABoxes, p.row_blocks - rb0);
  long long* tl = p.timeline ? p.timeline + 16 * ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
  if (tl && threadIdx.x == 0) { tl[0] = clock64(); unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); tl[7] = sm; }

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < (p.a_map_per_tap ? 4 : 1); ++i) prefetch_tmap(&tmaps_a.m[i]);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
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
  qeb_pdl_sync();
  if (tl && threadIdx.x == 0) tl[1] = clock64();

  if (t_begin < t_end) {
    if (warp == 0) {
      // ===== TMA producer: up to kABoxes A boxes + kBBoxes B boxes per stage. The per-box constants (tap shift, channel
      // offset, tensor map) are computed once and the tile coordinates are carried as counters, so the steady-state loop
      // has no integer division; control flow is warp-uniform and one elected lane issues, which keeps all of it in
      // uniform registers (see conv_fprop_tc_kernel) =====
      int ac0[W::kABoxes], asx[W::kABoxes], asy[W::kABoxes], amap[W::kABoxes];
#pragma unroll
      for (int j = 0; j < W::kABoxes; ++j) {
        const int rb = rb0 + (j < n_rb ? j : 0);
        const int tap = rb / p.a_groups, cg = rb - tap * p.a_groups;
        ac0[j] = cg * W::kCh;
        amap[j] = p.a_map_per_tap ? tap : 0;
        asy[j] = p.a_map_per_tap ? 0 : tap / p.kw - p.ph;
        asx[j] = p.a_map_per_tap ? 0 : tap % p.kw - p.pw;
      }
      int tw = t_begin % p.tiles_w, th = (t_begin / p.tiles_w) % p.tiles_h, tn = t_begin / (p.tiles_w * p.tiles_h);
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)(n_rb + kBBoxes) * W::kBox;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sbase = smem + stage * kStage;
          mbar_expect_tx(&full_bar[stage], bytes);
#pragma unroll
          for (int j = 0; j < W::kABoxes; ++j)
            if (j < n_rb)
              tma_load_4d(sbase + j * W::kBox, &tmaps_a.m[amap[j]], &full_bar[stage], ac0[j], tw * p.wt + asx[j], th * p.ht + asy[j],
                          tn * p.nt);
#pragma unroll
          for (int j = 0; j < kBBoxes; ++j)
            tma_load_4d(sbase + W::kASide + j * W::kBox, &tmap_b, &full_bar[stage], tile_n * BLOCK_N + j * W::kCh, tw * p.wt, th * p.ht,
                        tn * p.nt);
        }
        __syncwarp();
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) { th = 0; ++tn; }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1) {
      // MMA issuer: warp-uniform loop, elected lane issues (see conv_fprop_tc_kernel)
      constexpr uint32_t idesc = W::kF16 ? instr_desc_f16(kBlockM, BLOCK_N, 0, 0, 1, 1) : instr_desc_tf32(kBlockM, BLOCK_N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (tl && lane == 0 && t == t_begin) tl[2] = clock64();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * kStage);
          const uint32_t sb = sa + W::kASide;
          if constexpr (W::kF16) {
#pragma unroll
            for (int k = 0; k < kWgPix / 16; ++k) {
              // 16 pixels (two 8-row K groups) per MMA: +16 rows inside every [32 px][kCh ch] box
              const uint64_t adesc = smem_desc_mnmajor_16<W::kRowBytes>(sa + k * 16 * W::kRowBytes, W::kBox, 8 * W::kRowBytes);
              const uint64_t bdesc = smem_desc_mnmajor_16<W::kRowBytes>(sb + k * 16 * W::kRowBytes, W::kBox, 8 * W::kRowBytes);
              mma_f16_ss(tmem_base, adesc, bdesc, idesc, (t > t_begin) || (k != 0));
            }
          } else {
#pragma unroll
            for (int k = 0; k < kWgPix / 8; ++k) {
              // 8 pixels (one swizzle atom of K) per MMA: +1024 B inside every [32 px][32 ch] block
              const uint64_t adesc = smem_desc_mnmajor_sw128_32b(sa + k * 1024, kBoxBytes, 512);
              const uint64_t bdesc = smem_desc_mnmajor_sw128_32b(sb + k * 1024, kBoxBytes, 512);
              mma_tf32_ss(tmem_base, adesc, bdesc, idesc, (t > t_begin) || (k != 0));
            }
          }
          mma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) mma_commit(tmem_full_bar);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int rb = rb0 + r / W::kCh;  // row block of this TMEM lane
      const bool valid = rb < p.row_blocks;
      const int tap = valid ? rb / p.a_groups : 0, cg = valid ? rb - tap * p.a_groups : 0;
      float* row_out = p.out + (long long)(cg * W::kCh + (r % W::kCh)) * p.s_rowc + (long long)(tap / p.kw) * p.s_kh +
                       (long long)(tap % p.kw) * p.s_kw;
      const float alpha = p.alpha ? __ldg(p.alpha) : 1.f;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      if (tl && threadIdx.x == 64) tl[4] = clock64();
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        const int ncol = tile_n * BLOCK_N + c0;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ncol + j < p.n_total) atomicAdd(row_out + (long long)(ncol + j) * p.s_col, v[j] * alpha);
        }
      }
      if (tl && threadIdx.x == 64) tl[5] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
  if (tl && threadIdx.x == 32) { tl[6] = clock64(); tl[3] = t_end - t_begin; }
}

template <int BLOCK_N, int ROWB>
int launch_wgrad(const TmapArray4& ta, const CUtensorMap& tb, const WgradParams& p_in, dim3 grid, cudaStream_t st) {
  using Cfg = FpropCfg<BLOCK_N>;
  using W = WgCfg<ROWB>;
  static bool attr = false;
  if (!attr) {
    QEB_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BLOCK_N, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kMaxSmem));
    attr = true;
  }
  WgradParams p = p_in;
  const int stage_bytes = W::kF16 ? W::kASide + (BLOCK_N / W::kCh) * W::kBox : Cfg::kStageBytes;
  const long long ctas = (long long)grid.x * grid.y * grid.z;
  p.stages = (Cfg::kSmBudget / Cfg::resident(ctas) - 1280) / stage_bytes;
  p.stages = max(2, min(p.stages, W::kF16 ? 12 : Cfg::kMaxStages));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + kBarBytes;
  ProfScope prof(W::kF16 ? "tc_conv_wgrad.f16" : "tc_conv_wgrad.tf32", st, 2.0 * p.n_img * p.h_out * p.w_out * (double)p.n_total * p.row_blocks * W::kCh,
                 (W::kF16 ? 2.0 : 4.0) * ((double)p.n_img * p.h_out * p.w_out * (p.a_groups * W::kCh + p.n_total)) +
                     4.0 * (double)p.n_total * p.row_blocks * W::kCh);
  QEB_CUDA(qeb_launch(conv_wgrad_tc_kernel<BLOCK_N, ROWB>, grid, kThreads, smem, st, ta, tb, p));
  qeb_count_launch();
  return QEB_OK;
}

}  // namespace

namespace {

// rowb: 0 = tf32 operands, 128 / 64 = fp16 operand shadows in rows of that many bytes
int wgrad_common(const TmapArray4& ta, const CUtensorMap& tb, WgradParams& p, int a_c, int b_c, cudaStream_t st, int rowb = 0) {
  int bn = min(256, max(rowb == 128 ? 64 : 32, pow2_ceil(b_c)));
  const int a_boxes = rowb == 128 ? 2 : 4;
  const int m_tiles = qeb_cdiv(p.row_blocks, a_boxes), n_tiles = qeb_cdiv(b_c, bn);
  // split-K so that the grid is about ONE CTA per SM, at least 8 pixel tiles per CTA: every CTA ends with a red.global.add
  // epilogue that retires at ~13 B per clock and SM (10 k cycles for a 128 x 256 tile, scripts/exp/wgrad_timeline.py), so a
  // second wave of CTAs pays setup + epilogue twice (same-box A/B: 296 CTAs 3.68 ms per step, 148 CTAs 3.65 ms)
  // CTA slots the cost model below fills per round: 168 (two CTAs per SM are resident for BLOCK_N < 256; same-box A/B of the step
  // 148 -> 3.007 ms, 160-176 -> 2.978-2.991 ms, 192 -> 2.999, 222 -> 3.004)
  static const int wg_ctas = getenv("QEB_WG_CTAS") ? atoi(getenv("QEB_WG_CTAS")) : 168;
  static const int wg_model = getenv("QEB_WG_SPLIT_MODEL") ? atoi(getenv("QEB_WG_SPLIT_MODEL")) : 1;
  int splits = qeb_cdiv(wg_ctas, m_tiles * n_tiles);
  splits = max(1, min(splits, qeb_cdiv(p.tiles_total, 8)));
  if (wg_model) {
    // The rounded-up split count above overshoots the SM count (216, 180, 162, 150 CTAs ...): the kernel then lasts as long as
    // the SMs that got TWO CTAs. Pick the split count from a cost model instead: rounds of CTAs per SM x (K steps x ~600 cycles
    // + the reduction epilogue, ~40 cycles per accumulator column), minimised over the candidates.
    const int tiles = m_tiles * n_tiles, max_splits = max(1, p.tiles_total / 4);
    long long best = -1;
    for (int s = 1; s <= max_splits && s <= 4 * wg_ctas; ++s) {
      const long long rounds = qeb_cdiv((long long)tiles * s, wg_ctas);
      static const int c_k = getenv("QEB_WG_CK") ? atoi(getenv("QEB_WG_CK")) : 600;
      static const int c_e = getenv("QEB_WG_CE") ? atoi(getenv("QEB_WG_CE")) : 40;
      const long long cost = rounds * ((long long)qeb_cdiv(p.tiles_total, s) * c_k + (long long)c_e * bn + 3000);
      if (best < 0 || cost < best) { best = cost; splits = s; }
    }
  }
  p.per_split = qeb_cdiv(p.tiles_total, splits);
  splits = qeb_cdiv(p.tiles_total, p.per_split);
  const dim3 grid(m_tiles, n_tiles, splits);
#define QEB_WG_CASE(BN)                                                                       \
  return rowb == 128 ? launch_wgrad<(BN < 64 ? 64 : BN), 128>(ta, tb, p, grid, st)            \
                     : (rowb == 64 ? launch_wgrad<BN, 64>(ta, tb, p, grid, st) : launch_wgrad<BN, 0>(ta, tb, p, grid, st))
  switch (bn) {
    case 32: QEB_WG_CASE(32);
    case 64: QEB_WG_CASE(64);
    case 128: QEB_WG_CASE(128);
    default: QEB_WG_CASE(256);
  }
#undef QEB_WG_CASE
}

void wgrad_geometry(WgradParams& p, int n_img, int h_out, int w_out, uint32_t* box) {
  p.n_img = n_img; p.h_out = h_out; p.w_out = w_out;
  p.wt = min(pow2_ceil(w_out), kWgPix);
  p.ht = min(pow2_ceil(h_out), kWgPix / p.wt);
  p.nt = kWgPix / (p.wt * p.ht);
  p.tiles_w = qeb_cdiv(w_out, p.wt);
  p.tiles_h = qeb_cdiv(h_out, p.ht);
  p.tiles_total = p.tiles_w * p.tiles_h * qeb_cdiv(n_img, p.nt);
  box[0] = 32; box[1] = p.wt; box[2] = p.ht; box[3] = p.nt;
}

// fp16 operand shadows usable: both given, strides 16-byte friendly; returns the row bytes (128 when both channel counts are
// multiples of 64, else 64), 0 = stay on tf32
int wgrad_rowb16(const Img& x, const Img& dy, const WgradShadows* sh) {
  static const int allow = getenv("QEB_FP16_BWD") ? atoi(getenv("QEB_FP16_BWD")) : 1;
  if (!allow || !sh || !sh->x16 || !sh->dy16 || !strides_ok16(x) || !strides_ok16(dy)) return 0;
  if (((uintptr_t)sh->x16 & 15) || ((uintptr_t)sh->dy16 & 15)) return 0;
  return (x.c % 64 == 0 && dy.c % 64 == 0) ? 128 : 64;
}

}  // namespace

int tc_conv_wgrad(const Img& x, const Img& dy, int kh, int kw, int ph, int pw, float* dw, long long s_co, long long s_ci,
                  long long s_kh, long long s_kw, cudaStream_t st, const WgradShadows* sh) {
  QEB_REQUIRE(x.p && dy.p && dw, "tc_conv_wgrad: null pointer");
  QEB_REQUIRE(x.c > 0 && x.c % 32 == 0, "tc_conv_wgrad: input channels %d must be a multiple of 32", x.c);
  QEB_REQUIRE(strides_ok(x) && strides_ok(dy), "tc_conv_wgrad: operands must be 16-byte aligned, strides multiples of 4");
  QEB_REQUIRE(x.n == dy.n, "tc_conv_wgrad: batch mismatch");
  WgradParams p;
  p.timeline = g_timeline;
  uint32_t box[4];
  wgrad_geometry(p, dy.n, dy.h, dy.w, box);
  p.kh = kh; p.kw = kw; p.ph = ph; p.pw = pw;
  const int rowb = wgrad_rowb16(x, dy, sh);
  const int ch = rowb ? rowb / 2 : 32;   // channels per operand row / box
  p.a_groups = x.c / ch;
  p.row_blocks = kh * kw * p.a_groups;
  p.n_total = dy.c;
  p.a_map_per_tap = 0;
  p.out = dw;
  p.s_rowc = s_ci; p.s_kh = s_kh; p.s_kw = s_kw; p.s_col = s_co;
  p.alpha = rowb ? sh->alpha : nullptr;
  TmapArray4 ta;
  CUtensorMap tb;
  int rc;
  if (rowb) {
    box[0] = (uint32_t)ch;
    rc = tmap_img16(&ta.m[0], x, sh->x16, box);
    if (rc) return rc;
    rc = tmap_img16(&tb, dy, sh->dy16, box);
  } else {
    rc = tmap_img(&ta.m[0], x, x.p, x.c, x.sn, x.sh, x.sw, x.w, x.h, box, 1);
    if (rc) return rc;
    rc = tmap_img(&tb, dy, dy.p, dy.c, dy.sn, dy.sh, dy.sw, dy.w, dy.h, box, 1);
  }
  if (rc) return rc;
  ta.m[1] = ta.m[2] = ta.m[3] = ta.m[0];
  return wgrad_common(ta, tb, p, x.c, dy.c, st, rowb);
}

int tc_convT_wgrad(const Img& x, const Img& dy, float* dw, cudaStream_t st, const WgradShadows* sh) {
  QEB_REQUIRE(x.p && dy.p && dw, "tc_convT_wgrad: null pointer");
  QEB_REQUIRE(x.c % 32 == 0 && dy.c % 32 == 0, "tc_convT_wgrad: channels must be multiples of 32");
  QEB_REQUIRE(strides_ok(x) && strides_ok(dy), "tc_convT_wgrad: operand alignment");
  QEB_REQUIRE(x.n == dy.n && dy.h == 2 * x.h && dy.w == 2 * x.w, "tc_convT_wgrad: dy must be 2x x");
  // M side: (tap, co) from the four sub-lattices of dy; N side: ci. dw[ci][co][dh][dw].
  WgradParams p;
  p.timeline = g_timeline;
  uint32_t box[4];
  wgrad_geometry(p, x.n, x.h, x.w, box);
  p.kh = 2; p.kw = 2; p.ph = 0; p.pw = 0;
  const int rowb = wgrad_rowb16(x, dy, sh);
  const int ch = rowb ? rowb / 2 : 32;
  p.a_groups = dy.c / ch;
  p.row_blocks = 4 * p.a_groups;
  p.n_total = x.c;
  p.a_map_per_tap = 1;
  p.out = dw;
  p.alpha = rowb ? sh->alpha : nullptr;
  p.s_rowc = 4; p.s_kh = 2; p.s_kw = 1; p.s_col = (long long)dy.c * 4;
  TmapArray4 ta;
  CUtensorMap tb;
  if (rowb) box[0] = (uint32_t)ch;
  for (int t = 0; t < 4; ++t) {
    const int dh = t >> 1, dwi = t & 1;
    int rc;
    if (rowb) {
      Img sub = dy;
      sub.w = x.w; sub.h = x.h; sub.sh = 2 * dy.sh; sub.sw = 2 * dy.sw;
      rc = tmap_img16(&ta.m[t], sub, static_cast<const __half*>(sh->dy16) + dh * dy.sh + dwi * dy.sw, box);
    } else {
      rc = tmap_img(&ta.m[t], dy, dy.p + dh * dy.sh + dwi * dy.sw, dy.c, dy.sn, 2 * dy.sh, 2 * dy.sw, x.w, x.h, box, 1);
    }
    if (rc) return rc;
  }
  int rc = rowb ? tmap_img16(&tb, x, sh->x16, box) : tmap_img(&tb, x, x.p, x.c, x.sn, x.sh, x.sw, x.w, x.h, box, 1);
  if (rc) return rc;
  return wgrad_common(ta, tb, p, dy.c, x.c, st, rowb);
}

// the same contraction with fp16 operands (x16: NHWC fp16 image, w16: packed fp16 weights) and an optional fp16 shadow of
// the fp32 output
QEB_API int qeb_conv_fprop_tc16(const void* x16, int n_img, int h_in, int w_in, int cin, int x_cstride, const void* w16,
                                int n_total, int kh, int kw, int ph, int pw, const float* bias, const float* scale, int relu,
                                float* out, int out_cstride, void* out16, void* stream) {
  QEB_REQUIRE(x16 && w16, "conv_fprop_tc16: null operand");
  QEB_REQUIRE(h_in + 2 * ph - kh + 1 > 0 && w_in + 2 * pw - kw + 1 > 0, "conv_fprop_tc16: empty output");
  QEB_REQUIRE(x_cstride % 8 == 0, "conv_fprop_tc16: the channel stride must be a multiple of 8 (16-byte rows)");
  Img xi = img_nhwc(reinterpret_cast<float*>(const_cast<void*>(x16)), n_img, h_in, w_in, cin, x_cstride);   // geometry only
  Img oi = img_nhwc(out, n_img, h_in + 2 * ph - kh + 1, w_in + 2 * pw - kw + 1, n_total, out_cstride);
  TcEpilogue ep;
  ep.bias = bias; ep.scale = scale; ep.relu = relu;
  ep.in16 = x16; ep.w16 = w16; ep.out16 = out16;
  return tc_conv_fprop(xi, reinterpret_cast<const float*>(w16), n_total, kh, kw, ph, pw, oi, ep, (cudaStream_t)stream);
}

// C ABI: weight gradient of a stride-1 convolution / Linear in torch's layout dw[co][ci][kh][kw], accumulated.
QEB_API int qeb_conv_wgrad_tc(const float* x, int cin, int x_cstride, int h_in, int w_in, const float* dy, int cout,
                              int dy_cstride, int n_img, int kh, int kw, int ph, int pw, float* dw, void* stream) {
  Img xi = img_nhwc(const_cast<float*>(x), n_img, h_in, w_in, cin, x_cstride);
  Img di = img_nhwc(const_cast<float*>(dy), n_img, h_in + 2 * ph - kh + 1, w_in + 2 * pw - kw + 1, cout, dy_cstride);
  return tc_conv_wgrad(xi, di, kh, kw, ph, pw, dw, (long long)cin * kh * kw, (long long)kh * kw, kw, 1, (cudaStream_t)stream);
}

// the same with fp16 operand shadows (x16 / dy16: same element layout as x / dy; dy16 holds dy * S, alpha_dev -> 1 / S or NULL)
QEB_API int qeb_conv_wgrad_tc16(const float* x, const void* x16, int cin, int x_cstride, int h_in, int w_in, const float* dy,
                                const void* dy16, int cout, int dy_cstride, int n_img, int kh, int kw, int ph, int pw,
                                const float* alpha_dev, float* dw, void* stream) {
  Img xi = img_nhwc(const_cast<float*>(x), n_img, h_in, w_in, cin, x_cstride);
  Img di = img_nhwc(const_cast<float*>(dy), n_img, h_in + 2 * ph - kh + 1, w_in + 2 * pw - kw + 1, cout, dy_cstride);
  WgradShadows sh;
  sh.x16 = x16; sh.dy16 = dy16; sh.alpha = alpha_dev;
  return tc_conv_wgrad(xi, di, kh, kw, ph, pw, dw, (long long)cin * kh * kw, (long long)kh * kw, kw, 1, (cudaStream_t)stream, &sh);
}

QEB_API int qeb_convT2x2_wgrad_tc(const float* x, int cin, int x_cstride, int h, int w, const float* dy, int cout,
                                  int dy_cstride, int n_img, float* dw, void* stream) {
  Img xi = img_nhwc(const_cast<float*>(x), n_img, h, w, cin, x_cstride);
  Img di = img_nhwc(const_cast<float*>(dy), n_img, 2 * h, 2 * w, cout, dy_cstride);
  return tc_convT_wgrad(xi, di, dw, (cudaStream_t)stream);
}

// Debugging aid: when set, every CTA of conv_fprop_tc_kernel writes 8 int64 {t_entry, t_setup_done, t_first_stage_landed,
// t_mma_issued, t_accum_ready, t_stores_done, t_exit, smid, ...} (clock64) at buf[16 * CTA index]. NULL switches it off.
QEB_API void qeb_debug_set_timeline(long long* buf) { g_timeline = buf; }
long long* qeb_debug_timeline() { return g_timeline; }
