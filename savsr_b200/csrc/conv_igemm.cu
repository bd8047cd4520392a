// Batched implicit-GEMM 3x3 / 1x1 convolution on tcgen05 tensor cores (sm_100a).
//
//   D[128 pixels, N] += A[128 pixels, 64 ch] * B[64 ch, N]      per (source, tap) K-block
//
//  * A operand: NHWC bf16 arena, fetched by TMA as a 4-D box (64 ch, x, y, 1 image) straight into the
//    128-byte-swizzled K-major layout tcgen05.mma reads.  Zero padding of the convolution is the TMA
//    out-of-bounds fill.  Two fetch modes:
//      TAP  : one 8x16-pixel box (16 KB) per filter tap, box origin shifted by the tap.
//      HALO : one (8+2) x (16+2) halo box per source; the nine taps are nine UMMA descriptors into the
//             same tile (start address shifted by whole 128-byte pixel rows, SBO = halo row pitch).
//             Measured on B200: the swizzle XOR is a function of absolute shared-memory address bits,
//             so shifted starts and an SBO that is not a multiple of 1024 need no base_offset.
//  * B operand (weights): STATIONARY.  All CTAs of a launch need the same few K-blocks, and streaming
//    them per tile from L2 makes 148 SMs hammer the same 64 lines (measured: ~1000 cycles per K-block
//    vs 128 cycles of MMA).  So each persistent CTA works on a contiguous chunk of tiles and keeps up to
//    18 pre-swizzled K-blocks (144 KB) resident in shared memory, reloading only when the (conv, sample)
//    weight set changes; K-blocks beyond that are streamed through a small ring.  Loads are issued in a
//    per-CTA rotated order so that concurrent CTAs pull different L2 lines.
//  * accumulator: fp32 in TMEM, double buffered (2 x N columns) so the epilogue of tile i overlaps the
//    main loop of tile i+1.
//  * warp roles: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one thread, owns TMEM),
//    warps 2-9 = epilogue: each warp owns a 32-row TMEM lane quadrant x 32 columns
//    (tcgen05.ld -> bias/activation/mask/residuals -> bf16 NHWC store, optional pooled partial sums).
//
// The CUDA-core kernel at the bottom implements the same contract with plain loads and shares the
// epilogue; it exists only as an on-device checker for the TMA/UMMA main loop (tests, bring-up).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace savsr {

// Cycle counters around the barrier waits are a bring-up aid: compiled in only with -DSAVSR_DEBUG_COUNTERS (the release
// kernels execute no clock reads on the issue path).
#ifdef SAVSR_DEBUG_COUNTERS
#define DBG_CLOCK() clock64()
#define DBG_ONLY(...) __VA_ARGS__
#else
#define DBG_CLOCK() 0ll
#define DBG_ONLY(...)
#endif

constexpr int kMaxAStages = 4;
constexpr int kHaloPitch = 10;                        // halo row pitch in pixels (tile width 8 + 2)
constexpr int kHaloStageBytes = 23552;               // 10 x 18 x 128 = 23040, rounded up to 1 KB
constexpr int kARegionBytes = 3 * kHaloStageBytes;   // 3 halo stages | 4 tap stages (4 x 16 KB)
constexpr int kBBlocks = 18;                         // resident weight K-blocks per CTA (N=64: 144 KB)
constexpr int kRing = 3;                             // ring stages carved out of the B region when K is larger
constexpr int kEpiWarps = 8;
constexpr int kNumThreads = 64 + 32 * kEpiWarps;

struct ConvParams {
  CUtensorMap tm_tile;
  CUtensorMap tm_halo;
  savsr_conv_group g[SAVSR_MAX_GROUPS];
  __nv_bfloat16* arena;
  int ngroups, batch, height, width, tiles_x, tiles_y;
  int ntaps;            // 1 or 9
  int nsrc;             // identical for all groups of a launch
  int halo;             // A fetch mode
  int n_res;            // K-blocks held resident; the remaining nsrc*ntaps - n_res go through the ring
  int chunk;            // consecutive work items per CTA
  uint32_t per_sample_mask;  // bit g set: conv g has per-sample weights (OSA-Conv)
  long long* dbg;            // optional [grid][8] cycle counters (bring-up aid), nullptr in production
  int dst_mode;
  int fmt;              // enum savsr_format of the arena and the packed weights
  int issuers;          // batched kernel: MMA-issuing warps, 2 (default) or 1 (bring-up knob SAVSR_BIGK_ISSUERS)
  int ksteps;           // K = 16 steps per tap that carry data: 4 = all 64 channels of a source; fewer when the groups declare src_channels
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == SAVSR_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == SAVSR_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Epilogue shared by the tensor-core kernel and the checker, in two phases so that every global load that does
// not depend on the accumulator (bias, residuals, mask, bilinear skip) is issued BEFORE waiting for the MMAs of the
// tile and its latency hides behind them.  `v` holds NC consecutive accumulator columns (output channels
// col0 .. col0+NC) of pixel m = quad * 32 + lane of the tile.
// NC = 32: bf16 arena destination (two warps per quadrant); NC = 16: AUX16 destination.
template <int NC>
struct EpiCtx {
  float bias[NC];      // reloaded only when the conv (group) changes
  uint4 r1[NC / 8], r2[NC / 8];
  float mk;
  long pix;
  bool valid;
  int bias_group;
  // per-conv constants hoisted out of the per-element code (reloaded only when the conv changes)
  float neg_slope;     // activation as max(t, neg_slope * t), 0 <= neg_slope <= 1: 1 = none, 0 = ReLU, s = LeakyReLU(s)
  float res2_scale;
  int res1_slot, res2_slot, dst_slot;
  bool has_mask;
  float* pool;
};

template <int NC>
__device__ __forceinline__ void epi_prefetch(const ConvParams& p, const savsr_conv_group& g, int gi, int n, int tile, int quad,
                                             int lane, int col0, EpiCtx<NC>& c) {
  const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int m = quad * 32 + lane;
  const int px = tx * kTileW + (m & (kTileW - 1));
  const int py = ty * kTileH + (m >> 3);
  c.valid = (px < p.width) && (py < p.height);
  const long npix = static_cast<long>(p.height) * p.width;
  c.pix = static_cast<long>(py) * p.width + px;
  if (gi != c.bias_group) {
    c.bias_group = gi;
    c.neg_slope = g.act == SAVSR_ACT_NONE ? 1.f : (g.act == SAVSR_ACT_RELU ? 0.f : g.slope);
    c.res2_scale = g.res2_scale;
    c.res1_slot = g.res1_slot; c.res2_slot = g.res2_slot; c.dst_slot = g.dst_slot;
    c.has_mask = g.mask != nullptr;
    c.pool = g.pool;
    if (g.bias != nullptr) {
      const float4* b4 = reinterpret_cast<const float4*>(g.bias + col0);
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        const float4 b = __ldg(b4 + j);
        c.bias[4 * j + 0] = b.x; c.bias[4 * j + 1] = b.y; c.bias[4 * j + 2] = b.z; c.bias[4 * j + 3] = b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) c.bias[j] = 0.f;
    }
  }
  if constexpr (NC == 32) {
    c.mk = (c.has_mask && c.valid) ? __ldg(g.mask + n * npix + c.pix) : 0.f;
    if (c.res1_slot >= 0 && c.valid) {
      const uint4* r = reinterpret_cast<const uint4*>(p.arena + ((static_cast<long>(c.res1_slot) * p.batch + n) * npix + c.pix) * kC + col0);
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) c.r1[j] = r[j];
    }
    if (c.res2_slot >= 0 && c.valid) {
      const uint4* r = reinterpret_cast<const uint4*>(p.arena + ((static_cast<long>(c.res2_slot) * p.batch + n) * npix + c.pix) * kC + col0);
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) c.r2[j] = r[j];
    }
  }
}

__device__ __forceinline__ void add_h16x8(float* v, const uint4& u, float s, int fmt) {
  v[0] += s * h_lo(u.x, fmt); v[1] += s * h_hi(u.x, fmt); v[2] += s * h_lo(u.y, fmt); v[3] += s * h_hi(u.y, fmt);
  v[4] += s * h_lo(u.z, fmt); v[5] += s * h_hi(u.z, fmt); v[6] += s * h_lo(u.w, fmt); v[7] += s * h_hi(u.w, fmt);
}

template <int NC>
__device__ __forceinline__ void epi_finish(const ConvParams& p, const savsr_conv_group& g, int n, int tile, int quad, int lane,
                                           float (&v)[NC], int col0, const EpiCtx<NC>& c) {
  const long npix = static_cast<long>(p.height) * p.width;
  const float ns = c.neg_slope;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const float t = v[j] + c.bias[j];
    v[j] = fmaxf(t, ns * t);   // none / ReLU / LeakyReLU without branches (slope in [0, 1]); two instructions
  }
  if constexpr (NC == 32) {
    if (c.has_mask) {
#pragma unroll
      for (int j = 0; j < NC; ++j) v[j] *= c.mk;
    }
    if (c.res1_slot >= 0 && c.valid) {
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) add_h16x8(v + 8 * j, c.r1[j], 1.f, p.fmt);
    }
    if (c.res2_slot >= 0 && c.valid) {
      const float r2s = c.res2_scale;
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) add_h16x8(v + 8 * j, c.r2[j], r2s, p.fmt);
    }
    if (c.valid) {
      uint4* d = reinterpret_cast<uint4*>(p.arena + ((static_cast<long>(c.dst_slot) * p.batch + n) * npix + c.pix) * kC + col0);
#pragma unroll
      for (int j = 0; j < NC / 8; ++j) {
        uint4 u;
        u.x = pack_h2(v[8 * j + 0], v[8 * j + 1], p.fmt);
        u.y = pack_h2(v[8 * j + 2], v[8 * j + 3], p.fmt);
        u.z = pack_h2(v[8 * j + 4], v[8 * j + 5], p.fmt);
        u.w = pack_h2(v[8 * j + 6], v[8 * j + 7], p.fmt);
        d[j] = u;
      }
    }
    if (c.pool != nullptr) {
      // Per-warp channel sums of the (pre-rounding) outputs over valid pixels, by a shuffle
      // reduce-scatter: 31 shuffles leave channel col0 + lane in each lane.  Deterministic.
      if (!c.valid) {
#pragma unroll
        for (int j = 0; j < NC; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int off = 16, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
          const float send = upper ? v[i] : v[i + cnt];
          const float keep = upper ? v[i + cnt] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      const int npart = p.tiles_x * p.tiles_y * 4;
      c.pool[(static_cast<long>(n) * npart + tile * 4 + quad) * kC + col0 + lane] = v[0];
    }
  } else {
    if (c.valid) {   // SAVSR_DST_AUX16
      float4* d = reinterpret_cast<float4*>(static_cast<float*>(g.aux_dst) + (n * npix + c.pix) * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
}


// ------------------------------------------------------------------------------------------------ N = 64 epilogue
// Epilogue of the tensor-core kernels for the 16-bit arena destination.  Each of the 8 epilogue warps owns one TMEM
// lane quadrant (32 pixels: 4 tile rows of 8) and one 32-column half.  The accumulator is read with 16x256b loads and
// the weights are packed in SAVSR_ROWS_QUAD order, so thread (g = lane / 4, q = lane % 4) holds, for the four pixels
// (x = g, y = 4 quad + j), j = 0..3, the eight consecutive channels half * 32 + 8 q .. + 7: residual loads and stores
// are 16 bytes per thread, 64 contiguous bytes per thread quad, 8 lines per warp instruction.  (The first version
// read one pixel per lane with 32x32b loads; every 16-byte store then touched 32 different lines and the LSU, not the
// tensor core, set the pace of single-source convs.)
struct EpiQuad {
  float bias[8];
  uint4 r1[4], r2[4];
  float mk[4];
  long pix0;          // pixel index of row j = 0; row j is pix0 + j * width
  uint32_t valid;     // bit j: pixel j is inside the image
  int bias_group;
  float neg_slope, res2_scale;
  int res1_slot, res2_slot, dst_slot;
  bool has_mask;
  float* pool;
  float psum[8];      // pooled sums of the current run of tiles (same conv, same sample), reduced across lanes at its end
};

__device__ __forceinline__ void epiq_prefetch(const ConvParams& p, const savsr_conv_group& g, int gi, int n, int tile, int quad,
                                              int lane, int half, EpiQuad& c) {
  const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
  const int px = tx * kTileW + (lane >> 2);
  const int py0 = ty * kTileH + quad * 4;
  const int rows = p.height - py0;   // rows of this quadrant inside the image
  c.valid = (px < p.width && rows > 0) ? (rows >= 4 ? 0xfu : ((1u << rows) - 1u)) : 0u;
  const long npix = static_cast<long>(p.height) * p.width;
  c.pix0 = static_cast<long>(py0) * p.width + px;
  const int ch0 = half * 32 + (lane & 3) * 8;
  if (gi != c.bias_group) {
    c.bias_group = gi;
    c.neg_slope = g.act == SAVSR_ACT_NONE ? 1.f : (g.act == SAVSR_ACT_RELU ? 0.f : g.slope);
    c.res2_scale = g.res2_scale;
    c.res1_slot = g.res1_slot; c.res2_slot = g.res2_slot; c.dst_slot = g.dst_slot;
    c.has_mask = g.mask != nullptr;
    c.pool = g.pool;
    if (g.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + ch0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + ch0) + 1);
      c.bias[0] = b0.x; c.bias[1] = b0.y; c.bias[2] = b0.z; c.bias[3] = b0.w;
      c.bias[4] = b1.x; c.bias[5] = b1.y; c.bias[6] = b1.z; c.bias[7] = b1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) c.bias[j] = 0.f;
    }
  }
  if (c.has_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) c.mk[j] = ((c.valid >> j) & 1u) ? __ldg(g.mask + n * npix + c.pix0 + j * p.width) : 0.f;
  }
  if (c.res1_slot >= 0) {
    const __nv_bfloat16* r = p.arena + ((static_cast<long>(c.res1_slot) * p.batch + n) * npix + c.pix0) * kC + ch0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((c.valid >> j) & 1u) c.r1[j] = *reinterpret_cast<const uint4*>(r + static_cast<long>(j) * p.width * kC);
  }
  if (c.res2_slot >= 0) {
    const __nv_bfloat16* r = p.arena + ((static_cast<long>(c.res2_slot) * p.batch + n) * npix + c.pix0) * kC + ch0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((c.valid >> j) & 1u) c.r2[j] = *reinterpret_cast<const uint4*>(r + static_cast<long>(j) * p.width * kC);
  }
}

// `taddr`: TMEM address of (first lane of the quadrant, first column of the half) of the tile's accumulator.
// Reads the accumulator, calls `release()` (the accumulator may be overwritten from then on), applies bias / activation /
// mask / residuals, then calls `next()` -- the caller starts the NEXT tile's epiq_prefetch into `c` there, so those loads
// have the stores, the pooling and the next accumulator wait to land (when the epilogue is the critical path, as in
// single-source convs, nothing else would hide their latency) -- and finally packs, stores and pools this tile.
// `last_of_run`: the next tile this warp will process belongs to another (conv, sample) or there is none.
template <class Release, class Next>
__device__ __forceinline__ void epiq_finish(const ConvParams& p, int n, int tile, int quad, int lane, int half, uint32_t taddr,
                                            EpiQuad& c, bool last_of_run, Release release, Next next) {
  uint32_t ra[16], rb[16];
  tmem_ld_16x256b_x4(taddr, ra);
  tmem_ld_16x256b_x4(taddr + (16u << 16), rb);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  release();
  // v[j][2 k + e]: pixel row j, channel ch0 + 2 k + e
  float v[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e >> 1][2 * k + (e & 1)] = __uint_as_float(ra[4 * k + e]);
      v[2 + (e >> 1)][2 * k + (e & 1)] = __uint_as_float(rb[4 * k + e]);
    }
  }
  const long npix = static_cast<long>(p.height) * p.width;
  const int ch0 = half * 32 + (lane & 3) * 8;
  const float ns = c.neg_slope;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = v[j][i] + c.bias[i];
      v[j][i] = fmaxf(t, ns * t);   // none / ReLU / LeakyReLU without branches (slope in [0, 1]): the epilogue's instruction count is
                                    // what limits single-source convs, every instruction per element counts
    }
  }
  if (c.has_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[j][i] *= c.mk[j];
    }
  }
  if (c.res1_slot >= 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((c.valid >> j) & 1u) add_h16x8(v[j], c.r1[j], 1.f, p.fmt);
  }
  if (c.res2_slot >= 0) {
    const float r2s = c.res2_scale;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((c.valid >> j) & 1u) add_h16x8(v[j], c.r2[j], r2s, p.fmt);
  }
  const uint32_t valid = c.valid;
  float* const pool = c.pool;
  __nv_bfloat16* d = p.arena + ((static_cast<long>(c.dst_slot) * p.batch + n) * npix + c.pix0) * kC + ch0;
  next();   // `c` belongs to the next tile from here on
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if ((valid >> j) & 1u) {
      uint4 u;
      u.x = pack_h2(v[j][0], v[j][1], p.fmt);
      u.y = pack_h2(v[j][2], v[j][3], p.fmt);
      u.z = pack_h2(v[j][4], v[j][5], p.fmt);
      u.w = pack_h2(v[j][6], v[j][7], p.fmt);
      *reinterpret_cast<uint4*>(d + static_cast<long>(j) * p.width * kC) = u;
    }
  }
  if (pool != nullptr) {
    // Channel sums of the (pre-rounding) outputs over the valid pixels: accumulated per thread over the warp's run of
    // tiles of one (conv, sample) -- the shuffle chain at the end of every tile used to delay the next tile's start --
    // and reduced across the 8 thread groups (7 shuffles, reduce-scatter: channel ch0 + 4 b4 + 2 b3 + b2 per lane) only
    // when the run ends.  The run's total lands in the partial-sum slot of its last tile, the other tiles' slots get zero:
    // consumers still add all tiles * 4 slots per sample.  Deterministic (the tile -> CTA partition is fixed per launch).
    if (valid == 0xfu) {   // interior pixels (all but the tiles on the right / bottom image edge): plain adds
#pragma unroll
      for (int i = 0; i < 8; ++i) c.psum[i] += (v[0][i] + v[1][i]) + (v[2][i] + v[3][i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = (valid & 1u) ? v[0][i] : 0.f, a1 = (valid & 2u) ? v[1][i] : 0.f;
        const float a2 = (valid & 4u) ? v[2][i] : 0.f, a3 = (valid & 8u) ? v[3][i] : 0.f;
        c.psum[i] += (a0 + a1) + (a2 + a3);
      }
    }
    const int npart = p.tiles_x * p.tiles_y * 4;
    const int ch = ch0 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    float total = 0.f;
    if (last_of_run) {   // warp-uniform
      float s[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] = c.psum[i]; c.psum[i] = 0.f; }
#pragma unroll
      for (int off = 16, cnt = 4; off >= 4; off >>= 1, cnt >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
          const float send = upper ? s[i] : s[i + cnt];
          const float keep = upper ? s[i + cnt] : s[i];
          s[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      total = s[0];
    }
    pool[(static_cast<long>(n) * npart + tile * 4 + quad) * kC + ch] = total;
  }
}

// ------------------------------------------------------------------------------------------------ tcgen05 kernel
// Upper 32 bits of a K-major SWIZZLE_128B shared-memory matrix descriptor (version 1, base_offset 0).
__device__ __forceinline__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint64_t make_desc(uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; }

template <int BN, int KS, bool HALO>
__global__ void __launch_bounds__(kNumThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  constexpr int NT = KS * KS;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: the 128-byte swizzle pattern repeats every 1024 bytes of shared address.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr int kBBytes = BN * 128;
  constexpr int kAStage = HALO ? kHaloStageBytes : kTileM * 128;
  constexpr int kAStages = HALO ? 3 : 4;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kARegionBytes;                       // kBBlocks resident blocks (last kRing double as ring)
  uint8_t* smem_ring = smem_b + (kBBlocks - kRing) * kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kBBlocks * kBBytes);
  uint64_t* a_full = bars;                     // [kMaxAStages]
  uint64_t* a_empty = a_full + kMaxAStages;    // [kMaxAStages]
  uint64_t* res_full = a_empty + kMaxAStages;  // [kBBlocks]
  uint64_t* ring_full = res_full + kBBlocks;   // [kRing]
  uint64_t* ring_empty = ring_full + kRing;    // [kRing]
  uint64_t* res_free = ring_empty + kRing;     // [1]
  uint64_t* t_full = res_free + 1;             // [2]
  uint64_t* t_empty = t_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  const int total = p.ngroups * p.batch * tiles;
  const int item_begin = blockIdx.x * p.chunk;
  const int item_end = min(item_begin + p.chunk, total);
  const int nsrc = p.nsrc;
  const int n_res = p.n_res;
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  constexpr int kEpiArrivals = BN == 64 ? 8 : 4;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(HALO ? &p.tm_halo : &p.tm_tile);
    for (int i = 0; i < kMaxAStages; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < kBBlocks; ++i) mbar_init(res_full + i, 1);
    for (int i = 0; i < kRing; ++i) { mbar_init(ring_full + i, 1); mbar_init(ring_empty + i, 1); }
    mbar_init(res_free, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, kEpiArrivals); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (one thread) ================================
    if (lane == 0) {
      int sa = 0, pa = 0, sr = 0, pr = 0;
      int gen = -1, cur_key = -1;
      long long dbg_ae = 0;
      const int nkb = nsrc * NT;
      int tile = item_begin % tiles;
      int gn = item_begin / tiles;
      for (int item = item_begin; item < item_end; ++item) {
        const int n = gn % p.batch;
        const int gi = gn / p.batch;
        const savsr_conv_group& g = p.g[gi];
        const int ty = tile / p.tiles_x;
        const int x0 = (tile - ty * p.tiles_x) * kTileW, y0 = ty * kTileH;
        const uint8_t* wptr = static_cast<const uint8_t*>(g.weight) + static_cast<long>(n) * g.weight_sample_stride;
        const int key = gi * p.batch + (g.weight_sample_stride != 0 ? n : 0);
        if (key != cur_key) {
          // new weight set: wait until every MMA that reads the old resident blocks has retired
          if (gen >= 0) mbar_wait(res_free, gen & 1);
          ++gen;
          cur_key = key;
          // rotated issue order: concurrent CTAs pull different L2 lines of the shared weights
          int kb = blockIdx.x % nkb;
          for (int i = 0; i < nkb; ++i) {
            if (kb < n_res) {
              mbar_expect_tx(res_full + kb, kBBytes);
              bulk_load(smem_b + kb * kBBytes, wptr + static_cast<long>(kb) * kBBytes, kBBytes, res_full + kb);
            }
            if (++kb == nkb) kb = 0;
          }
        }
        for (int s = 0; s < nsrc; ++s) {
          const int img = g.src_slot[s] * p.batch + n;
          if constexpr (HALO) {
            const long long c0 = DBG_CLOCK();
            mbar_wait(a_empty + sa, pa ^ 1);
            dbg_ae += DBG_CLOCK() - c0;
            mbar_expect_tx(a_full + sa, kHaloPitch * (kTileH + 2) * 128u);
            tma_load_4d(smem_a + sa * kAStage, &p.tm_halo, a_full + sa, 0, x0 - 1, y0 - 1, img);
            if (++sa == kAStages) { sa = 0; pa ^= 1; }
          }
#pragma unroll
          for (int tap = 0; tap < NT; ++tap) {
            const int kb = s * NT + tap;
            if constexpr (!HALO) {
              const int dx = KS == 3 ? tap % 3 - 1 : 0;
              const int dy = KS == 3 ? tap / 3 - 1 : 0;
              mbar_wait(a_empty + sa, pa ^ 1);
              mbar_expect_tx(a_full + sa, kTileM * 128u);
              tma_load_4d(smem_a + sa * kAStage, &p.tm_tile, a_full + sa, 0, x0 + dx, y0 + dy, img);
              if (++sa == kAStages) { sa = 0; pa ^= 1; }
            }
            if (kb >= n_res) {
              mbar_wait(ring_empty + sr, pr ^ 1);
              mbar_expect_tx(ring_full + sr, kBBytes);
              bulk_load(smem_ring + sr * kBBytes, wptr + static_cast<long>(kb) * kBBytes, kBBytes, ring_full + sr);
              if (++sr == kRing) { sr = 0; pr ^= 1; }
            }
          }
        }
        if (++tile == tiles) { tile = 0; ++gn; }
      }
      DBG_ONLY(if (p.dbg != nullptr) p.dbg[blockIdx.x * 8 + 6] = dbg_ae;)
      (void)dbg_ae;
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // One elected thread feeds the tensor core, so this loop is the critical path of the kernel.  The whole warp
    // runs it with warp-uniform control flow and values (so descriptors live in uniform registers and ptxas emits
    // straight UTCHMMA sequences); only the tcgen05 instructions themselves sit under elect.sync.
    const uint32_t idesc = umma_idesc_f16(BN, p.fmt);
    constexpr uint32_t a_hi = desc_hi(HALO ? kHaloPitch * 128u : 1024u);
    constexpr uint32_t b_hi = desc_hi(1024u);
    const uint32_t a_lo0 = (smem_u32(smem_a) >> 4) & 0x3fffu;
    const uint32_t b_lo0 = (smem_u32(smem_b) >> 4) & 0x3fffu;
    const uint32_t ring_lo0 = (smem_u32(smem_ring) >> 4) & 0x3fffu;
    const bool all_resident = nsrc * NT <= n_res;
    int sa = 0, pa = 0, sr = 0, pr = 0;
    int gen = -1, cur_key = -1;
    int it = 0;
    // (tile, n, gi) advance incrementally: no divisions or parameter loads on the per-tile critical path
    int tile = item_begin % tiles;
    int n = (item_begin / tiles) % p.batch;
    int gi = (item_begin / tiles) / p.batch;
    const uint32_t ps_mask = p.per_sample_mask;
    const int batch = p.batch;
    long long dbg_te = 0, dbg_af = 0, dbg_t0 = DBG_CLOCK();
    for (int item = item_begin; item < item_end; ++item, ++it) {
      const int key = gi * batch + (((ps_mask >> gi) & 1u) ? n : 0);
      bool fresh = false;
      if (key != cur_key) {
        if (gen >= 0 && elect_one()) umma_commit(res_free);  // fires when all MMAs of the previous weight set are done
        __syncwarp();
        ++gen;
        cur_key = key;
        fresh = true;
      }
      const int acc = it & 1;
      long long c0 = DBG_CLOCK();
      mbar_wait(t_empty + acc, ((it >> 1) & 1) ^ 1);
      dbg_te += DBG_CLOCK() - c0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      if (HALO && all_resident && !fresh) {
        // ---- fast path: weights resident and already landed; one wait + one elected region per source
        uint32_t b_lo = b_lo0;
        for (int s = 0; s < nsrc; ++s) {
          c0 = DBG_CLOCK();
          mbar_wait(a_full + sa, pa);
          dbg_af += DBG_CLOCK() - c0;
          tc_fence_after();
          const uint32_t al0 = a_lo0 + sa * (kAStage >> 4);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < NT; ++tap) {
              const uint32_t al = al0 + ((tap / 3) * kHaloPitch + tap % 3) * 8;
              const uint32_t bl = b_lo + tap * (kBBytes >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, make_desc(a_hi, al + 2 * k), make_desc(b_hi, bl + 2 * k), idesc, (s | tap | k) ? 1u : 0u);
            }
            umma_commit(a_empty + sa);
          }
          __syncwarp();
          b_lo += NT * (kBBytes >> 4);
          if (++sa == kAStages) { sa = 0; pa ^= 1; }
        }
      } else {
        // ---- general path: per-K-block waits (first tile of a weight set, streamed K-blocks, TAP mode)
        uint32_t accumulate = 0;
        uint32_t b_lo = b_lo0;
        int kb = 0;
        for (int s = 0; s < nsrc; ++s) {
          if constexpr (HALO) {
            mbar_wait(a_full + sa, pa);
            tc_fence_after();
          }
#pragma unroll
          for (int tap = 0; tap < NT; ++tap, ++kb, b_lo += kBBytes >> 4) {
            if constexpr (!HALO) {
              mbar_wait(a_full + sa, pa);
              tc_fence_after();
            }
            uint32_t bl = b_lo;
            const bool streamed = kb >= n_res;
            if (streamed) {
              mbar_wait(ring_full + sr, pr);
              tc_fence_after();
              bl = ring_lo0 + sr * (kBBytes >> 4);
            } else if (fresh) {
              mbar_wait(res_full + kb, gen & 1);
              tc_fence_after();
            }
            // halo: tap (dy, dx) starts (dy * pitch + dx) pixel rows (128 B = 8 descriptor units) into the tile
            const uint32_t al = a_lo0 + sa * (kAStage >> 4) + (HALO ? ((tap / 3) * kHaloPitch + tap % 3) * 8 : 0);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // 4 x (K = 16 bf16 = 32 bytes = 2 units) inside the 128-byte swizzle atom
                umma_bf16(d_tmem, make_desc(a_hi, al + 2 * k), make_desc(b_hi, bl + 2 * k), idesc, accumulate);
                accumulate = 1;
              }
              if (streamed) umma_commit(ring_empty + sr);
              if (!HALO) umma_commit(a_empty + sa);
            }
            __syncwarp();
            accumulate = 1;
            if (streamed) {
              if (++sr == kRing) { sr = 0; pr ^= 1; }
            }
            if constexpr (!HALO) {
              if (++sa == kAStages) { sa = 0; pa ^= 1; }
            }
          }
          if constexpr (HALO) {
            if (elect_one()) umma_commit(a_empty + sa);
            __syncwarp();
            if (++sa == kAStages) { sa = 0; pa ^= 1; }
          }
        }
      }
      if (elect_one()) umma_commit(t_full + acc);
      __syncwarp();
      if (++tile == tiles) {
        tile = 0;
        if (++n == batch) { n = 0; ++gi; }
      }
    }
    DBG_ONLY(if (p.dbg != nullptr && lane == 0) {
      p.dbg[blockIdx.x * 8 + 0] = DBG_CLOCK() - dbg_t0;
      p.dbg[blockIdx.x * 8 + 1] = dbg_te;
      p.dbg[blockIdx.x * 8 + 2] = dbg_af;
      p.dbg[blockIdx.x * 8 + 3] = item_end - item_begin;
    })
    (void)dbg_te; (void)dbg_af; (void)dbg_t0;
  } else {
    // ================================ epilogue warps ================================
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access (warp id % 4)
    const int half = (warp - 2) >> 2;   // which 32-column half of the accumulator (N = 64 only)
    if (BN == 64 || half == 0) {
      constexpr int NC = 16;             // columns per warp of the N = 16 destination (AUX16)
      EpiCtx<NC> ec;
      EpiQuad eq;
      ec.bias_group = -1;
      eq.bias_group = -1;
#pragma unroll
      for (int i = 0; i < 8; ++i) eq.psum[i] = 0.f;
      int it = 0;
      int tile = item_begin % tiles;
      int gn = item_begin / tiles;
      long long dbg_tf = 0, dbg_t0 = DBG_CLOCK();
      if constexpr (BN == 64) {
        if (item_begin < item_end) epiq_prefetch(p, p.g[gn / p.batch], gn / p.batch, gn % p.batch, tile, quad, lane, half, eq);
      }
      for (int item = item_begin; item < item_end; ++item, ++it) {
        const int n = gn % p.batch;
        const int gi = gn / p.batch;
        const savsr_conv_group& g = p.g[gi];
        const int acc = it & 1;
        // loads that do not depend on the accumulator fly while the tile's MMAs run
        if constexpr (BN != 64) epi_prefetch<NC>(p, g, gi, n, tile, quad, lane, 0, ec);
        const long long c0 = DBG_CLOCK();
        mbar_wait(t_full + acc, (it >> 1) & 1);
        dbg_tf += DBG_CLOCK() - c0;
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN + half * 32);
        if constexpr (BN == 64) {
          epiq_finish(p, n, tile, quad, lane, half, taddr, eq, item + 1 >= item_end || tile + 1 == tiles,
                      [&] { if (lane == 0) mbar_arrive(t_empty + acc); }, [&] {
            if (item + 1 < item_end) {
              const int nt = tile + 1 == tiles ? 0 : tile + 1, ngn = tile + 1 == tiles ? gn + 1 : gn;
              epiq_prefetch(p, p.g[ngn / p.batch], ngn / p.batch, ngn % p.batch, nt, quad, lane, half, eq);
            }
          });
        } else {
          float v[NC];
          uint32_t r[16];
          tmem_ld16(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(r[c]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + acc);
          epi_finish<NC>(p, g, n, tile, quad, lane, v, 0, ec);
        }
        if (++tile == tiles) { tile = 0; ++gn; }
      }
      DBG_ONLY(if (p.dbg != nullptr && warp == 2 && lane == 0) {
        p.dbg[blockIdx.x * 8 + 4] = DBG_CLOCK() - dbg_t0;
        p.dbg[blockIdx.x * 8 + 5] = dbg_tf;
      })
      (void)dbg_tf; (void)dbg_t0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ big-K variant
// Convs whose weights do not fit the resident region (K > 18 blocks: OSA 192->64 / 320->64, merge 192->64, 320->128).
// Loop order is source-major over a BATCH of up to 4 tiles: the 9 K-blocks of one source are loaded once per batch
// (two 72 KB sets, double buffered) and used for every tile of the batch, each tile accumulating in its own TMEM
// buffer (2 x 4 accumulators of 64 columns = all 512 TMEM columns, so the epilogue of batch b overlaps batch b+1).
// Weight traffic per tile drops 4x versus streaming per tile and no per-tap waits remain on the issue path.
constexpr int kBatchTiles = 4;

constexpr int kBigkThreads = kNumThreads + 32;   // + a second MMA-issuing warp (warp 10)

__global__ void __launch_bounds__(kBigkThreads, 1) conv_igemm_bigk_kernel(const __grid_constant__ ConvParams p) {
  constexpr int BN = 64, NT = 9;
  constexpr int kBBytes = BN * 128;
  constexpr int kAStage = kHaloStageBytes, kAStages = 3;
  constexpr int kSetBlocks = 9, kSetBytes = kSetBlocks * kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kARegionBytes;   // two sets of 9 K-blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + 2 * kSetBytes);
  uint64_t* a_full = bars;                       // [4]
  uint64_t* a_empty = a_full + kMaxAStages;      // [4]
  uint64_t* set_full = a_empty + kMaxAStages;    // [2]
  uint64_t* set_empty = set_full + 2;            // [2]
  uint64_t* t_full = set_empty + 2;              // [8]
  uint64_t* t_empty = t_full + 2 * kBatchTiles;  // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2 * kBatchTiles);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  const int total = p.ngroups * p.batch * tiles;
  const int item_begin = blockIdx.x * p.chunk;
  const int item_end = min(item_begin + p.chunk, total);
  const int nsrc = p.nsrc;
  // With at most two sources both weight sets stay resident (set s = source s) for as long as consecutive batches
  // use the same weights; otherwise the two sets double-buffer the sources of one batch.
  const bool pinned = nsrc <= 2;
  const uint32_t ps_mask = p.per_sample_mask;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tm_halo);
    for (int i = 0; i < kMaxAStages; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, p.issuers); }   // owner's commit (+ the other issuer's pass)
    for (int i = 0; i < 2; ++i) { mbar_init(set_full + i, 1); mbar_init(set_empty + i, p.issuers); }   // every issuer releases a set
    for (int i = 0; i < 2 * kBatchTiles; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();        // everything above touched only this CTA's own state; from here on the previous kernel's results are read
  pdl_trigger();

  // every role walks the same sequence of batches: up to 4 consecutive tiles of one (conv, sample)
  int item = item_begin;
  int tile = item_begin % tiles;
  int gn = item_begin / tiles;
  int bcount = 0;
  uint32_t use_bits = 0;      // per accumulator: parity of how many times it has been used
  uint32_t set_loads[2] = {0, 0};  // per weight set: how many times it has been (re)loaded
  int cur_key = -1;

  if (warp == 0) {
    // ================================ TMA producer (one thread) ================================
    if (lane == 0) {
      int sa = 0, pa = 0;
      uint32_t u = 0;
      while (item < item_end) {
        const int cnt = min(kBatchTiles, min(item_end - item, tiles - tile));
        const int n = gn % p.batch;
        const int gi = gn / p.batch;
        const savsr_conv_group& g = p.g[gi];
        const uint8_t* wptr = static_cast<const uint8_t*>(g.weight) + static_cast<long>(n) * g.weight_sample_stride;
        const int key = gi * p.batch + (((ps_mask >> gi) & 1u) ? n : 0);
        const bool load = !pinned || key != cur_key;
        cur_key = key;
        for (int s = 0; s < nsrc; ++s) {
          if (load) {
            const int set = pinned ? s : static_cast<int>(u++ & 1);
            mbar_wait(set_empty + set, (set_loads[set] & 1u) ^ 1u);
            ++set_loads[set];
            mbar_expect_tx(set_full + set, kSetBytes);
            int kb = blockIdx.x % kSetBlocks;  // rotated issue order across CTAs
            for (int i = 0; i < kSetBlocks; ++i) {
              bulk_load(smem_b + set * kSetBytes + kb * kBBytes, wptr + static_cast<long>(s * NT + kb) * kBBytes, kBBytes, set_full + set);
              if (++kb == kSetBlocks) kb = 0;
            }
          }
          const int img = g.src_slot[s] * p.batch + n;
          for (int j = 0; j < cnt; ++j) {
            const int tj = tile + j;
            const int ty = tj / p.tiles_x;
            mbar_wait(a_empty + sa, pa ^ 1);
            mbar_expect_tx(a_full + sa, kHaloPitch * (kTileH + 2) * 128u);
            tma_load_4d(smem_a + sa * kAStage, &p.tm_halo, a_full + sa, 0, (tj - ty * p.tiles_x) * kTileW - 1, ty * kTileH - 1, img);
            if (++sa == kAStages) { sa = 0; pa ^= 1; }
          }
        }
        item += cnt; tile += cnt;
        if (tile == tiles) { tile = 0; ++gn; }
      }
    }
  } else if (warp == 1 || (warp == 10 && p.issuers == 2)) {
    // ================================ two MMA issuers (warp-uniform, elected issue) ================================
    // In situ one issuing thread also waits on barriers, commits and walks the batch structure between tiles, and the
    // tensor core idles meanwhile (64 cycles per MMA measured with one issuer, 52 with two): warp 1 takes the even
    // tiles of a batch, warp 10 the odd ones, each into its own TMEM accumulators, so one fills the other's gaps.
    const int my = warp == 1 ? 0 : 1;
    const int own_mask = p.issuers == 2 ? 1 : 0;   // tile j of a batch belongs to issuer (j & own_mask)
    const uint32_t idesc = umma_idesc_f16(BN, p.fmt);
    constexpr uint32_t a_hi = desc_hi(kHaloPitch * 128u);
    constexpr uint32_t b_hi = desc_hi(1024u);
    const uint32_t a_lo0 = (smem_u32(smem_a) >> 4) & 0x3fffu;
    const uint32_t b_lo0 = (smem_u32(smem_b) >> 4) & 0x3fffu;
    int sa = 0, pa = 0;
    uint32_t u = 0;
    long long dbg_te = 0, dbg_af = 0, dbg_sf = 0, dbg_t0 = DBG_CLOCK(), c0;
    while (item < item_end) {
      const int cnt = min(kBatchTiles, min(item_end - item, tiles - tile));
      const int bb = bcount & 1;
      const int n = gn % p.batch;
      const int gi = gn / p.batch;
      const int key = gi * p.batch + (((ps_mask >> gi) & 1u) ? n : 0);
      const bool load = !pinned || key != cur_key;
      if (pinned && load && cur_key >= 0) {
        // the pinned sets are about to be replaced: release them once every MMA issued so far has retired
        if (elect_one()) {
          for (int s = 0; s < nsrc; ++s) umma_commit(set_empty + s);
        }
        __syncwarp();
      }
      cur_key = key;
      c0 = DBG_CLOCK();
#pragma unroll
      for (int j = 0; j < kBatchTiles; ++j) {
        if (j < cnt && (j & own_mask) == my) mbar_wait(t_empty + bb * kBatchTiles + j, ((use_bits >> (bb * kBatchTiles + j)) & 1u) ^ 1u);
      }
      dbg_te += DBG_CLOCK() - c0;
      tc_fence_after();
      for (int s = 0; s < nsrc; ++s) {
        const int set = pinned ? s : static_cast<int>(u & 1);
        if (load) {
          c0 = DBG_CLOCK();
          mbar_wait(set_full + set, set_loads[set] & 1u);
          dbg_sf += DBG_CLOCK() - c0;
          ++set_loads[set];
          tc_fence_after();
          ++u;
        }
        const uint32_t bl0 = b_lo0 + set * (kSetBytes >> 4);
#pragma unroll
        for (int j = 0; j < kBatchTiles; ++j) {
          if (j < cnt) {
            // BOTH issuers observe every stage fill in order and the stage is refilled only after both have passed it
            // (a_empty counts 2): mbarrier parity cannot distinguish phases two apart, so neither issuer may run
            // more than one fill ahead of, or behind, the barrier it waits on.
            c0 = DBG_CLOCK();
            mbar_wait(a_full + sa, pa);
            dbg_af += DBG_CLOCK() - c0;
            if ((j & own_mask) != my) {
              if (elect_one()) mbar_arrive(a_empty + sa);
              __syncwarp();
            } else {
              tc_fence_after();
              const uint32_t al0 = a_lo0 + sa * (kAStage >> 4);
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((bb * kBatchTiles + j) * BN);
              if (elect_one()) {
                // The tap-row loop is deliberately NOT unrolled: fully unrolled, ptxas precomputes all 72 descriptors,
                // spills uniform registers and needs 58 cycles per MMA from one thread; this compact form
                // (UTCHMMA / UIADD3.64 pairs) issues at 50, the tensor core retiring one N = 64 MMA per 48
                // (scripts/umma_bench.cu, "rolled tap loop").
                uint32_t al = al0, bl = bl0, acc = s ? 1u : 0u;
                if (p.ksteps == 4) {
#pragma unroll 1
                  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        umma_bf16(d_tmem, make_desc(a_hi, al + dx * 8 + 2 * k), make_desc(b_hi, bl + dx * (kBBytes >> 4) + 2 * k), idesc, acc);
                        acc = 1u;
                      }
                    }
                    al += kHaloPitch * 8;
                    bl += 3 * (kBBytes >> 4);
                  }
                } else {
                  // sources whose trailing channels are zero by construction (the packed 7 x 3 input frames of the first layer):
                  // the K steps over those channels would multiply zeros by zero-expanded filter columns -- skip them
#pragma unroll 1
                  for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
#pragma unroll 1
                      for (int k = 0; k < p.ksteps; ++k) {
                        umma_bf16(d_tmem, make_desc(a_hi, al + dx * 8 + 2 * k), make_desc(b_hi, bl + dx * (kBBytes >> 4) + 2 * k), idesc, acc);
                        acc = 1u;
                      }
                    }
                    al += kHaloPitch * 8;
                    bl += 3 * (kBBytes >> 4);
                  }
                }
                umma_commit(a_empty + sa);
                // the tile's accumulator is complete once its LAST source has been applied: hand it to the epilogue right away
                // instead of at the end of the batch (the epilogue of tiles 0, 1 then overlaps the MMAs of tiles 2, 3: shorter
                // pipeline fill and drain per launch, accumulators recycle earlier)
                if (s == nsrc - 1) umma_commit(t_full + bb * kBatchTiles + j);
              }
              __syncwarp();
            }
            if (++sa == kAStages) { sa = 0; pa ^= 1; }
          }
        }
        if (!pinned) {
          if (elect_one()) umma_commit(set_empty + set);
          __syncwarp();
        }
      }
#pragma unroll
      for (int j = 0; j < kBatchTiles; ++j)
        if (j < cnt) use_bits ^= 1u << (bb * kBatchTiles + j);
      ++bcount;
      item += cnt; tile += cnt;
      if (tile == tiles) { tile = 0; ++gn; }
    }
    DBG_ONLY(if (p.dbg != nullptr && lane == 0 && my == 0) {
      p.dbg[blockIdx.x * 8 + 0] = DBG_CLOCK() - dbg_t0;
      p.dbg[blockIdx.x * 8 + 1] = dbg_te;
      p.dbg[blockIdx.x * 8 + 2] = dbg_af;
      p.dbg[blockIdx.x * 8 + 3] = item_end - item_begin;
      p.dbg[blockIdx.x * 8 + 7] = dbg_sf;
    })
    (void)dbg_te; (void)dbg_af; (void)dbg_sf; (void)dbg_t0;
  } else if (warp < 10) {   // (warp 10 is the optional second issuer; idle when p.issuers == 1)
    // ================================ epilogue warps ================================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    EpiQuad eq;
    eq.bias_group = -1;
#pragma unroll
    for (int i = 0; i < 8; ++i) eq.psum[i] = 0.f;
    long long dbg_tf = 0, dbg_t0 = DBG_CLOCK();
    if (item < item_end) epiq_prefetch(p, p.g[gn / p.batch], gn / p.batch, gn % p.batch, tile, quad, lane, half, eq);
    while (item < item_end) {
      const int cnt = min(kBatchTiles, min(item_end - item, tiles - tile));
      const int bb = bcount & 1;
      const int n = gn % p.batch;
      for (int j = 0; j < cnt; ++j) {
        const int acc = bb * kBatchTiles + j;
        const long long c0 = DBG_CLOCK();
        mbar_wait(t_full + acc, (use_bits >> acc) & 1u);
        dbg_tf += DBG_CLOCK() - c0;
        use_bits ^= 1u << acc;
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN + half * 32);
        epiq_finish(p, n, tile + j, quad, lane, half, taddr, eq, item + j + 1 >= item_end || tile + j + 1 == tiles,
                    [&] { if (lane == 0) mbar_arrive(t_empty + acc); }, [&] {
          if (item + j + 1 < item_end) {   // a batch never straddles a (conv, sample) boundary
            const int nt = tile + j + 1 == tiles ? 0 : tile + j + 1, ngn = tile + j + 1 == tiles ? gn + 1 : gn;
            epiq_prefetch(p, p.g[ngn / p.batch], ngn / p.batch, ngn % p.batch, nt, quad, lane, half, eq);
          }
        });
      }
      ++bcount;
      item += cnt; tile += cnt;
      if (tile == tiles) { tile = 0; ++gn; }
    }
    DBG_ONLY(if (p.dbg != nullptr && warp == 2 && lane == 0) {
      p.dbg[blockIdx.x * 8 + 4] = DBG_CLOCK() - dbg_t0;
      p.dbg[blockIdx.x * 8 + 5] = dbg_tf;
    })
    (void)dbg_tf; (void)dbg_t0;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ checker kernel
// One block of 128 threads per (conv, sample, tile); thread m accumulates all N outputs of its pixel with
// plain loads, reading the same packed (swizzled) weights.  Same epilogue.  Slow by design.
template <int BN>
__global__ void __launch_bounds__(128) conv_check_kernel(const __grid_constant__ ConvParams p) {
  const int tiles = p.tiles_x * p.tiles_y;
  const int item = blockIdx.x;
  const int tile = item % tiles;
  const int gn = item / tiles;
  const int n = gn % p.batch;
  const int gi = gn / p.batch;
  const savsr_conv_group& g = p.g[gi];
  const int m = threadIdx.x;
  const int px = (tile % p.tiles_x) * kTileW + (m & 7);
  const int py = (tile / p.tiles_x) * kTileH + (m >> 3);
  const long npix = static_cast<long>(p.height) * p.width;
  const uint8_t* wptr = static_cast<const uint8_t*>(g.weight) + static_cast<long>(n) * g.weight_sample_stride;
  constexpr int NC = BN == 64 ? 32 : 16;
  for (int col0 = 0; col0 < BN; col0 += NC) {
    float v[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) v[c] = 0.f;
    for (int s = 0; s < g.nsrc; ++s) {
      const __nv_bfloat16* src = p.arena + (static_cast<long>(g.src_slot[s]) * p.batch + n) * npix * kC;
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int dx = p.ntaps == 9 ? tap % 3 - 1 : 0, dy = p.ntaps == 9 ? tap / 3 - 1 : 0;
        const int sx = px + dx, sy = py + dy;
        const bool in = sx >= 0 && sx < p.width && sy >= 0 && sy < p.height;
        const uint8_t* wb = wptr + static_cast<long>(s * p.ntaps + tap) * (BN * 128);
        for (int k = 0; k < 64; ++k) {
          const float a = in ? h_to_float(reinterpret_cast<const uint16_t*>(src)[(static_cast<long>(sy) * p.width + sx) * kC + k], p.fmt) : 0.f;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int row = BN == 64 ? quad_row(col0 + c) : col0 + c;   // packed rows of N = 64 blocks are in SAVSR_ROWS_QUAD order
            const int off = row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
            v[c] += a * h_to_float(*reinterpret_cast<const uint16_t*>(wb + off), p.fmt);
          }
        }
      }
    }
    EpiCtx<NC> ec;
    ec.bias_group = -1;
    epi_prefetch<NC>(p, g, gi, n, tile, m >> 5, m & 31, col0, ec);
    epi_finish<NC>(p, g, n, tile, m >> 5, m & 31, v, col0, ec);
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// fp32 OIHW -> packed bf16 blocks [co/n_tile][ci/64 * k*k][n_tile][64], 128-byte swizzled rows.
__global__ void pack_weight_kernel(const float* __restrict__ w, int co_real, int co, int ci, int ks, int n_tile, int fmt, int quad,
                                   uint16_t* __restrict__ out) {
  const long total = static_cast<long>(co) * ci * ks * ks;
  const int taps = ks * ks;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    // idx enumerates the OUTPUT position: (ng, kb, n, k)
    const int k = idx & 63;
    long r = idx >> 6;
    const int n = r % n_tile; r /= n_tile;
    const int kb = r % ((ci / 64) * taps);
    const int ng = r / ((ci / 64) * taps);
    const int s = kb / taps, tap = kb % taps;
    const int o = ng * n_tile + (quad ? quad_row(n) : n), i = s * 64 + k;   // row n of the block holds channel o
    const float val = o < co_real ? w[(static_cast<long>(o) * ci + i) * taps + tap] : 0.f;
    const long block = static_cast<long>(ng) * ((ci / 64) * taps) + kb;
    const long off = block * (n_tile * 64) + n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7));
    out[off] = float_to_h(val, fmt);
  }
}

// ------------------------------------------------------------------------------------------------ arena import / export
__global__ void arena_import_kernel(const float* __restrict__ nchw, uint16_t* __restrict__ dst, int batch, long npix, int fmt) {
  const long total = static_cast<long>(batch) * npix * kC;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = idx & 63;
    const long pn = idx >> 6;
    const long pix = pn % npix, n = pn / npix;
    dst[idx] = float_to_h(nchw[(n * kC + c) * npix + pix], fmt);
  }
}
__global__ void arena_export_kernel(const uint16_t* __restrict__ src, float* __restrict__ nchw, int batch, long npix, int fmt) {
  const long total = static_cast<long>(batch) * npix * kC;
  for (long idx = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const long pix = idx % npix;
    const long nc = idx / npix;
    const int c = nc % kC;
    const long n = nc / kC;
    nchw[idx] = h_to_float(src[(n * npix + pix) * kC + c], fmt);
  }
}

template <int BN, int KS, bool HALO>
static int launch_igemm(savsr_ctx* ctx, ConvParams& p, int total, cudaStream_t st) {
  const size_t smem = 1024 + kARegionBytes + kBBlocks * BN * 128 + 512;
  constexpr int variant = (BN == 64 ? 0 : 3) + (KS == 1 ? 0 : (HALO ? 1 : 2));
  if (int rc = ensure_smem_attr(ctx, kAttrIgemm + variant, conv_igemm_kernel<BN, KS, HALO>, smem)) return rc;
  // persistent CTAs over contiguous chunks of work items (weights stay resident within a chunk)
  p.chunk = (total + ctx->sm_count - 1) / ctx->sm_count;
  const int grid = (total + p.chunk - 1) / p.chunk;
  conv_igemm_kernel<BN, KS, HALO><<<grid, kNumThreads, smem, st>>>(p);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

template <int BN>
static int launch_conv(savsr_ctx* ctx, ConvParams& p, int impl, cudaStream_t st) {
  const int total = p.ngroups * p.batch * p.tiles_x * p.tiles_y;
  if (total == 0) return 0;
  if (impl == SAVSR_IMPL_CHECK) {
    conv_check_kernel<BN><<<total, 128, 0, st>>>(p);
    SAVSR_CUDA(cudaGetLastError());
    return 0;
  }
  if (p.ntaps == 1) return launch_igemm<BN, 1, false>(ctx, p, total, st);
  if constexpr (BN == 64) {
    // the batched dual-issuer kernel is the default for every 3x3 HALO conv; SAVSR_BIGK_ALL=0 restricts it to K > 18 blocks
    if (p.halo && (ctx->opt[SAVSR_OPT_BIGK_ALL] || p.nsrc * p.ntaps > kBBlocks)) {
      const size_t smem = 1024 + kARegionBytes + kBBlocks * BN * 128 + 512;
      if (int rc = ensure_smem_attr(ctx, kAttrBigk, conv_igemm_bigk_kernel, smem)) return rc;
      p.chunk = (total + ctx->sm_count - 1) / ctx->sm_count;
      const int grid = (total + p.chunk - 1) / p.chunk;
      SAVSR_CUDA(launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, conv_igemm_bigk_kernel, dim3(grid), dim3(kBigkThreads), smem, st, p));
      return 0;
    }
  }
  if (p.halo) return launch_igemm<BN, 3, true>(ctx, p, total, st);
  return launch_igemm<BN, 3, false>(ctx, p, total, st);
}

}  // namespace savsr

using namespace savsr;

#ifdef SAVSR_DEBUG_COUNTERS
// bring-up builds only (not in the public header): device buffer [grid][8] receiving cycle counters of the context's next conv launches
extern "C" void savsr_debug_conv_counters(savsr_ctx* ctx, long long* dev_buf) { if (ctx) ctx->conv_dbg = dev_buf; }
#endif

extern "C" size_t savsr_packed_weight_bytes(int co, int ci, int ksize) {
  return static_cast<size_t>(co) * ci * ksize * ksize * sizeof(__nv_bfloat16);
}

extern "C" int savsr_pack_conv_weight(const float* w_oihw, int co_real, int co, int ci, int ksize, int n_tile, int format,
                                      int row_order, void* packed, savsr_stream st) {
  SAVSR_REQUIRE(format == SAVSR_FMT_BF16 || format == SAVSR_FMT_FP16, "savsr_pack_conv_weight: unknown format %d", format);
  SAVSR_REQUIRE(row_order == SAVSR_ROWS_LINEAR || row_order == SAVSR_ROWS_QUAD, "savsr_pack_conv_weight: unknown row_order %d", row_order);
  SAVSR_REQUIRE(w_oihw && packed, "savsr_pack_conv_weight: null pointer");
  SAVSR_REQUIRE(ksize == 1 || ksize == 3, "savsr_pack_conv_weight: ksize must be 1 or 3, got %d", ksize);
  SAVSR_REQUIRE(n_tile == 64 || n_tile == 16, "savsr_pack_conv_weight: n_tile must be 64 or 16, got %d", n_tile);
  SAVSR_REQUIRE(ci > 0 && ci % 64 == 0, "savsr_pack_conv_weight: ci (%d) must be a positive multiple of 64", ci);
  SAVSR_REQUIRE(co > 0 && co % n_tile == 0 && co_real <= co && co_real > 0,
                "savsr_pack_conv_weight: co (%d) must be a multiple of n_tile (%d) and >= co_real (%d)", co, n_tile, co_real);
  const long total = static_cast<long>(co) * ci * ksize * ksize;
  const int blocks = static_cast<int>((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  pack_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(st)>>>(w_oihw, co_real, co, ci, ksize, n_tile, format,
                                                                       (n_tile == 64 && row_order == SAVSR_ROWS_QUAD) ? 1 : 0,
                                                                       static_cast<uint16_t*>(packed));
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_arena_import(savsr_arena* a, int slot, const float* nchw, savsr_stream st) {
  SAVSR_REQUIRE(a && nchw, "savsr_arena_import: null pointer");
  SAVSR_REQUIRE(slot >= 0 && slot < a->nslots, "savsr_arena_import: slot %d out of range [0,%d)", slot, a->nslots);
  const long npix = static_cast<long>(a->height) * a->width;
  arena_import_kernel<<<1024, 256, 0, static_cast<cudaStream_t>(st)>>>(
      nchw, reinterpret_cast<uint16_t*>(a->base + static_cast<long>(slot) * a->batch * npix * kC), a->batch, npix, a->ctx->fmt);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_arena_export(savsr_arena* a, int slot, float* nchw, savsr_stream st) {
  SAVSR_REQUIRE(a && nchw, "savsr_arena_export: null pointer");
  SAVSR_REQUIRE(slot >= 0 && slot < a->nslots, "savsr_arena_export: slot %d out of range [0,%d)", slot, a->nslots);
  const long npix = static_cast<long>(a->height) * a->width;
  arena_export_kernel<<<1024, 256, 0, static_cast<cudaStream_t>(st)>>>(
      reinterpret_cast<const uint16_t*>(a->base + static_cast<long>(slot) * a->batch * npix * kC), nchw, a->batch, npix, a->ctx->fmt);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_conv(savsr_ctx* ctx, savsr_arena* arena, const savsr_conv_group* groups, int ngroups, int ksize,
                          int n_tile, int dst_mode, int impl, savsr_stream st) {
  SAVSR_REQUIRE(ctx && arena && groups, "savsr_conv: null pointer");
  DeviceGuard guard(ctx->device);
  SAVSR_REQUIRE(ngroups >= 0 && ngroups <= SAVSR_MAX_GROUPS, "savsr_conv: ngroups %d out of range [0,%d]", ngroups, SAVSR_MAX_GROUPS);
  SAVSR_REQUIRE(ksize == 1 || ksize == 3, "savsr_conv: ksize must be 1 or 3, got %d", ksize);
  SAVSR_REQUIRE(impl >= SAVSR_IMPL_TCGEN05_TAP && impl <= SAVSR_IMPL_CHECK, "savsr_conv: unknown impl %d", impl);
  SAVSR_REQUIRE((n_tile == 64 && dst_mode == SAVSR_DST_ARENA) || (n_tile == 16 && dst_mode == SAVSR_DST_AUX16),
                "savsr_conv: n_tile %d does not match dst_mode %d", n_tile, dst_mode);
  if (ngroups == 0) return 0;

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tm_tile = arena->tm_tile;
  p.tm_halo = arena->tm_halo;
  for (int i = 0; i < ngroups; ++i) {
    const savsr_conv_group& g = groups[i];
    SAVSR_REQUIRE(g.nsrc >= 1 && g.nsrc <= SAVSR_MAX_SRC, "savsr_conv: group %d nsrc %d out of range", i, g.nsrc);
    SAVSR_REQUIRE(g.weight != nullptr, "savsr_conv: group %d has no weights", i);
    SAVSR_REQUIRE(g.nsrc == groups[0].nsrc, "savsr_conv: all groups of a launch must have the same nsrc (%d vs %d)", g.nsrc, groups[0].nsrc);
    SAVSR_REQUIRE(g.src_channels == groups[0].src_channels && g.src_channels >= 0 && g.src_channels <= 64 && g.src_channels % 16 == 0,
                  "savsr_conv: src_channels (%d) must be 0, 16, 32, 48 or 64 and equal for all groups of a launch", g.src_channels);
    SAVSR_REQUIRE(g.act != SAVSR_ACT_LRELU || (g.slope >= 0.f && g.slope <= 1.f), "savsr_conv: group %d LeakyReLU slope %g outside [0, 1]", i, g.slope);
    for (int s = 0; s < g.nsrc; ++s) {
      SAVSR_REQUIRE(g.src_slot[s] >= 0 && g.src_slot[s] < arena->nslots, "savsr_conv: group %d source slot %d out of range", i, g.src_slot[s]);
      SAVSR_REQUIRE(dst_mode != SAVSR_DST_ARENA || g.src_slot[s] != g.dst_slot, "savsr_conv: group %d writes slot %d that it also convolves", i, g.dst_slot);
    }
    if (dst_mode == SAVSR_DST_ARENA) {
      SAVSR_REQUIRE(g.dst_slot >= 0 && g.dst_slot < arena->nslots, "savsr_conv: group %d dst slot %d out of range", i, g.dst_slot);
      SAVSR_REQUIRE(g.res1_slot < arena->nslots && g.res2_slot < arena->nslots, "savsr_conv: group %d residual slot out of range", i);
    } else {
      SAVSR_REQUIRE(g.aux_dst != nullptr, "savsr_conv: group %d needs aux_dst for dst_mode %d", i, dst_mode);
    }
    p.g[i] = g;
  }
  p.arena = arena->base;
  p.ngroups = ngroups;
  p.batch = arena->batch;
  p.height = arena->height;
  p.width = arena->width;
  p.tiles_x = arena->tiles_x;
  p.tiles_y = arena->tiles_y;
  p.ntaps = ksize * ksize;
  p.halo = (impl == SAVSR_IMPL_TCGEN05_HALO && ksize == 3) ? 1 : 0;
  p.nsrc = groups[0].nsrc;
  for (int i = 0; i < ngroups; ++i) if (groups[i].weight_sample_stride != 0) p.per_sample_mask |= 1u << i;
  const int nkb = p.nsrc * p.ntaps;
  p.n_res = nkb <= kBBlocks ? nkb : kBBlocks - kRing;
  p.dst_mode = dst_mode;
  p.fmt = ctx->fmt;
  p.dbg = ctx->conv_dbg;
  p.issuers = ctx->opt[SAVSR_OPT_BIGK_ISSUERS];
  p.ksteps = groups[0].src_channels ? groups[0].src_channels / 16 : 4;   // honoured by the batched 3x3 kernel; the others always run all four
  if (n_tile == 64) return launch_conv<64>(ctx, p, impl, static_cast<cudaStream_t>(st));
  return launch_conv<16>(ctx, p, impl, static_cast<cudaStream_t>(st));
}
