// SATU kernels (savsr_arch.py:217-376): bit-exact coordinate / index kernel + per-scale MLP table,
// spatio-temporal filtering, and the fused HR gather with the routed compress/expand experts.
//
// Coordinate arithmetic follows the reference's fp32 operation order exactly (CPU semantics: IEEE
// division by the scale, no FMA contraction), hence the explicit __fadd_rn/__fmul_rn/__fdiv_rn.
#include "common.cuh"

namespace savsr {

__device__ __forceinline__ float rel_q(int i, float s) { return __fdiv_rn(__fadd_rn(static_cast<float>(i), 0.5f), s); }
// R(i) = (q - floor(q + 1e-3)) - 0.5   (savsr_arch.py:331, 333)
__device__ __forceinline__ float rel_coord(int i, float s, int* cell) {
  const float q = rel_q(i, s);
  const float fl = floorf(__fadd_rn(q, 1e-3f));
  if (cell) *cell = static_cast<int>(fl);
  return __fsub_rn(__fsub_rn(q, fl), 0.5f);
}
// normalised base grid coordinate (savsr_arch.py:275-280), zero offset
__device__ __forceinline__ float base_norm(int i, float s, int n_lr) {
  float g = __fsub_rn(__fdiv_rn(__fadd_rn(static_cast<float>(i), 0.5f), s), 0.5f);
  g = __fsub_rn(__fdiv_rn(__fmul_rn(g, 2.f), static_cast<float>(n_lr - 1)), 1.f);
  return g;
}
// ATen grid_sampler un-normalisation, align_corners = True
__device__ __forceinline__ float unnormalize(float g, int n_lr) {
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), static_cast<float>(n_lr - 1));
}

__host__ __device__ constexpr uint32_t desc_hi_1024() { return (1024u >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint64_t make_desc64(uint32_t hi, uint32_t lo) { return (static_cast<uint64_t>(hi) << 32) | lo; }

__global__ void satu_axis_kernel(int n_out, int n_lr, float s, float* rel, int32_t* cell, float* base, int32_t* corner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  int c;
  const float r = rel_coord(i, s, &c);
  if (rel) rel[i] = r;
  if (cell) cell[i] = c;
  const float g = base_norm(i, s, n_lr);
  if (base) base[i] = g;
  if (corner) corner[i] = static_cast<int32_t>(floorf(unnormalize(g, n_lr)));
}

// table[pix] = (offset_x, offset_y, st_offset_x, st_offset_y, r0..r3), the 4 -> 64 -> 64 MLP and its heads
// (savsr_arch.py:335-351).  One HR pixel per thread, weights in shared memory.
__global__ void __launch_bounds__(128) satu_table_kernel(const savsr_satu_weights wts, int H, int W, float s_h, float s_w,
                                                         float* __restrict__ table) {
  __shared__ __align__(16) float w0[4 * 64];    // [in][out]
  __shared__ __align__(16) float w2[64 * 64];   // [in][out]
  __shared__ __align__(16) float wh[8 * 64];    // [head][in]: offset(2), st_offset(2), routing(4)
  __shared__ float b0[64], b2[64], bh[8];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) w0[i] = wts.body0_w[(i & 63) * 4 + (i >> 6)];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) w2[i] = wts.body2_w[(i & 63) * 64 + (i >> 6)];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const int hd = i >> 6, j = i & 63;
    wh[i] = hd < 2 ? wts.offset_w[hd * 64 + j] : hd < 4 ? wts.st_offset_w[(hd - 2) * 64 + j] : wts.routing_w[(hd - 4) * 64 + j];
  }
  if (threadIdx.x < 64) { b0[threadIdx.x] = wts.body0_b[threadIdx.x]; b2[threadIdx.x] = wts.body2_b[threadIdx.x]; }
  if (threadIdx.x < 8) {
    const int hd = threadIdx.x;
    bh[hd] = hd < 2 ? wts.offset_b[hd] : hd < 4 ? wts.st_offset_b[hd - 2] : wts.routing_b[hd - 4];
  }
  __syncthreads();
  const long pix = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (pix >= static_cast<long>(H) * W) return;
  const int i = pix / W, j = pix % W;
  float in[4];
  in[0] = __fdiv_rn(1.f, s_w);  // channel 0 is 1/s_w (savsr_arch.py:336)
  in[1] = __fdiv_rn(1.f, s_h);
  in[2] = rel_coord(i, s_h, nullptr);
  in[3] = rel_coord(j, s_w, nullptr);
  float e1[64];
#pragma unroll
  for (int o = 0; o < 64; ++o) {
    float a = b0[o];
#pragma unroll
    for (int k = 0; k < 4; ++k) a += w0[k * 64 + o] * in[k];
    e1[o] = fmaxf(a, 0.f);
  }
  float e2[64];
#pragma unroll
  for (int o = 0; o < 64; ++o) e2[o] = b2[o];
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    const float a = e1[k];
    const float4* wr = reinterpret_cast<const float4*>(w2 + k * 64);
#pragma unroll
    for (int o4 = 0; o4 < 16; ++o4) {
      const float4 wv = wr[o4];
      e2[4 * o4 + 0] += a * wv.x; e2[4 * o4 + 1] += a * wv.y;
      e2[4 * o4 + 2] += a * wv.z; e2[4 * o4 + 3] += a * wv.w;
    }
  }
  float out[8];
#pragma unroll
  for (int hd = 0; hd < 8; ++hd) {
    float a = bh[hd];
#pragma unroll
    for (int k = 0; k < 64; ++k) a += wh[hd * 64 + k] * fmaxf(e2[k], 0.f);
    out[hd] = hd < 4 ? a : 1.f / (1.f + expf(-a));
  }
  float4* d = reinterpret_cast<float4*>(table + pix * 8);
  d[0] = make_float4(out[0], out[1], out[2], out[3]);
  d[1] = make_float4(out[4], out[5], out[6], out[7]);
}

// ------------------------------------------------------------------------------------------------ kernel_conv + sta_conv, fused
// sta[p][c] = sum_{t in 5x5} x[clamp(p + d_t)][c] * K_t[p][c],   K_t = LeakyReLU_0.1(W_t a[p] + b_t)   (savsr_arch.py:297-313)
// The unfused route materialises the 25 per-pixel kernels K_t (1600 channels, 83 MB per sample) and reads them back.  Here
// each tap is one M=128 (8x16-pixel tile) x N=64 x K=64 GEMM on tcgen05 whose accumulator is consumed straight from TMEM:
// the epilogue threads multiply it with the tap's neighbour feature (16-byte loads that hit L1) and keep the 25-tap sum
// in registers, so K never leaves the SM.  Work unit = batch of 2 tiles sharing one pass over the 25 weight blocks (8 KB each,
// streamed through a ring); TMEM holds 4 taps x 2 tiles of accumulators, so the tensor core runs up to 4 taps ahead.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..17 = epilogue (TMEM lane quadrant = warp % 4; 2 column halves x 2 tiles).
struct KstaParams {
  CUtensorMap tm_tile;            // box [64 ch, 8, 16, 1] over the LR arena
  const __nv_bfloat16* arena;
  const uint8_t* weights;         // packed [25][64][64], rows in SAVSR_ROWS_QUAD order
  const float* bias;              // [25][64]
  float slope;
  int batch, hp, wp, h, w;
  int a_slot, x_slot, dst_slot;
  int tiles_x, tiles_y;
  int nbatches, chunk;            // work items = 2-tile batches; `chunk` consecutive ones per CTA
  int fmt;
};
constexpr int kKstaThreads = 18 * 32;
constexpr int kKstaWStages = 6;
constexpr int kKstaSmem = 1024 + 4 * 16384 + kKstaWStages * 8192 + 512;

__global__ void __launch_bounds__(kKstaThreads, 1) satu_kconv_sta_kernel(const __grid_constant__ KstaParams p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;                       // [2 buffers][2 tiles][16 KB]
  uint8_t* smem_w = smem + 4 * 16384;           // ring of weight blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + kKstaWStages * 8192);
  uint64_t* a_full = bars;                      // [2]
  uint64_t* a_empty = a_full + 2;               // [2]
  uint64_t* w_full = a_empty + 2;               // [6]
  uint64_t* w_empty = w_full + kKstaWStages;    // [6]
  uint64_t* t_full = w_empty + kKstaWStages;    // [8]  accumulator (tap & 3) * 2 + tile
  uint64_t* t_empty = t_full + 8;               // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.tiles_x * p.tiles_y;
  const int per_img = (tiles + 1) >> 1;          // 2-tile batches per sample (the last one may hold a single tile)
  const int item_begin = blockIdx.x * p.chunk;
  const int item_end = min(item_begin + p.chunk, p.nbatches);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tm_tile);
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + i, 1); mbar_init(a_empty + i, 1); }
    for (int i = 0; i < kKstaWStages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      int ws = 0, wph = 0;
      for (int item = item_begin, it = 0; item < item_end; ++item, ++it) {
        const int n = item / per_img, t0 = (item - n * per_img) * 2;
        const int cnt = min(2, tiles - t0);
        const int ab = it & 1;
        mbar_wait(a_empty + ab, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(a_full + ab, cnt * 16384u);
        for (int j = 0; j < cnt; ++j) {
          const int tj = t0 + j, ty = tj / p.tiles_x, tx = tj - ty * p.tiles_x;
          tma_load_4d(smem_a + (ab * 2 + j) * 16384, &p.tm_tile, a_full + ab, 0, tx * kTileW, ty * kTileH, p.a_slot * p.batch + n);
        }
        for (int tap = 0; tap < 25; ++tap) {
          mbar_wait(w_empty + ws, wph ^ 1);
          mbar_expect_tx(w_full + ws, 8192u);
          bulk_load(smem_w + ws * 8192, p.weights + tap * 8192, 8192u, w_full + ws);
          if (++ws == kKstaWStages) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = umma_idesc_f16(64, p.fmt);
    constexpr uint32_t hi = desc_hi_1024();
    const uint32_t a_lo0 = (smem_u32(smem_a) >> 4) & 0x3fffu, w_lo0 = (smem_u32(smem_w) >> 4) & 0x3fffu;
    int ws = 0, wph = 0;
    uint32_t use_bits = 0;
    for (int item = item_begin, it = 0; item < item_end; ++item, ++it) {
      const int n = item / per_img, t0 = (item - n * per_img) * 2;
      const int cnt = min(2, tiles - t0);
      const int ab = it & 1;
      mbar_wait(a_full + ab, (it >> 1) & 1);
      tc_fence_after();
      for (int tap = 0; tap < 25; ++tap) {
        mbar_wait(w_full + ws, wph);
        tc_fence_after();
        const uint32_t wl = w_lo0 + ws * (8192 >> 4);
        for (int j = 0; j < cnt; ++j) {
          const int acc = (tap & 3) * 2 + j;
          mbar_wait(t_empty + acc, ((use_bits >> acc) & 1u) ^ 1u);
          use_bits ^= 1u << acc;
          tc_fence_after();
          const uint32_t al = a_lo0 + (ab * 2 + j) * (16384 >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + acc * 64, make_desc64(hi, al + 2 * k), make_desc64(hi, wl + 2 * k), idesc, k ? 1u : 0u);
            umma_commit(t_full + acc);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(w_empty + ws);
        __syncwarp();
        if (++ws == kKstaWStages) { ws = 0; wph ^= 1; }
      }
      if (elect_one()) umma_commit(a_empty + ab);
      __syncwarp();
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    const int r = (warp - 2) >> 2;
    const int half = r & 1, tj = r >> 1;
    const int g = lane >> 2, q = lane & 3;
    const int ch0 = half * 32 + q * 8;
    const long npix = static_cast<long>(p.hp) * p.wp;
    const float slope = p.slope;
    uint32_t use_bits = 0;
    for (int item = item_begin; item < item_end; ++item) {
      const int n = item / per_img, t0 = (item - n * per_img) * 2;
      const int cnt = min(2, tiles - t0);
      if (tj >= cnt) continue;   // single-tile batch: the accumulators of tile 1 are not touched, their phases stay put
      const int tile = t0 + tj, ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
      const int px = tx * kTileW + g;
      const int py0 = ty * kTileH + quad * 4;
      const __nv_bfloat16* xs = p.arena + (static_cast<long>(p.x_slot) * p.batch + n) * npix * kC + ch0;
      float acc[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
      for (int tap = 0; tap < 25; ++tap) {
        const int a = (tap & 3) * 2 + tj;
        const int dy = tap / 5 - 2, dx = tap - (tap / 5) * 5 - 2;
        // neighbour features first: their latency hides behind the accumulator wait
        const int sx = min(max(px + dx, 0), p.w - 1);
        uint4 nb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int sy = min(max(py0 + j + dy, 0), p.h - 1);
          nb[j] = __ldg(reinterpret_cast<const uint4*>(xs + (static_cast<long>(sy) * p.wp + sx) * kC));
        }
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + tap * 64 + ch0));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + tap * 64 + ch0) + 1);
        const float bias[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        mbar_wait(t_full + a, (use_bits >> a) & 1u);
        use_bits ^= 1u << a;
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(a * 64 + half * 32);
        uint32_t ra[16], rb[16];
        tmem_ld_16x256b_x4(taddr, ra);
        tmem_ld_16x256b_x4(taddr + (16u << 16), rb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + a);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 2 * k + (e & 1);
            {
              const float t = __uint_as_float(ra[4 * k + e]) + bias[i];
              const float kv = fmaxf(t, slope * t);   // LeakyReLU for 0 <= slope <= 1
              const uint32_t w32 = i < 2 ? nb[e >> 1].x : i < 4 ? nb[e >> 1].y : i < 6 ? nb[e >> 1].z : nb[e >> 1].w;
              acc[e >> 1][i] += kv * ((i & 1) ? h_hi(w32, p.fmt) : h_lo(w32, p.fmt));
            }
            {
              const float t = __uint_as_float(rb[4 * k + e]) + bias[i];
              const float kv = fmaxf(t, slope * t);
              const uint32_t w32 = i < 2 ? nb[2 + (e >> 1)].x : i < 4 ? nb[2 + (e >> 1)].y : i < 6 ? nb[2 + (e >> 1)].z : nb[2 + (e >> 1)].w;
              acc[2 + (e >> 1)][i] += kv * ((i & 1) ? h_hi(w32, p.fmt) : h_lo(w32, p.fmt));
            }
          }
        }
      }
      // zeros outside the unpadded h x w region (like the unfused kernel); nothing outside the arena
      __nv_bfloat16* d = const_cast<__nv_bfloat16*>(p.arena) + (static_cast<long>(p.dst_slot) * p.batch + n) * npix * kC + ch0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int py = py0 + j;
        if (px < p.wp && py < p.hp) {
          const bool in = px < p.w && py < p.h;
          uint4 u;
          u.x = in ? pack_h2(acc[j][0], acc[j][1], p.fmt) : 0u;
          u.y = in ? pack_h2(acc[j][2], acc[j][3], p.fmt) : 0u;
          u.z = in ? pack_h2(acc[j][4], acc[j][5], p.fmt) : 0u;
          u.w = in ? pack_h2(acc[j][6], acc[j][7], p.fmt) : 0u;
          *reinterpret_cast<uint4*>(d + (static_cast<long>(py) * p.wp + px) * kC) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ bilinear corners
struct Corner4 {
  int off[4];   // element offset of the corner pixel inside the LR image (pixel index * 64), -1 = outside
  float wt[4];
};

// bilinear corners, zeros padding, align_corners = True (ATen grid_sampler_2d)
__device__ __forceinline__ Corner4 make_corners(float gx, float gy, int h, int w, int wp) {
  const float ix = unnormalize(gx, w), iy = unnormalize(gy, h);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = static_cast<int>(x0f), y0 = static_cast<int>(y0f);
  const float tx = ix - x0f, ty = iy - y0f;
  Corner4 c;
  const int xs[2] = {x0, x0 + 1}, ys[2] = {y0, y0 + 1};
  const float wx[2] = {1.f - tx, tx}, wy[2] = {1.f - ty, ty};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const bool in = xs[b] >= 0 && xs[b] < w && ys[a] >= 0 && ys[a] < h;
      c.off[a * 2 + b] = in ? (ys[a] * wp + xs[b]) * kC : -1;
      c.wt[a * 2 + b] = wy[a] * wx[b];
    }
  return c;
}

// ------------------------------------------------------------------------------------------------ fused HR kernel
// gather(x) -> routed experts -> + gather(sta) -> 128->64 fusion conv, one 128-pixel HR tile at a time, with the three
// small GEMMs on tcgen05 (accumulators in TMEM) and the gathers / routing on CUDA cores:
//   U[128,32] = F[128,64] Wc^T            (all 4 experts' compress matrices stacked, savsr_arch.py:353-355, 368)
//   V[e*8+k]  = r_e * sum_e' r_e' U[e'*8+k]                                     (routing, per pixel, in registers)
//   O[128,64] = V[128,32] We^T ; fea = O + F                                    (expand + residual, 357-370)
//   Y[128,64] = [S | fea][128,128] Wf^T + b                                     (fusion, 374; S = gathered sta)
// F, S, V, fea tiles are written by the threads as bf16 in the 128-byte-swizzled K-major layout the UMMA reads.
// The chain is sequential per tile; two CTAs per SM overlap each other's phases.
struct FusedParams {
  const __nv_bfloat16* lr;
  __nv_bfloat16* hr;
  const float* table;
  const float* base_y;
  const float* base_x;
  const uint8_t* w_compress;  // packed [32][64]  (4 KB)
  const uint8_t* w_expand;    // packed [64][64]  (8 KB, K columns 32..63 zero)
  const uint8_t* w_fusion;    // packed 2 x [64][64] (16 KB): sta block, fea block
  const float* bias;          // [64]
  int batch, hp, wp, h, w, H, W;
  int x_slot, sta_slot, dst_slot;
  int tiles_per_img;
  int fmt;
};

constexpr int kFusedThreads = 256;   // 8 warps: warp w owns TMEM lane quadrant w & 3 and column half w >> 2
constexpr int kFusedSmem = 1024 + 3 * 16384 + 4096 + 8192 + 16384 + 2 * 128 * 32 + 128 * 16 + 256 + 64;

__device__ __forceinline__ void gather8(const __nv_bfloat16* img, const Corner4& c, int chunk, float (&a)[8], int fmt) {
  uint4 v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = c.off[q] >= 0 ? *reinterpret_cast<const uint4*>(img + c.off[q] + chunk * 8) : make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = 0.f;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float wgt = c.wt[q];
    a[0] += wgt * h_lo(v[q].x, fmt); a[1] += wgt * h_hi(v[q].x, fmt); a[2] += wgt * h_lo(v[q].y, fmt); a[3] += wgt * h_hi(v[q].y, fmt);
    a[4] += wgt * h_lo(v[q].z, fmt); a[5] += wgt * h_hi(v[q].z, fmt); a[6] += wgt * h_lo(v[q].w, fmt); a[7] += wgt * h_hi(v[q].w, fmt);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&a)[8], int fmt) {
  uint4 o;
  o.x = pack_h2(a[0], a[1], fmt); o.y = pack_h2(a[2], a[3], fmt); o.z = pack_h2(a[4], a[5], fmt); o.w = pack_h2(a[6], a[7], fmt);
  return o;
}
// byte offset of 16-byte chunk `c` of row `r` in a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ int swz(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(kFusedThreads, 2) satu_fused_kernel(const FusedParams p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sF = smem;                 // F, later fea (in place)      [128][128 B]
  uint8_t* sS = smem + 16384;         // gathered sta                 [128][128 B]
  uint8_t* sV = smem + 32768;         // V (K = 32 used)              [128][128 B]
  uint8_t* sWc = smem + 49152;        // [32][128 B]
  uint8_t* sWe = sWc + 4096;          // [64][128 B]
  uint8_t* sWf = sWe + 8192;          // 2 x [64][128 B]
  Corner4* cx = reinterpret_cast<Corner4*>(sWf + 16384);
  Corner4* cs = cx + 128;
  float* rt = reinterpret_cast<float*>(cs + 128);   // [128][4] routing weights
  float* sbias = rt + 128 * 4;                      // [64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sbias + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;                 // tile row (= TMEM lane) this thread serves in the register phases
  for (int i = tid; i < (4096 + 8192 + 16384) / 16; i += kFusedThreads) {
    const uint8_t* src = i < 256 ? p.w_compress + i * 16 : i < 768 ? p.w_expand + (i - 256) * 16 : p.w_fusion + (i - 768) * 16;
    *reinterpret_cast<uint4*>(sWc + i * 16) = *reinterpret_cast<const uint4*>(src);
  }
  // zero the V tile once (chunks 4..7 of every row, the unused K half, are never written again)
  for (int i = tid; i < 1024; i += kFusedThreads) *reinterpret_cast<uint4*>(sV + i * 16) = make_uint4(0, 0, 0, 0);
  if (tid < 64) sbias[tid] = __ldg(p.bias + tid);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const uint32_t tU = tm, tO = tm + 32, tY = tm + 96;
  const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
  constexpr uint32_t hi = desc_hi_1024();
  const uint32_t loF = (smem_u32(sF) >> 4) & 0x3fff, loS = (smem_u32(sS) >> 4) & 0x3fff, loV = (smem_u32(sV) >> 4) & 0x3fff;
  const uint32_t loWc = (smem_u32(sWc) >> 4) & 0x3fff, loWe = (smem_u32(sWe) >> 4) & 0x3fff, loWf = (smem_u32(sWf) >> 4) & 0x3fff;
  uint32_t phase = 0;
  const long NPIX = static_cast<long>(p.H) * p.W;
  const long lr_img = static_cast<long>(p.hp) * p.wp * kC;
  const int total = p.batch * p.tiles_per_img;

  for (int t = blockIdx.x; t < total; t += gridDim.x) {
    const int n = t / p.tiles_per_img;
    const long pix0 = static_cast<long>(t - n * p.tiles_per_img) * 128;
    // ---- P0: per-pixel sampling corners and routing weights (threads 0..127, one pixel each)
    if (tid < 128) {
      const long pix = pix0 + tid;
      if (pix < NPIX) {
        const int i = pix / p.W, j = pix - static_cast<long>(i) * p.W;
        const float4 t0 = *reinterpret_cast<const float4*>(p.table + pix * 8);
        const float4 t1 = *reinterpret_cast<const float4*>(p.table + pix * 8 + 4);
        const float bx = p.base_x[j], by = p.base_y[i];
        const float wm1 = static_cast<float>(p.w - 1), hm1 = static_cast<float>(p.h - 1);
        cx[tid] = make_corners(__fadd_rn(bx, __fdiv_rn(__fmul_rn(t0.x, 2.f), wm1)), __fadd_rn(by, __fdiv_rn(__fmul_rn(t0.y, 2.f), hm1)), p.h, p.w, p.wp);
        cs[tid] = make_corners(__fadd_rn(bx, __fdiv_rn(__fmul_rn(t0.z, 2.f), wm1)), __fadd_rn(by, __fdiv_rn(__fmul_rn(t0.w, 2.f), hm1)), p.h, p.w, p.wp);
        *reinterpret_cast<float4*>(rt + tid * 4) = t1;
      } else {
        Corner4 z;
#pragma unroll
        for (int q = 0; q < 4; ++q) { z.off[q] = -1; z.wt[q] = 0.f; }
        cx[tid] = z; cs[tid] = z;
        *reinterpret_cast<float4*>(rt + tid * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    // ---- P1: bilinear gathers, (pixel, 8-channel chunk) per thread-iteration: 8 lanes read one 128-byte pixel row
    const __nv_bfloat16* xs = p.lr + (static_cast<long>(p.x_slot) * p.batch + n) * lr_img;
    const __nv_bfloat16* ss = p.lr + (static_cast<long>(p.sta_slot) * p.batch + n) * lr_img;
#pragma unroll 2
    for (int it = 0; it < 4; ++it) {
      const int id = it * kFusedThreads + tid;
      const int lp = id >> 3, chunk = id & 7;
      float a[8], b[8];
      gather8(xs, cx[lp], chunk, a, p.fmt);
      gather8(ss, cs[lp], chunk, b, p.fmt);
      *reinterpret_cast<uint4*>(sF + swz(lp, chunk)) = pack8(a, p.fmt);
      *reinterpret_cast<uint4*>(sS + swz(lp, chunk)) = pack8(b, p.fmt);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    // ---- P2: U = F Wc^T   (M 128, N 32, K 64)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tU, make_desc64(hi, loF + 2 * k), make_desc64(hi, loWc + 2 * k), umma_idesc_f16(32, p.fmt), k ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    // ---- P3: routing in registers -> V tile (warps 0-3)
    if (half == 0) {
      uint32_t u0[16], u1[16];
      tmem_ld16(tU + lane_addr, u0);
      tmem_ld16(tU + lane_addr + 16, u1);
      tmem_ld_wait();
      const float4 r = *reinterpret_cast<const float4*>(rt + row * 4);
      float tk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        tk[k] = r.x * __uint_as_float(u0[k]) + r.y * __uint_as_float(u0[8 + k]) + r.z * __uint_as_float(u1[k]) + r.w * __uint_as_float(u1[8 + k]);
      const float re[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = re[e] * tk[k];
        *reinterpret_cast<uint4*>(sV + swz(row, e)) = pack8(v, p.fmt);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    // ---- P4: O = V We^T   (M 128, N 64, K 32)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_bf16(tO, make_desc64(hi, loV + 2 * k), make_desc64(hi, loWe + 2 * k), umma_idesc_f16(64, p.fmt), k ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    // ---- P5: fea = O + F, in place over the F tile (each warp: 32 rows x 32 columns)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t o[16];
      tmem_ld16(tO + lane_addr + half * 32 + 16 * j, o);
      tmem_ld_wait();
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint4* ptr = reinterpret_cast<uint4*>(sF + swz(row, half * 4 + 2 * j + c2));
        const uint4 f = *ptr;
        float v[8];
        v[0] = __uint_as_float(o[8 * c2 + 0]) + h_lo(f.x, p.fmt); v[1] = __uint_as_float(o[8 * c2 + 1]) + h_hi(f.x, p.fmt);
        v[2] = __uint_as_float(o[8 * c2 + 2]) + h_lo(f.y, p.fmt); v[3] = __uint_as_float(o[8 * c2 + 3]) + h_hi(f.y, p.fmt);
        v[4] = __uint_as_float(o[8 * c2 + 4]) + h_lo(f.z, p.fmt); v[5] = __uint_as_float(o[8 * c2 + 5]) + h_hi(f.z, p.fmt);
        v[6] = __uint_as_float(o[8 * c2 + 6]) + h_lo(f.w, p.fmt); v[7] = __uint_as_float(o[8 * c2 + 7]) + h_hi(f.w, p.fmt);
        *ptr = pack8(v, p.fmt);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    // ---- P6: Y = [S | fea] Wf^T   (M 128, N 64, K 128)
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tY, make_desc64(hi, loS + 2 * k), make_desc64(hi, loWf + 2 * k), umma_idesc_f16(64, p.fmt), k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tY, make_desc64(hi, loF + 2 * k), make_desc64(hi, loWf + 512 + 2 * k), umma_idesc_f16(64, p.fmt), 1u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase); phase ^= 1;
    tc_fence_after();
    // ---- P7: + bias -> 16-bit HR feature.  16x256b TMEM loads over SAVSR_ROWS_QUAD-ordered fusion weights: thread
    // (g, q) holds pixels g + 8 j of its quadrant and the 8 consecutive channels half * 32 + 8 q .., so a store
    // instruction writes 64 contiguous bytes per thread quad, 8 lines per warp (one pixel per lane touched 32 lines).
    {
      const int g = lane >> 2, q = lane & 3;
      uint32_t ya[16], yb[16];
      tmem_ld_16x256b_x4(tY + lane_addr + half * 32, ya);
      tmem_ld_16x256b_x4(tY + lane_addr + (16u << 16) + half * 32, yb);
      tmem_ld_wait();
      const int ch0 = half * 32 + q * 8;
      float bias8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) bias8[i] = sbias[ch0 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long pix = pix0 + quad * 32 + g + 8 * j;
        if (pix < NPIX) {
          const uint32_t* y = j < 2 ? ya : yb;
          float v[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            v[2 * k] = __uint_as_float(y[4 * k + 2 * (j & 1)]) + bias8[2 * k];
            v[2 * k + 1] = __uint_as_float(y[4 * k + 2 * (j & 1) + 1]) + bias8[2 * k + 1];
          }
          *reinterpret_cast<uint4*>(p.hr + ((static_cast<long>(p.dst_slot) * p.batch + n) * NPIX + pix) * kC + ch0) = pack8(v, p.fmt);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // the next tile overwrites cx/cs and the F/S/V tiles
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

}  // namespace savsr

using namespace savsr;

extern "C" int savsr_satu_index(savsr_ctx* ctx, const savsr_satu_weights* wts, int h, int w, int H, int W, float s_h,
                                float s_w, float* rel_y, float* rel_x, int32_t* cell_y, int32_t* cell_x, float* base_y,
                                float* base_x, int32_t* corner_y, int32_t* corner_x, float* table, savsr_stream st_) {
  SAVSR_REQUIRE(ctx, "savsr_satu_index: null context");
  DeviceGuard guard(ctx->device);
  SAVSR_REQUIRE(h >= 2 && w >= 2 && H >= 1 && W >= 1, "savsr_satu_index: bad sizes lr %dx%d hr %dx%d", h, w, H, W);
  SAVSR_REQUIRE(s_h > 0.f && s_w > 0.f, "savsr_satu_index: scale must be positive, got (%g, %g)", s_h, s_w);
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  satu_axis_kernel<<<(H + 127) / 128, 128, 0, st>>>(H, h, s_h, rel_y, cell_y, base_y, corner_y);
  satu_axis_kernel<<<(W + 127) / 128, 128, 0, st>>>(W, w, s_w, rel_x, cell_x, base_x, corner_x);
  if (table) {
    SAVSR_REQUIRE(wts && wts->body0_w && wts->body0_b && wts->body2_w && wts->body2_b && wts->routing_w && wts->routing_b &&
                  wts->offset_w && wts->offset_b && wts->st_offset_w && wts->st_offset_b,
                  "savsr_satu_index: table requested but MLP weights missing");
    const long npix = static_cast<long>(H) * W;
    satu_table_kernel<<<static_cast<unsigned>((npix + 127) / 128), 128, 0, st>>>(*wts, H, W, s_h, s_w, table);
  }
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_satu_kconv_sta(savsr_ctx* ctx, savsr_arena* arena, int a_slot, int x_slot, int dst_slot, int h, int w,
                                    const void* weights, const float* bias, float slope, savsr_stream st) {
  SAVSR_REQUIRE(ctx && arena && weights && bias, "savsr_satu_kconv_sta: null pointer");
  SAVSR_REQUIRE(a_slot >= 0 && a_slot < arena->nslots && x_slot >= 0 && x_slot < arena->nslots && dst_slot >= 0 &&
                dst_slot < arena->nslots, "savsr_satu_kconv_sta: slot out of range");
  SAVSR_REQUIRE(dst_slot != a_slot && dst_slot != x_slot, "savsr_satu_kconv_sta: the destination slot must differ from both sources");
  SAVSR_REQUIRE(h >= 1 && w >= 1 && h <= arena->height && w <= arena->width, "savsr_satu_kconv_sta: region %dx%d exceeds arena", h, w);
  SAVSR_REQUIRE(slope >= 0.f && slope <= 1.f, "savsr_satu_kconv_sta: LeakyReLU slope %g outside [0, 1]", slope);
  if (arena->batch == 0) return 0;
  KstaParams p;
  memset(&p, 0, sizeof(p));
  p.tm_tile = arena->tm_tile;
  p.arena = arena->base;
  p.weights = static_cast<const uint8_t*>(weights);
  p.bias = bias;
  p.slope = slope;
  p.batch = arena->batch; p.hp = arena->height; p.wp = arena->width; p.h = h; p.w = w;
  p.a_slot = a_slot; p.x_slot = x_slot; p.dst_slot = dst_slot;
  p.tiles_x = arena->tiles_x; p.tiles_y = arena->tiles_y;
  const int tiles = p.tiles_x * p.tiles_y;
  p.nbatches = arena->batch * ((tiles + 1) / 2);
  p.chunk = (p.nbatches + ctx->sm_count - 1) / ctx->sm_count;
  p.fmt = ctx->fmt;
  const int grid = (p.nbatches + p.chunk - 1) / p.chunk;
  DeviceGuard guard(ctx->device);
  if (int rc = ensure_smem_attr(ctx, kAttrKsta, satu_kconv_sta_kernel, kKstaSmem)) return rc;
  satu_kconv_sta_kernel<<<grid, kKstaThreads, kKstaSmem, static_cast<cudaStream_t>(st)>>>(p);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_satu_fused(savsr_ctx* ctx, savsr_arena* lr, int x_slot, int sta_slot, int h, int w, savsr_arena* hr, int dst_slot,
                                const float* table, const float* base_y, const float* base_x, const void* w_compress,
                                const void* w_expand, const void* w_fusion, const float* fusion_bias, savsr_stream st) {
  SAVSR_REQUIRE(ctx && lr && hr && table && base_y && base_x && w_compress && w_expand && w_fusion && fusion_bias, "savsr_satu_fused: null pointer");
  SAVSR_REQUIRE(lr->batch == hr->batch, "savsr_satu_fused: LR batch %d != HR batch %d", lr->batch, hr->batch);
  SAVSR_REQUIRE(x_slot >= 0 && x_slot < lr->nslots && sta_slot >= 0 && sta_slot < lr->nslots, "savsr_satu_fused: LR slot out of range");
  SAVSR_REQUIRE(dst_slot >= 0 && dst_slot < hr->nslots, "savsr_satu_fused: HR slot out of range");
  SAVSR_REQUIRE(h >= 2 && w >= 2 && h <= lr->height && w <= lr->width, "savsr_satu_fused: region %dx%d exceeds LR arena", h, w);
  if (lr->batch == 0) return 0;
  DeviceGuard guard(ctx->device);
  if (int rc = ensure_smem_attr(ctx, kAttrFused, satu_fused_kernel, kFusedSmem)) return rc;
  FusedParams p;
  p.lr = lr->base; p.hr = hr->base; p.table = table; p.base_y = base_y; p.base_x = base_x;
  p.w_compress = static_cast<const uint8_t*>(w_compress); p.w_expand = static_cast<const uint8_t*>(w_expand);
  p.w_fusion = static_cast<const uint8_t*>(w_fusion); p.bias = fusion_bias;
  p.batch = lr->batch; p.hp = lr->height; p.wp = lr->width; p.h = h; p.w = w; p.H = hr->height; p.W = hr->width;
  p.x_slot = x_slot; p.sta_slot = sta_slot; p.dst_slot = dst_slot; p.fmt = ctx->fmt;
  const long npix = static_cast<long>(p.H) * p.W;
  p.tiles_per_img = static_cast<int>((npix + 127) / 128);
  const int total = p.batch * p.tiles_per_img;
  const int grid = total < 2 * ctx->sm_count ? total : 2 * ctx->sm_count;
  satu_fused_kernel<<<grid, kFusedThreads, kFusedSmem, static_cast<cudaStream_t>(st)>>>(p);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}
