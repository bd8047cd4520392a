// Weight gradient of the 3x3 convolution on tcgen05 (row f1 of the scope table: backward of the conv that carries 98 % of the
// FLOPs; lbasicsr/models/sr_model.py:101-128 runs it through autograd / cuDNN wgrad).
//
//   dW[o][i][ky][kx] = sum over samples n and pixels p of  dY[n][o][p] * X[n][i][p + (ky-1, kx-1)]        (zero padding)
//
// The contraction index is the PIXEL, so both operands must have pixels contiguous: the kernel reads 16-bit NCHW tensors
// (the layout autograd hands over anyway), where a TMA box of 64 channels x 64 consecutive pixels of one image row lands in shared
// memory as a [64 rows][128 B] K-major SWIZZLE_128B tile -- the same operand layout every other kernel of this library uses.
// TMA moves 16-byte granules, so a box cannot start one 2-byte pixel to the left or right (measured: illegal instruction); the
// caller therefore passes X as three copies shifted by -1 / 0 / +1 pixel along x.  The row shifts are plain box coordinates.
//   A (M = 128): two X tiles stacked (two taps of the same image row: dx = -1 | 0, and dx = +1 | an unused half), rows = input channel
//   B (N = 64) : the dY tile, rows = output channel;   D[tap pair][i][o] accumulates in TMEM over the CTA's whole work range.
// Per image row y the three input rows y-1, y, y+1 (each as three x-shifted tiles, TMA zero fill = the conv's zero padding) are kept
// in a 4-slot ring, so a row is fetched once per 3 uses.  6 accumulators x 64 columns; flushed with fp32 atomics when the weight
// block (source slot / sample) changes or the CTA ends.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = flush (TMEM lane quadrant = warp % 4).
#include "common.cuh"

namespace savsr {

struct WgradParams {
  CUtensorMap tm_x;      // [3 x-shifts][B*Ci][H][Wp] 16-bit, box (64 px, 1 row, 64 channels)
  CUtensorMap tm_dy;     // [B*64][H][Wp]
  float* dw;             // [per_sample ? B : 1][64][Ci][3][3] fp32, accumulated atomically (caller zeroes it)
  int batch, ci, height, width;     // ci = 64 * nsrc; width = real width (pixels >= width are zero in both tensors)
  int nsrc, nxseg, nychunk, rows_per_chunk;
  int per_sample;
  int nitems, chunk;     // work items (source, sample, x segment, row chunk), `chunk` consecutive ones per CTA
  int fmt;
};

constexpr int kWgRowSlots = 4;
constexpr int kWgRowBytes = 4 * 8192;                  // dx = -1 | 0 | +1 | unused half of the second M = 128 operand
constexpr int kWgDyStages = 2;
constexpr int kWgThreads = 6 * 32;
constexpr int kWgSmem = 1024 + kWgRowSlots * kWgRowBytes + kWgDyStages * 8192 + 256;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int x, int y, int c) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(c), "r"(smem_u32(bar))
      : "memory");
}

struct WgItem { int s, n, x0, y0, y1; };
__device__ __forceinline__ WgItem wg_item(const WgradParams& p, int item) {
  WgItem it;
  const int yc = item % p.nychunk; item /= p.nychunk;
  const int xs = item % p.nxseg; item /= p.nxseg;
  it.n = item % p.batch;
  it.s = item / p.batch;
  it.x0 = xs * 64;
  it.y0 = yc * p.rows_per_chunk;
  it.y1 = min(it.y0 + p.rows_per_chunk, p.height);
  return it;
}

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_rows = smem;
  uint8_t* s_dy = smem + kWgRowSlots * kWgRowBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dy + kWgDyStages * 8192);
  uint64_t* row_full = bars;                    // [4]
  uint64_t* row_empty = bars + 4;               // [4]
  uint64_t* dy_full = bars + 8;                 // [2]
  uint64_t* dy_empty = bars + 10;               // [2]
  uint64_t* acc_full = bars + 12;               // issuer -> flush warps
  uint64_t* acc_empty = bars + 13;              // flush warps -> issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item_begin = blockIdx.x * p.chunk, item_end = min(item_begin + p.chunk, p.nitems);
  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.tm_x); prefetch_tensormap(&p.tm_dy);
    for (int i = 0; i < 4; ++i) { mbar_init(row_full + i, 1); mbar_init(row_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(dy_full + i, 1); mbar_init(dy_empty + i, 1); }
    mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t rcount = 0, dcount = 0;                 // rows / dY tiles loaded so far (ring positions)
      for (int item = item_begin; item < item_end; ++item) {
        const WgItem w = wg_item(p, item);
        const int cx = w.n * p.ci + w.s * 64, cy = w.n * 64;
        // input rows y0-1 .. y1 (inclusive) in order; the dY row y is issued right after input row y+1 (what the MMAs of y need last)
        for (int r = w.y0 - 1; r <= w.y1; ++r, ++rcount) {
          const int slot = rcount & 3;
          mbar_wait(row_empty + slot, ((rcount >> 2) & 1u) ^ 1u);
          mbar_expect_tx(row_full + slot, 3 * 8192u);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx)
            tma_load_3d(s_rows + slot * kWgRowBytes + dx * 8192, &p.tm_x, row_full + slot, w.x0, r, dx * p.batch * p.ci + cx);   // out-of-range rows read as zero
          const int y = r - 1;
          if (y >= w.y0) {
            const int st = dcount & 1;
            mbar_wait(dy_empty + st, ((dcount >> 1) & 1u) ^ 1u);
            mbar_expect_tx(dy_full + st, 8192u);
            tma_load_3d(s_dy + st * 8192, &p.tm_dy, dy_full + st, w.x0, y, cy);
            ++dcount;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    const uint32_t idesc = umma_idesc_f16(64, p.fmt);
    constexpr uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lo_rows = (smem_u32(s_rows) >> 4) & 0x3fffu, lo_dy = (smem_u32(s_dy) >> 4) & 0x3fffu;
    uint32_t rcount = 0, dcount = 0, flushes = 0;
    int cur_key = -1;
    bool fresh = true;                                  // the accumulators hold nothing yet
    for (int item = item_begin; item < item_end; ++item) {
      const WgItem w = wg_item(p, item);
      const int key = w.s * p.batch + (p.per_sample ? w.n : 0);
      if (key != cur_key && cur_key >= 0) {
        // another weight block starts: hand the accumulators to the flush warps and wait until they have been read
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
        mbar_wait(acc_empty, flushes & 1u);
        ++flushes;
        tc_fence_after();
        fresh = true;
      }
      cur_key = key;
      // rows of this item occupy ring positions rcount .. rcount + (y1 - y0 + 1); row r sits at position rcount + (r - (y0 - 1))
      for (int y = w.y0; y < w.y1; ++y) {
        const uint32_t base = rcount + (y - w.y0);      // position of input row y - 1
        if (y == w.y0) {                                // first image row of the item: its three input rows are all new
          mbar_wait(row_full + (base & 3), (base >> 2) & 1u);
          mbar_wait(row_full + ((base + 1) & 3), ((base + 1) >> 2) & 1u);
        }
        mbar_wait(row_full + ((base + 2) & 3), ((base + 2) >> 2) & 1u);
        const int st = dcount & 1;
        mbar_wait(dy_full + st, (dcount >> 1) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t bl = lo_dy + st * (8192 >> 4);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint32_t al = lo_rows + ((base + dy) & 3) * (kWgRowBytes >> 4);
#pragma unroll
            for (int pair = 0; pair < 2; ++pair) {
              const uint32_t d = tm + static_cast<uint32_t>((dy * 2 + pair) * 64);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d, (static_cast<uint64_t>(hi) << 32) | (al + pair * (16384 >> 4) + 2 * k), (static_cast<uint64_t>(hi) << 32) | (bl + 2 * k), idesc,
                          (fresh && k == 0) ? 0u : 1u);
            }
          }
          umma_commit(dy_empty + st);
          umma_commit(row_empty + (base & 3));          // input row y - 1 is not needed after image row y
          if (y == w.y1 - 1) {                          // last image row of the item: rows y and y + 1 retire too
            umma_commit(row_empty + ((base + 1) & 3));
            umma_commit(row_empty + ((base + 2) & 3));
          }
        }
        __syncwarp();
        fresh = false;
        ++dcount;
      }
      rcount += static_cast<uint32_t>(w.y1 - w.y0 + 2);
    }
    if (cur_key >= 0) {
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ================================ flush warps ================================
    const int quad = warp & 3;                          // warps 2..5 -> TMEM lane quadrants 2, 3, 0, 1
    const int row = quad * 32 + lane;                   // accumulator row: rows 0..63 = first tap of the pair, 64..127 = second
    int cur_key = -1, cur_s = 0, cur_n = 0;
    uint32_t flushes = 0;
    auto flush = [&]() {
      mbar_wait(acc_full, flushes & 1u);
      ++flushes;
      tc_fence_after();
      float* dwb = p.dw + (p.per_sample ? static_cast<long>(cur_n) * 64 * p.ci * 9 : 0);
      const int i = cur_s * 64 + (row & 63);
#pragma unroll 1
      for (int acc = 0; acc < 6; ++acc) {
        const int dy = acc >> 1, pair = acc & 1;
        const int dx = pair == 0 ? (row < 64 ? 0 : 1) : (row < 64 ? 2 : -1);     // tiles of a row slot: dx index 0,1,2 = shift -1,0,+1; 4th half unused
        uint32_t v[16];
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
          tmem_ld16(tm + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * 64 + c0), v);
          tmem_ld_wait();
          if (dx >= 0) {
#pragma unroll
            for (int c = 0; c < 16; ++c) atomicAdd(dwb + (static_cast<long>(c0 + c) * p.ci + i) * 9 + dy * 3 + dx, __uint_as_float(v[c]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    };
    for (int item = item_begin; item < item_end; ++item) {
      const WgItem w = wg_item(p, item);
      const int key = w.s * p.batch + (p.per_sample ? w.n : 0);
      if (key != cur_key && cur_key >= 0) flush();
      cur_key = key; cur_s = w.s; cur_n = w.n;
    }
    if (cur_key >= 0) flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D map over a 16-bit NCHW tensor viewed as [planes][H][pitch]: box (64 pixels, 1 row, 64 planes), 128-byte swizzle, zero fill.
static int encode_nchw_map(savsr_ctx* ctx, CUtensorMap* tm, const void* base, long planes, int height, int pitch) {
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(pitch), static_cast<cuuint64_t>(height), static_cast<cuuint64_t>(planes)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(pitch) * 2, static_cast<cuuint64_t>(pitch) * height * 2};
  const cuuint32_t box[3] = {64, 1, 64};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = reinterpret_cast<EncodeTiledFn3>(ctx->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("savsr_conv_wgrad: cuTensorMapEncodeTiled failed with CUresult %d (planes %ld, %dx%d)", static_cast<int>(r), planes, height, pitch);
    return 3;
  }
  return 0;
}

}  // namespace savsr

using namespace savsr;

extern "C" int savsr_conv_wgrad(savsr_ctx* ctx, const void* x3_nchw16, const void* dy_nchw16, int batch, int ci, int height, int width,
                                int pitch, int per_sample, float* dw, savsr_stream st) {
  SAVSR_REQUIRE(ctx && x3_nchw16 && dy_nchw16 && dw, "savsr_conv_wgrad: null pointer");
  DeviceGuard guard(ctx->device);
  SAVSR_REQUIRE(batch >= 1 && height >= 1 && width >= 1, "savsr_conv_wgrad: empty problem (batch %d, %dx%d)", batch, height, width);
  SAVSR_REQUIRE(ci > 0 && ci % 64 == 0 && ci <= 64 * SAVSR_MAX_SRC, "savsr_conv_wgrad: ci (%d) must be 64, 128, ... %d", ci, 64 * SAVSR_MAX_SRC);
  SAVSR_REQUIRE(pitch >= width && pitch % 8 == 0, "savsr_conv_wgrad: row pitch %d must be >= width %d and a multiple of 8 elements", pitch, width);
  SAVSR_REQUIRE((reinterpret_cast<uintptr_t>(x3_nchw16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy_nchw16) & 15) == 0,
                "savsr_conv_wgrad: tensors must be 16-byte aligned");
  WgradParams p;
  memset(&p, 0, sizeof(p));
  if (int rc = encode_nchw_map(ctx, &p.tm_x, x3_nchw16, 3L * batch * ci, height, pitch)) return rc;
  if (int rc = encode_nchw_map(ctx, &p.tm_dy, dy_nchw16, static_cast<long>(batch) * 64, height, pitch)) return rc;
  p.dw = dw;
  p.batch = batch; p.ci = ci; p.height = height; p.width = width;
  p.nsrc = ci / 64;
  p.nxseg = (width + 63) / 64;
  p.per_sample = per_sample ? 1 : 0;
  // row chunks sized so that the launch has about two work items per SM (each item re-reads two halo rows)
  const int base_items = p.nsrc * batch * p.nxseg;
  int chunks = (2 * ctx->sm_count + base_items - 1) / base_items;
  if (chunks > height) chunks = height;
  if (chunks < 1) chunks = 1;
  p.rows_per_chunk = (height + chunks - 1) / chunks;
  p.nychunk = (height + p.rows_per_chunk - 1) / p.rows_per_chunk;
  p.nitems = base_items * p.nychunk;
  p.chunk = (p.nitems + ctx->sm_count - 1) / ctx->sm_count;
  p.fmt = ctx->fmt;
  const int grid = (p.nitems + p.chunk - 1) / p.chunk;
  if (int rc = ensure_smem_attr(ctx, kAttrWgrad, conv_wgrad_kernel, kWgSmem)) return rc;
  conv_wgrad_kernel<<<grid, kWgThreads, kWgSmem, static_cast<cudaStream_t>(st)>>>(p);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}
