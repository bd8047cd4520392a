// C-ABI plumbing: error reporting, device context, activation arenas and their TMA descriptors.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace savsr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return 2;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 4-D NHWC bf16 map: dims (C = 64, W, H, images), box (64, bw, bh, 1), 128-byte swizzle, zero OOB fill.
static int encode_map(savsr_ctx* ctx, CUtensorMap* tm, void* base, int nimg, int height, int width, int bw, int bh) {
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(height), static_cast<cuuint64_t>(nimg)};
  const cuuint64_t strides[3] = {128, static_cast<cuuint64_t>(width) * 128, static_cast<cuuint64_t>(height) * width * 128};
  const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(bw), static_cast<cuuint32_t>(bh), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (images %d, %dx%d, box %dx%d)", static_cast<int>(r), nimg, height,
              width, bw, bh);
    return 3;
  }
  return 0;
}

}  // namespace savsr

using namespace savsr;

extern "C" int savsr_abi_version(void) { return SAVSR_ABI_VERSION; }
extern "C" const char* savsr_last_error(void) { return g_err; }

extern "C" int savsr_ctx_create(int device, savsr_ctx** out) {
  SAVSR_REQUIRE(out, "savsr_ctx_create: null output pointer");
  *out = nullptr;
  int count = 0;
  SAVSR_CUDA(cudaGetDeviceCount(&count));
  SAVSR_REQUIRE(device >= 0 && device < count, "savsr_ctx_create: device %d not present (%d CUDA devices)", device, count);
  cudaDeviceProp prop;
  SAVSR_CUDA(cudaGetDeviceProperties(&prop, device));
  SAVSR_REQUIRE(prop.major == 10, "savsr_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only "
                "(tcgen05/TMEM/TMA) and has no fallback", device, prop.major, prop.minor);
  DeviceGuard guard(device);   // the caller's current device is restored on return
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SAVSR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SAVSR_REQUIRE(qres == cudaDriverEntryPointSuccess && fn, "savsr_ctx_create: driver lacks cuTensorMapEncodeTiled");
  savsr_ctx* c = static_cast<savsr_ctx*>(calloc(1, sizeof(savsr_ctx)));
  SAVSR_REQUIRE(c, "savsr_ctx_create: out of host memory");
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->encode_tiled = fn;
  c->opt[SAVSR_OPT_BIGK_ALL] = 1;
  c->opt[SAVSR_OPT_BIGK_ISSUERS] = 2;
  *out = c;
  return 0;
}

extern "C" int savsr_ctx_set_format(savsr_ctx* ctx, int format) {
  SAVSR_REQUIRE(ctx, "savsr_ctx_set_format: null context");
  SAVSR_REQUIRE(format == SAVSR_FMT_BF16 || format == SAVSR_FMT_FP16, "savsr_ctx_set_format: unknown format %d", format);
  ctx->fmt = format;
  return 0;
}
extern "C" int savsr_ctx_get_format(const savsr_ctx* ctx) { return ctx ? ctx->fmt : -1; }

extern "C" int savsr_ctx_set_option(savsr_ctx* ctx, int option, int value) {
  SAVSR_REQUIRE(ctx, "savsr_ctx_set_option: null context");
  SAVSR_REQUIRE(option >= 0 && option < SAVSR_OPT_COUNT, "savsr_ctx_set_option: unknown option %d", option);
  if (option == SAVSR_OPT_BIGK_ALL) SAVSR_REQUIRE(value == 0 || value == 1, "savsr_ctx_set_option: BIGK_ALL must be 0 or 1, got %d", value);
  if (option == SAVSR_OPT_BIGK_ISSUERS) SAVSR_REQUIRE(value == 1 || value == 2, "savsr_ctx_set_option: BIGK_ISSUERS must be 1 or 2, got %d", value);
  if (option == SAVSR_OPT_PDL) SAVSR_REQUIRE(value == 0 || value == 1, "savsr_ctx_set_option: PDL must be 0 or 1, got %d", value);
  ctx->opt[option] = value;
  return 0;
}
extern "C" int savsr_ctx_get_option(const savsr_ctx* ctx, int option) {
  return (ctx && option >= 0 && option < SAVSR_OPT_COUNT) ? ctx->opt[option] : -1;
}

extern "C" void savsr_ctx_destroy(savsr_ctx* ctx) { free(ctx); }
extern "C" int savsr_ctx_sm_count(const savsr_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" size_t savsr_arena_bytes(int nslots, int batch, int height, int width) {
  if (nslots <= 0 || batch <= 0 || height <= 0 || width <= 0) return 0;
  return static_cast<size_t>(nslots) * batch * height * width * kC * sizeof(__nv_bfloat16);
}

extern "C" int savsr_arena_create(savsr_ctx* ctx, void* base, int nslots, int batch, int height, int width, savsr_arena** out) {
  SAVSR_REQUIRE(ctx && out, "savsr_arena_create: null pointer");
  *out = nullptr;
  DeviceGuard guard(ctx->device);
  SAVSR_REQUIRE(base && (reinterpret_cast<uintptr_t>(base) & 255) == 0, "savsr_arena_create: base must be a 256-byte aligned device pointer");
  SAVSR_REQUIRE(nslots > 0 && batch > 0 && height > 0 && width > 0, "savsr_arena_create: empty arena (%d slots, batch %d, %dx%d)", nslots,
                batch, height, width);
  savsr_arena* a = static_cast<savsr_arena*>(calloc(1, sizeof(savsr_arena)));
  SAVSR_REQUIRE(a, "savsr_arena_create: out of host memory");
  a->ctx = ctx;
  a->base = static_cast<__nv_bfloat16*>(base);
  a->nslots = nslots; a->batch = batch; a->height = height; a->width = width;
  a->tiles_x = (width + kTileW - 1) / kTileW;
  a->tiles_y = (height + kTileH - 1) / kTileH;
  int rc = encode_map(ctx, &a->tm_tile, base, nslots * batch, height, width, kTileW, kTileH);
  if (rc == 0) rc = encode_map(ctx, &a->tm_halo, base, nslots * batch, height, width, kTileW + 2, kTileH + 2);
  if (rc != 0) { free(a); return rc; }
  *out = a;
  return 0;
}

extern "C" void savsr_arena_destroy(savsr_arena* a) { free(a); }
extern "C" int savsr_arena_tiles(const savsr_arena* a) { return a ? a->tiles_x * a->tiles_y : 0; }
