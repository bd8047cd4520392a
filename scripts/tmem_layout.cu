// Probe of the tcgen05.ld .16x256b register layout: fill TMEM with value = lane * 1000 + column through .32x32b stores
// (thread i <-> TMEM lane i), read back with .16x256b.x4 at lane offsets 0 and 16 and print which (lane, column) each
// thread register received.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tmem_layout.bin scripts/tmem_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../savsr_b200/csrc/common.cuh"
namespace savsr { void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*) { return 1; } }
using namespace savsr;

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<64>(&slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t mine = tm + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = (warp * 32 + lane) * 1000 + c0 + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(mine + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    const uint32_t addr = mine + (static_cast<uint32_t>(h * 16) << 16) + 32;   // columns 32..63
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(addr));
    tmem_ld_wait();
    for (int c = 0; c < 16; ++c) out[((warp * 2 + h) * 32 + lane) * 16 + c] = r[c];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tm); }
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int hf = 0; hf < 2; ++hf) for (int t = 0; t < 32; ++t) for (int c = 0; c < 16; ++c) {
    const uint32_t v = h[((w * 2 + hf) * 32 + t) * 16 + c];
    const int k = c >> 2, e2 = c & 3;   // hypothesis: reg 4k+e: row t/4 + 8*(e>>1), column 8k + 2*(t%4) + (e&1)
    const uint32_t expect = (w * 32 + hf * 16 + t / 4 + 8 * (e2 >> 1)) * 1000 + 32 + 8 * k + 2 * (t % 4) + (e2 & 1);
    if (v != expect) { if (bad < 10) printf("w%d h%d t%d r%d: got %u expect %u\n", w, hf, t, c, v, expect); ++bad; }
  }
  printf("hypothesis mismatches: %d\n", bad);
  for (int t = 0; t < 6; ++t) { printf("w1 h0 t%d:", t); for (int c = 0; c < 16; ++c) printf(" %u", h[((1 * 2 + 0) * 32 + t) * 16 + c]); printf("\n"); }
  return 0;
}
