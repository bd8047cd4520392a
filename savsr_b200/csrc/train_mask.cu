// Train-mode OSAdapt mask net and combination (savsr_arch.py:186-214 and the `+ gamma * share` of 727-732), forward and backward,
// for the native training step (row f1).  Everything after the 64->16 convolution (which runs on the tensor-core kernel) is
// 16-channel or 1-channel work on fp32 channel-last maps -- [batch][pixels][16] -- far too small for tensor cores: plain
// CUDA-core kernels whose job is to replace ~120 framework launches per OSAdapt by ~35.
//
//   m0 = conv0(R) + b0 (savsr_conv, N = 16)  ->  BN1 (batch statistics) -> ReLU -> AvgPool2 = t2
//   m4 = conv4(t2) -> BN5 -> ReLU = t3 ;  m7 = conv7(t3) -> BN8 -> ReLU -> bilinear x2 = t5 ;  m11 = conv11(t5) -> BN12 -> sigmoid = mask
//   h  = R + a * mask + gamma * share                       (a = OSA-Conv(R), arena slots)
//
// BatchNorm in train mode: biased batch variance normalises, running statistics move by `momentum` with the unbiased variance.
// Sums for the statistics are accumulated in double (atomics), finalised by a one-block kernel that also zeroes the scratch.
#include "common.cuh"

namespace savsr {

static_assert(sizeof(savsr_mask_train) == 456, "C ABI struct layout changed: update savsr_b200/_capi.py (MaskTrain)");

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// ------------------------------------------------------------------------------------------------ BatchNorm statistics
// x: [rows][C], C = 16 or 1.  sums[0..C) += sum x, sums[C..2C) += sum x^2.  blockDim 256 (a multiple of C): a thread keeps its channel.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long n, int C, double* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s1[256], s2[256];
  float a = 0.f, b = 0.f;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < n; e += gridDim.x * 256L) { const float v = x[e]; a += v; b += v * v; }
  s1[threadIdx.x] = a; s2[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x < C) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = threadIdx.x; k < 256; k += C) { t1 += s1[k]; t2 += s2[k]; }
    atomicAdd(sums + threadIdx.x, t1);
    atomicAdd(sums + C + threadIdx.x, t2);
  }
}
// stat[c] = mean, stat[16 + c] = rstd; running statistics updated; sums zeroed.  One block of 32 threads.
__global__ void bn_finalize_kernel(double* sums, int C, double count, float eps, float momentum, float* rm, float* rv, float* stat) {
  pdl_wait();
  pdl_trigger();
  const int c = threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  stat[c] = static_cast<float>(mean);
  stat[16 + c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (rm) {
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    rm[c] = (1.f - momentum) * rm[c] + momentum * static_cast<float>(mean);
    rv[c] = (1.f - momentum) * rv[c] + momentum * static_cast<float>(unb);
  }
  sums[c] = 0.0; sums[C + c] = 0.0;
}

// ------------------------------------------------------------------------------------------------ forward pieces
// t2[b][q][c] = mean over the 2x2 block of ReLU(BN1(m0))
__global__ void __launch_bounds__(256) mask_norm_pool_kernel(const float* __restrict__ m0, const float* __restrict__ stat, const float* __restrict__ w,
                                                             const float* __restrict__ bb, float* __restrict__ t2, int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  const int h2 = H / 2, w2 = W / 2;
  const long total = static_cast<long>(B) * h2 * w2 * 16;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
    const int c = e & 15;
    const long q = e >> 4;
    const int x = q % w2, y = (q / w2) % h2, b = q / (static_cast<long>(w2) * h2);
    const float sc = stat[16 + c] * w[c], sh = bb[c] - stat[c] * sc;
    const float* src = m0 + ((static_cast<long>(b) * H + 2 * y) * W + 2 * x) * 16 + c;
    const float v = fmaxf(src[0] * sc + sh, 0.f) + fmaxf(src[16] * sc + sh, 0.f) + fmaxf(src[W * 16] * sc + sh, 0.f) + fmaxf(src[W * 16 + 16] * sc + sh, 0.f);
    t2[e] = 0.25f * v;
  }
}
// out = ReLU(BN(x)) elementwise, [rows][16]
__global__ void __launch_bounds__(256) mask_norm_relu_kernel(const float* __restrict__ x, const float* __restrict__ stat, const float* __restrict__ w,
                                                             const float* __restrict__ bb, float* __restrict__ out, long n) {
  pdl_wait();
  pdl_trigger();
  for (long e = blockIdx.x * 256L + threadIdx.x; e < n; e += gridDim.x * 256L) {
    const int c = e & 15;
    const float sc = stat[16 + c] * w[c];
    out[e] = fmaxf(x[e] * sc + (bb[c] - stat[c] * sc), 0.f);
  }
}
// t5 = bilinear x2 (align_corners = False) of ReLU(BN8(m7)); [B][h2*w2][16] -> [B][H*W][16]
__global__ void __launch_bounds__(256) mask_norm_up_kernel(const float* __restrict__ m7, const float* __restrict__ stat, const float* __restrict__ w,
                                                           const float* __restrict__ bb, float* __restrict__ t5, int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  const int h2 = H / 2, w2 = W / 2;
  const long total = static_cast<long>(B) * H * W * 16;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
    const int c = e & 15;
    const long p = e >> 4;
    const int x = p % W, y = (p / W) % H, b = p / (static_cast<long>(W) * H);
    int y0, y1, x0, x1; float ly, lx;
    bilinear_src(y, h2, H, y0, y1, ly);
    bilinear_src(x, w2, W, x0, x1, lx);
    const float sc = stat[16 + c] * w[c], sh = bb[c] - stat[c] * sc;
    const float* src = m7 + static_cast<long>(b) * h2 * w2 * 16 + c;
    const float v00 = fmaxf(src[(y0 * w2 + x0) * 16] * sc + sh, 0.f), v01 = fmaxf(src[(y0 * w2 + x1) * 16] * sc + sh, 0.f);
    const float v10 = fmaxf(src[(y1 * w2 + x0) * 16] * sc + sh, 0.f), v11 = fmaxf(src[(y1 * w2 + x1) * 16] * sc + sh, 0.f);
    t5[e] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}
// 3x3 convolution, 16 -> CO channels (CO = 16 or 1), zero padding, + bias.  in [B][hh*ww][16] -> out [B][hh*ww][CO].  Thread = pixel.
template <int CO>
__global__ void __launch_bounds__(128) mask_conv_kernel(const float* __restrict__ in, const float* __restrict__ wgt, const float* __restrict__ bias,
                                                        float* __restrict__ out, int B, int hh, int ww) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ws[9 * 16 * CO];          // [tap][c][o]
  for (int k = threadIdx.x; k < 9 * 16 * CO; k += blockDim.x) {
    const int o = k % CO, c = (k / CO) & 15, tap = k / (CO * 16);
    ws[k] = wgt[(o * 16 + c) * 9 + tap];
  }
  __syncthreads();
  const long total = static_cast<long>(B) * hh * ww;
  for (long q = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; q < total; q += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = q % ww, y = (q / ww) % hh;
    const long b = q / (static_cast<long>(ww) * hh);
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = bias[o];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= hh || xx < 0 || xx >= ww) continue;
      const float4* src = reinterpret_cast<const float4*>(in + ((b * hh + yy) * ww + xx) * 16);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 v = src[c4];
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* wr = ws + (tap * 16 + c4 * 4 + j) * CO;
#pragma unroll
          for (int o = 0; o < CO; ++o) acc[o] += vv[j] * wr[o];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) out[q * CO + o] = acc[o];
  }
}
// mask = sigmoid(BN12(m11)); h = R + a * mask + gamma * share on arena slots.  Thread = (pixel, 8-channel group).
__global__ void __launch_bounds__(256) mask_combine_kernel(const float* __restrict__ m11, const float* __restrict__ stat, const float* __restrict__ w,
                                                           const float* __restrict__ bb, const float* __restrict__ gamma, const uint16_t* __restrict__ R,
                                                           const uint16_t* __restrict__ A, const uint16_t* __restrict__ S, uint16_t* __restrict__ Hh,
                                                           float* __restrict__ mask, long npix_total, int fmt) {
  pdl_wait();
  pdl_trigger();
  const float sc = stat[16] * w[0], sh = bb[0] - stat[0] * sc, gm = gamma[0];
  for (long e = blockIdx.x * 256L + threadIdx.x; e < npix_total * 8; e += gridDim.x * 256L) {
    const long p = e >> 3;
    const float mk = sigm(m11[p] * sc + sh);
    if ((e & 7) == 0) mask[p] = mk;
    const uint4 r = reinterpret_cast<const uint4*>(R)[e], a = reinterpret_cast<const uint4*>(A)[e], s = reinterpret_cast<const uint4*>(S)[e];
    const uint32_t rw[4] = {r.x, r.y, r.z, r.w}, aw[4] = {a.x, a.y, a.z, a.w}, sw[4] = {s.x, s.y, s.z, s.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = pack_h2(h_lo(rw[j], fmt) + h_lo(aw[j], fmt) * mk + gm * h_lo(sw[j], fmt), h_hi(rw[j], fmt) + h_hi(aw[j], fmt) * mk + gm * h_hi(sw[j], fmt), fmt);
    reinterpret_cast<uint4*>(Hh)[e] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------ backward pieces
// da = dh * mask, gshare = gamma * dh (slots), dmask[p] = sum_c dh * a, d gamma += sum dh * share.
__global__ void __launch_bounds__(256) mask_combine_bwd_kernel(const uint16_t* __restrict__ dH, const uint16_t* __restrict__ A, const uint16_t* __restrict__ S,
                                                               const float* __restrict__ mask, const float* __restrict__ gamma, uint16_t* __restrict__ dA,
                                                               uint16_t* __restrict__ gS, float* __restrict__ dmask, float* __restrict__ d_gamma,
                                                               long npix_total, int fmt) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8];
  const float gm = gamma[0];
  float dg = 0.f;
  for (long e0 = blockIdx.x * 256L; e0 < npix_total * 8; e0 += gridDim.x * 256L) {      // block-uniform trip count: the shuffles below need full warps
    const long e = e0 + threadIdx.x;
    const bool ok = e < npix_total * 8;
    float dm = 0.f;
    if (ok) {
      const long p = e >> 3;
      const float mk = mask[p];
      const uint4 d = reinterpret_cast<const uint4*>(dH)[e], a = reinterpret_cast<const uint4*>(A)[e], s = reinterpret_cast<const uint4*>(S)[e];
      const uint32_t dw[4] = {d.x, d.y, d.z, d.w}, aw[4] = {a.x, a.y, a.z, a.w}, sw[4] = {s.x, s.y, s.z, s.w};
      uint32_t o1[4], o2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d0 = h_lo(dw[j], fmt), d1 = h_hi(dw[j], fmt);
        dm += d0 * h_lo(aw[j], fmt) + d1 * h_hi(aw[j], fmt);
        dg += d0 * h_lo(sw[j], fmt) + d1 * h_hi(sw[j], fmt);
        o1[j] = pack_h2(d0 * mk, d1 * mk, fmt);
        o2[j] = pack_h2(d0 * gm, d1 * gm, fmt);
      }
      reinterpret_cast<uint4*>(dA)[e] = make_uint4(o1[0], o1[1], o1[2], o1[3]);
      reinterpret_cast<uint4*>(gS)[e] = make_uint4(o2[0], o2[1], o2[2], o2[3]);
    }
    dm += __shfl_xor_sync(0xffffffffu, dm, 1);
    dm += __shfl_xor_sync(0xffffffffu, dm, 2);
    dm += __shfl_xor_sync(0xffffffffu, dm, 4);
    if (ok && (threadIdx.x & 7) == 0) dmask[e >> 3] = dm;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dg;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(d_gamma, t);
  }
}

// BatchNorm backward, pass 1.  dy' = dy * act'(BN(x)) (act 1 = ReLU, 2 = sigmoid); pooled: dy comes from the half-resolution map
// through AvgPool2 (dy = 0.25 * dsrc[b][y/2][x/2]).  sums[c] += dy', sums[C + c] += dy' * xhat.
struct BnBwd {
  const float* dy; const float* x; const float* stat; const float* w; const float* b;
  int C, act, pooled, H, W;      // H, W: full resolution (pooled only)
  long n;                        // elements of x
};
__device__ __forceinline__ float bn_bwd_dyp(const BnBwd& p, long e, int c, float& xhat) {
  float dy;
  if (p.pooled) {
    const long pix = e >> 4;
    const int x = pix % p.W, y = (pix / p.W) % p.H;
    const long b = pix / (static_cast<long>(p.W) * p.H);
    dy = 0.25f * p.dy[((b * (p.H / 2) + y / 2) * (p.W / 2) + x / 2) * 16 + c];
  } else {
    dy = p.dy[e];
  }
  xhat = (p.x[e] - p.stat[c]) * p.stat[16 + c];
  const float yv = xhat * p.w[c] + p.b[c];
  if (p.act == 1) return yv > 0.f ? dy : 0.f;
  if (p.act == 2) { const float s = sigm(yv); return dy * s * (1.f - s); }
  return dy;
}
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwd p, double* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s1[256], s2[256];
  float a = 0.f, b = 0.f;
  const int c = threadIdx.x % p.C;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < p.n; e += gridDim.x * 256L) {
    float xh;
    const float d = bn_bwd_dyp(p, e, c, xh);
    a += d; b += d * xh;
  }
  s1[threadIdx.x] = a; s2[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x < p.C) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = threadIdx.x; k < 256; k += p.C) { t1 += s1[k]; t2 += s2[k]; }
    atomicAdd(sums + threadIdx.x, t1);
    atomicAdd(sums + p.C + threadIdx.x, t2);
  }
}
// d weight += s2, d bias += s1; coef = [w * rstd | s1 / count | s2 / count] (16 each); sums zeroed.
__global__ void bn_bwd_finalize_kernel(double* sums, int C, double count, const float* stat, const float* w, float* dw, float* db, float* coef) {
  pdl_wait();
  pdl_trigger();
  const int c = threadIdx.x;
  if (c >= C) return;
  const double s1 = sums[c], s2 = sums[C + c];
  dw[c] += static_cast<float>(s2);
  db[c] += static_cast<float>(s1);
  coef[c] = w[c] * stat[16 + c];
  coef[16 + c] = static_cast<float>(s1 / count);
  coef[32 + c] = static_cast<float>(s2 / count);
  sums[c] = 0.0; sums[C + c] = 0.0;
}
// pass 2: dx = coef0 * (dy' - coef1 - xhat * coef2) -> fp32 [rows][C], or (slot != NULL) the first 16 channels of an arena slot
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwd p, const float* __restrict__ coef, float* __restrict__ dx, uint16_t* __restrict__ slot, int fmt) {
  pdl_wait();
  pdl_trigger();
  const int c = threadIdx.x % p.C;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < p.n; e += gridDim.x * 256L) {
    float xh;
    const float d = bn_bwd_dyp(p, e, c, xh);
    const float v = coef[c] * (d - coef[16 + c] - xh * coef[32 + c]);
    if (slot) slot[(e >> 4) * kC + c] = float_to_h(v, fmt);
    else dx[e] = v;
  }
}

// weight / bias gradient of a 3x3 16 -> CO convolution: dW[o][c][tap] += sum_q dout[q][o] in[q + tap][c].  Block = a chunk of pixels,
// thread = (o, c) for CO = 16, (c, tap) for CO = 1.
template <int CO>
__global__ void __launch_bounds__(256) mask_conv_wgrad_kernel(const float* __restrict__ dout, const float* __restrict__ in, float* __restrict__ dW,
                                                              float* __restrict__ dbias, int B, int hh, int ww, int chunk) {
  pdl_wait();
  pdl_trigger();
  const long total = static_cast<long>(B) * hh * ww;
  const long q0 = blockIdx.x * static_cast<long>(chunk), q1 = min(q0 + chunk, total);
  if (CO == 16) {
    const int o = threadIdx.x >> 4, c = threadIdx.x & 15;
    float acc[9], ab = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = 0.f;
    for (long q = q0; q < q1; ++q) {
      const int x = q % ww, y = (q / ww) % hh;
      const float d = dout[q * 16 + o];
      ab += d;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < hh && xx >= 0 && xx < ww) acc[tap] += d * in[(q + (tap / 3 - 1) * ww + (tap % 3 - 1)) * 16 + c];
      }
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) atomicAdd(dW + (o * 16 + c) * 9 + tap, acc[tap]);
    if (c == 0) atomicAdd(dbias + o, ab);
  } else {
    const int c = threadIdx.x & 15, tap = threadIdx.x >> 4;      // 144 active threads + thread 255 for the bias
    float acc = 0.f;
    if (tap < 9) {
      for (long q = q0; q < q1; ++q) {
        const int x = q % ww, y = (q / ww) % hh;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < hh && xx >= 0 && xx < ww) acc += dout[q] * in[(q + (tap / 3 - 1) * ww + (tap % 3 - 1)) * 16 + c];
      }
      atomicAdd(dW + c * 9 + tap, acc);
    } else if (threadIdx.x == 255) {
      for (long q = q0; q < q1; ++q) acc += dout[q];
      atomicAdd(dbias, acc);
    }
  }
}
// data gradient of the 16 -> 16 convolution: din[q][c] = sum_tap sum_o W[o][c][tap] dout[q - tap][o].  Thread = pixel.
__global__ void __launch_bounds__(128) mask_conv16_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ wgt, float* __restrict__ din,
                                                                int B, int hh, int ww) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ws[9 * 16 * 16];          // [tap][o][c]
  for (int k = threadIdx.x; k < 9 * 256; k += blockDim.x) {
    const int c = k & 15, o = (k >> 4) & 15, tap = k >> 8;
    ws[k] = wgt[(o * 16 + c) * 9 + tap];
  }
  __syncthreads();
  const long total = static_cast<long>(B) * hh * ww;
  for (long q = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; q < total; q += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = q % ww, y = (q / ww) % hh;
    const long b = q / (static_cast<long>(ww) * hh);
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);        // the output pixel that read this input through `tap`
      if (yy < 0 || yy >= hh || xx < 0 || xx >= ww) continue;
      const float4* src = reinterpret_cast<const float4*>(dout + ((b * hh + yy) * ww + xx) * 16);
#pragma unroll
      for (int o4 = 0; o4 < 4; ++o4) {
        const float4 v = src[o4];
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float* wr = ws + (tap * 16 + o4 * 4 + j) * 16;
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c] += vv[j] * wr[c];
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) din[q * 16 + c] = acc[c];
  }
}
// data gradient of the 16 -> 1 convolution followed by the transpose of the bilinear x2 upsample: dt4 (half resolution, zeroed) +=.
// Thread = full-resolution pixel p': dt5[p'][c] = sum_tap W11[c][tap] dm11[p' - tap], scattered to its four source pixels.
__global__ void __launch_bounds__(128) mask_conv1_dgrad_up_kernel(const float* __restrict__ dm11, const float* __restrict__ w11, float* __restrict__ dt4,
                                                                  int B, int H, int W) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ws[9 * 16];               // [tap][c]
  for (int k = threadIdx.x; k < 144; k += blockDim.x) ws[k] = w11[(k & 15) * 9 + (k >> 4)];
  __syncthreads();
  const int h2 = H / 2, w2 = W / 2;
  const long total = static_cast<long>(B) * H * W;
  for (long p = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; p < total; p += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = p % W, y = (p / W) % H;
    const long b = p / (static_cast<long>(W) * H);
    float g[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) g[c] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const float d = dm11[(b * H + yy) * W + xx];
#pragma unroll
      for (int c = 0; c < 16; ++c) g[c] += d * ws[tap * 16 + c];
    }
    int y0, y1, x0, x1; float ly, lx;
    bilinear_src(y, h2, H, y0, y1, ly);
    bilinear_src(x, w2, W, x0, x1, lx);
    float* base = dt4 + b * h2 * w2 * 16;
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11v = ly * lx;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      atomicAdd(base + (y0 * w2 + x0) * 16 + c, w00 * g[c]);
      if (w01 != 0.f) atomicAdd(base + (y0 * w2 + x1) * 16 + c, w01 * g[c]);
      if (w10 != 0.f) atomicAdd(base + (y1 * w2 + x0) * 16 + c, w10 * g[c]);
      if (w11v != 0.f) atomicAdd(base + (y1 * w2 + x1) * 16 + c, w11v * g[c]);
    }
  }
}

static int blocks_for(const savsr_ctx* ctx, long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  const long cap = 8L * ctx->sm_count;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace savsr

using namespace savsr;

static int check_mask(const char* who, const savsr_ctx* ctx, const savsr_arena* arena, const savsr_mask_train* m) {
  SAVSR_REQUIRE(ctx && arena && m, "%s: null pointer", who);
  SAVSR_REQUIRE(arena->height % 2 == 0 && arena->width % 2 == 0, "%s: OSAdapt needs even sizes, got %dx%d", who, arena->height, arena->width);
  SAVSR_REQUIRE(m->w4 && m->b4 && m->w7 && m->b7 && m->w11 && m->b11 && m->gamma && m->m0 && m->t2 && m->m4 && m->t3 && m->m7 && m->t5 && m->m11 && m->mask &&
                m->stat && m->sums, "%s: null parameter / activation pointer", who);
  for (int l = 0; l < 4; ++l) SAVSR_REQUIRE(m->bn_w[l] && m->bn_b[l], "%s: BatchNorm %d has no weight / bias", who, l);
  return 0;
}

extern "C" int savsr_mask_forward_train(savsr_ctx* ctx, savsr_arena* arena, const savsr_mask_train* m, int r_slot, int a_slot, int share_slot, int out_slot,
                                        savsr_stream st_) {
  if (int rc = check_mask("savsr_mask_forward_train", ctx, arena, m)) return rc;
  SAVSR_REQUIRE(r_slot >= 0 && r_slot < arena->nslots && a_slot >= 0 && a_slot < arena->nslots && share_slot >= 0 && share_slot < arena->nslots &&
                out_slot >= 0 && out_slot < arena->nslots, "savsr_mask_forward_train: slot out of range");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  const int B = arena->batch, H = arena->height, W = arena->width, h2 = H / 2, w2 = W / 2;
  const long P = static_cast<long>(B) * H * W, Q = static_cast<long>(B) * h2 * w2;
  auto stats = [&](const float* x, long rows, int C, int layer) {
    (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, bn_stats_kernel, dim3(blocks_for(ctx, rows * C, 256 * 8)), dim3(256), 0, st, x, rows * C, C, m->sums);
    (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, bn_finalize_kernel, dim3(1), dim3(32), 0, st, m->sums, C, static_cast<double>(rows), m->eps, m->momentum, m->bn_rm[layer], m->bn_rv[layer], m->stat + layer * 32);
  };
  stats(m->m0, P, 16, 0);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_norm_pool_kernel, dim3(blocks_for(ctx, Q * 16, 256)), dim3(256), 0, st, m->m0, m->stat, m->bn_w[0], m->bn_b[0], m->t2, B, H, W);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv_kernel<16>, dim3(blocks_for(ctx, Q, 128)), dim3(128), 0, st, m->t2, m->w4, m->b4, m->m4, B, h2, w2);
  stats(m->m4, Q, 16, 1);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_norm_relu_kernel, dim3(blocks_for(ctx, Q * 16, 256)), dim3(256), 0, st, m->m4, m->stat + 32, m->bn_w[1], m->bn_b[1], m->t3, Q * 16);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv_kernel<16>, dim3(blocks_for(ctx, Q, 128)), dim3(128), 0, st, m->t3, m->w7, m->b7, m->m7, B, h2, w2);
  stats(m->m7, Q, 16, 2);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_norm_up_kernel, dim3(blocks_for(ctx, P * 16, 256)), dim3(256), 0, st, m->m7, m->stat + 64, m->bn_w[2], m->bn_b[2], m->t5, B, H, W);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv_kernel<1>, dim3(blocks_for(ctx, P, 128)), dim3(128), 0, st, m->t5, m->w11, m->b11, m->m11, B, H, W);
  stats(m->m11, P, 1, 3);
  const long img = static_cast<long>(H) * W * kC * B;
  const uint16_t* base = reinterpret_cast<const uint16_t*>(arena->base);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_combine_kernel, dim3(blocks_for(ctx, P * 8, 256)), dim3(256), 0, st, m->m11, m->stat + 96, m->bn_w[3], m->bn_b[3], m->gamma, base + r_slot * img, base + a_slot * img,
                                                                 base + share_slot * img, reinterpret_cast<uint16_t*>(arena->base) + out_slot * img, m->mask, P, ctx->fmt);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_mask_backward_train(savsr_ctx* ctx, savsr_arena* arena, const savsr_mask_train* m, int dh_slot, int a_slot, int share_slot, int da_slot,
                                         int gshare_slot, int dm0_slot, savsr_stream st_) {
  if (int rc = check_mask("savsr_mask_backward_train", ctx, arena, m)) return rc;
  SAVSR_REQUIRE(m->d_w4 && m->d_b4 && m->d_w7 && m->d_b7 && m->d_w11 && m->d_b11 && m->d_gamma && m->dmask && m->dm11 && m->dt4 && m->dm7 && m->dt3 && m->dm4 &&
                m->dt2 && m->coef, "savsr_mask_backward_train: null gradient pointer");
  for (int l = 0; l < 4; ++l) SAVSR_REQUIRE(m->d_bn_w[l] && m->d_bn_b[l], "savsr_mask_backward_train: BatchNorm %d has no gradient buffers", l);
  const int slots[6] = {dh_slot, a_slot, share_slot, da_slot, gshare_slot, dm0_slot};
  for (int s : slots) SAVSR_REQUIRE(s >= 0 && s < arena->nslots, "savsr_mask_backward_train: slot %d out of range", s);
  DeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  const int B = arena->batch, H = arena->height, W = arena->width, h2 = H / 2, w2 = W / 2;
  const long P = static_cast<long>(B) * H * W, Q = static_cast<long>(B) * h2 * w2;
  const long img = static_cast<long>(H) * W * kC * B;
  uint16_t* base = reinterpret_cast<uint16_t*>(arena->base);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_combine_bwd_kernel, dim3(blocks_for(ctx, P * 8, 256)), dim3(256), 0, st, base + dh_slot * img, base + a_slot * img, base + share_slot * img, m->mask, m->gamma,
                                                                     base + da_slot * img, base + gshare_slot * img, m->dmask, m->d_gamma, P, ctx->fmt);
  auto bn_bwd = [&](const float* dy, const float* x, int layer, int C, int act, int pooled, long n, double count, float* dx, uint16_t* slot) {
    BnBwd p;
    p.dy = dy; p.x = x; p.stat = m->stat + layer * 32; p.w = m->bn_w[layer]; p.b = m->bn_b[layer];
    p.C = C; p.act = act; p.pooled = pooled; p.H = H; p.W = W; p.n = n;
    (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, bn_bwd_reduce_kernel, dim3(blocks_for(ctx, n, 256 * 8)), dim3(256), 0, st, p, m->sums);
    (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, bn_bwd_finalize_kernel, dim3(1), dim3(32), 0, st, m->sums, C, count, p.stat, p.w, m->d_bn_w[layer], m->d_bn_b[layer], m->coef + layer * 48);
    (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, bn_bwd_apply_kernel, dim3(blocks_for(ctx, n, 256 * 4)), dim3(256), 0, st, p, m->coef + layer * 48, dx, slot, ctx->fmt);
  };
  // sigmoid + BN12 -> dm11
  bn_bwd(m->dmask, m->m11, 3, 1, 2, 0, P, static_cast<double>(P), m->dm11, nullptr);
  // conv11 (16 -> 1) on the upsampled map
  const int chunk1 = 32;       // short pixel chunks: the per-thread loop is a chain of dependent loads, parallelism comes from the block count
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv_wgrad_kernel<1>, dim3(static_cast<int>((P + chunk1 - 1) / chunk1)), dim3(256), 0, st, m->dm11, m->t5, m->d_w11, m->d_b11, B, H, W, chunk1);
  SAVSR_CUDA(cudaMemsetAsync(m->dt4, 0, static_cast<size_t>(Q) * 16 * sizeof(float), st));
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv1_dgrad_up_kernel, dim3(blocks_for(ctx, P, 128)), dim3(128), 0, st, m->dm11, m->w11, m->dt4, B, H, W);
  // ReLU + BN8 -> dm7 ; conv7
  bn_bwd(m->dt4, m->m7, 2, 16, 1, 0, Q * 16, static_cast<double>(Q), m->dm7, nullptr);
  const int chunk16 = 16;
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv_wgrad_kernel<16>, dim3(static_cast<int>((Q + chunk16 - 1) / chunk16)), dim3(256), 0, st, m->dm7, m->t3, m->d_w7, m->d_b7, B, h2, w2, chunk16);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv16_dgrad_kernel, dim3(blocks_for(ctx, Q, 128)), dim3(128), 0, st, m->dm7, m->w7, m->dt3, B, h2, w2);
  // ReLU + BN5 -> dm4 ; conv4
  bn_bwd(m->dt3, m->m4, 1, 16, 1, 0, Q * 16, static_cast<double>(Q), m->dm4, nullptr);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv_wgrad_kernel<16>, dim3(static_cast<int>((Q + chunk16 - 1) / chunk16)), dim3(256), 0, st, m->dm4, m->t2, m->d_w4, m->d_b4, B, h2, w2, chunk16);
  (void)launch_k(ctx->opt[SAVSR_OPT_PDL] != 0, mask_conv16_dgrad_kernel, dim3(blocks_for(ctx, Q, 128)), dim3(128), 0, st, m->dm4, m->w4, m->dt2, B, h2, w2);
  // AvgPool2 + ReLU + BN1 -> dm0, written as the first 16 channels of an arena slot (the other 48 stay zero): operand of the tensor-core
  // data / weight gradient of the 64 -> 16 convolution
  bn_bwd(m->dt2, m->m0, 0, 16, 1, 1, P * 16, static_cast<double>(P), nullptr, base + dm0_slot * img);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ SATU: per-pixel 5x5 dynamic filter
// sta_conv of STAUpsample (savsr_arch.py:297-313) fused with the LeakyReLU of kernel_conv (226-228), for the training step:
//   out[b,c,p] = sum_t x[b,c,clamp(p + d_t)] * lrelu(kpre[b, c*25 + t, p])            (replicate padding, t = 5 u + v, d_t = (u-2, v-2))
// fp32 NCHW tensors (the island around it is ATen).  Without this fusion the framework materialises the 25-tap unfold of x, the
// activated kernels and two more products of that size (105 MB each at 4 x 64 x 64) in both directions.
namespace savsr {
__global__ void __launch_bounds__(256) sta_lrelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ kpre, float* __restrict__ out,
                                                            long planes, int h, int w, float slope) {
  const long npix = static_cast<long>(h) * w, total = planes * npix;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
    const long pl = e / npix, p = e - pl * npix;
    const int yy = p / w, xx = p - static_cast<long>(yy) * w;
    const float* xs = x + pl * npix;
    const float* ks = kpre + pl * 25 * npix + p;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 25; ++t) {
      const int sy = min(max(yy + t / 5 - 2, 0), h - 1), sx = min(max(xx + t % 5 - 2, 0), w - 1);
      const float k = ks[t * npix];
      acc += xs[sy * w + sx] * (k > 0.f ? k : slope * k);
    }
    out[e] = acc;
  }
}
// d kpre[b, c*25 + t, p] = dout[b,c,p] * x[b,c,clamp(p + d_t)] * lrelu'(kpre)
__global__ void __launch_bounds__(256) sta_lrelu_bwd_k_kernel(const float* __restrict__ x, const float* __restrict__ kpre, const float* __restrict__ dout,
                                                              float* __restrict__ dkpre, long planes, int h, int w, float slope) {
  const long npix = static_cast<long>(h) * w, total = planes * npix;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
    const long pl = e / npix, p = e - pl * npix;
    const int yy = p / w, xx = p - static_cast<long>(yy) * w;
    const float* xs = x + pl * npix;
    const long kb = pl * 25 * npix + p;
    const float g = dout[e];
#pragma unroll
    for (int t = 0; t < 25; ++t) {
      const int sy = min(max(yy + t / 5 - 2, 0), h - 1), sx = min(max(xx + t % 5 - 2, 0), w - 1);
      dkpre[kb + t * npix] = g * xs[sy * w + sx] * (kpre[kb + t * npix] > 0.f ? 1.f : slope);
    }
  }
}
// d x[b,c,p'] = sum over (p, t) with clamp(p + d_t) = p' of dout[b,c,p] * lrelu(kpre[b, c*25 + t, p]).  Gather form: for an interior
// coordinate the source is unique (p = p' - d_t); on a border every p whose shifted coordinate was clamped onto it contributes.
__device__ __forceinline__ void sta_sources(int target, int d, int size, int& lo, int& hi) {
  // all s in [0, size) with clamp(s + d, 0, size - 1) == target
  if (target > 0 && target < size - 1) { lo = hi = target - d; if (lo < 0 || lo >= size) { lo = 1; hi = 0; } return; }
  if (target == 0) { lo = 0; hi = min(size - 1, -d); if (size == 1) hi = 0; return; }          // s + d <= 0
  lo = max(0, size - 1 - d); hi = size - 1;                                                     // s + d >= size - 1
}
__global__ void __launch_bounds__(256) sta_lrelu_bwd_x_kernel(const float* __restrict__ kpre, const float* __restrict__ dout, float* __restrict__ dx,
                                                              long planes, int h, int w, float slope) {
  const long npix = static_cast<long>(h) * w, total = planes * npix;
  for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += gridDim.x * 256L) {
    const long pl = e / npix, p = e - pl * npix;
    const int ty = p / w, tx = p - static_cast<long>(ty) * w;
    const float* ds = dout + pl * npix;
    const float* ks = kpre + pl * 25 * npix;
    float acc = 0.f;
    for (int t = 0; t < 25; ++t) {
      int y0, y1, x0, x1;
      sta_sources(ty, t / 5 - 2, h, y0, y1);
      sta_sources(tx, t % 5 - 2, w, x0, x1);
      for (int sy = y0; sy <= y1; ++sy)
        for (int sx = x0; sx <= x1; ++sx) {
          const float k = ks[t * npix + sy * w + sx];
          acc += ds[sy * w + sx] * (k > 0.f ? k : slope * k);
        }
    }
    dx[e] = acc;
  }
}
}  // namespace savsr

extern "C" int savsr_sta_lrelu_forward(savsr_ctx* ctx, const float* x, const float* kpre, float* out, int batch, int channels, int height, int width,
                                       float slope, savsr_stream st) {
  SAVSR_REQUIRE(ctx && x && kpre && out, "savsr_sta_lrelu_forward: null pointer");
  SAVSR_REQUIRE(batch >= 1 && channels >= 1 && height >= 1 && width >= 1, "savsr_sta_lrelu_forward: empty problem");
  DeviceGuard guard(ctx->device);
  const long planes = static_cast<long>(batch) * channels;
  sta_lrelu_fwd_kernel<<<blocks_for(ctx, planes * height * width, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(x, kpre, out, planes, height, width, slope);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int savsr_sta_lrelu_backward(savsr_ctx* ctx, const float* x, const float* kpre, const float* dout, float* dx, float* dkpre, int batch, int channels,
                                        int height, int width, float slope, savsr_stream st) {
  SAVSR_REQUIRE(ctx && x && kpre && dout && dx && dkpre, "savsr_sta_lrelu_backward: null pointer");
  SAVSR_REQUIRE(batch >= 1 && channels >= 1 && height >= 1 && width >= 1, "savsr_sta_lrelu_backward: empty problem");
  DeviceGuard guard(ctx->device);
  const long planes = static_cast<long>(batch) * channels;
  const int blocks = blocks_for(ctx, planes * height * width, 256);
  sta_lrelu_bwd_k_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(st)>>>(x, kpre, dout, dkpre, planes, height, width, slope);
  sta_lrelu_bwd_x_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(st)>>>(kpre, dout, dx, planes, height, width, slope);
  SAVSR_CUDA(cudaGetLastError());
  return 0;
}
