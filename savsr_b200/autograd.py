"""Backward of the tensor-core 3x3 convolution (stage A).

The reference trains through autograd with cuDNN's dgrad / wgrad (lbasicsr/models/sr_model.py:101-128).  ``conv3x3`` below is
the same differentiable op on libsavsr_sm100:

  forward   savsr_conv (tcgen05 implicit GEMM) on a scratch NHWC 16-bit arena
  dgrad     the SAME kernel run on dY with the transposed, spatially flipped filter (one launch, one group per 64-channel source)
  wgrad     savsr_conv_wgrad (tcgen05, pixels as the contraction dimension, K-major tiles straight from 16-bit NCHW tensors)
  dbias     a reduction over (n, y, x)

Weights may be shared ``[64, Ci, 3, 3]`` or per sample ``[B, 64, Ci, 3, 3]`` (the folded kernels of OSA-Conv,
savsr_arch.py:158-167).  Operands are rounded to the context's 16-bit format (bf16 by default), accumulation is fp32 -- the
gradients therefore match an fp32 autograd reference on the same rounded operands to fp32 accumulation-order noise.

PyTorch provides tensors, streams and the autograd tape only; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _capi as K
from . import engine


def _ctx_of(t: torch.Tensor) -> K.Context:
    if not t.is_cuda:
        raise RuntimeError("savsr_b200.autograd runs on CUDA (sm_100a) only; there is no CPU fallback")
    return engine.context(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _h16(ctx: K.Context) -> torch.dtype:
    return torch.float16 if ctx.lib.savsr_ctx_get_format(ctx.handle) == K.FMT_FP16 else torch.bfloat16


def _pack(ctx: K.Context, w: torch.Tensor, st: int) -> torch.Tensor:
    """fp32 [co, ci, 3, 3] (co multiple of 64) -> packed tensor-core blocks [co/64][ci/64 * 9][64][64], QUAD row order."""
    w = w.detach().float().contiguous()
    co, ci = w.shape[0], w.shape[1]
    out = torch.empty(ctx.lib.savsr_packed_weight_bytes(co, ci, 3), dtype=torch.uint8, device=w.device)
    K.check(ctx.lib.savsr_pack_conv_weight(w.data_ptr(), co, co, ci, 3, 64, ctx.lib.savsr_ctx_get_format(ctx.handle), K.ROWS_QUAD,
                                           out.data_ptr(), st))
    return out                    # (temporaries are safe to drop: the caching allocator reuses memory in stream order only)


def _group(src, dst, weight: torch.Tensor, bias: Optional[torch.Tensor], wstride: int) -> K.ConvGroup:
    g = K.ConvGroup()
    for i, s in enumerate(src):
        g.src_slot[i] = s
    g.nsrc = len(src)
    g.dst_slot, g.res1_slot, g.res2_slot, g.res2_scale = dst, -1, -1, 0.0
    g.act, g.slope = K.ACT_NONE, 0.0
    g.weight, g.weight_sample_stride = weight.data_ptr(), wstride
    g.bias = bias.data_ptr() if bias is not None else None
    g.mask = g.pool = g.aux_dst = None
    return g


def _nchw16(t: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """fp32 [B, C, H, W] -> 16-bit [B, C, H, pitch] with pitch = W rounded up to 8 and zero padding (savsr_conv_wgrad's layout)."""
    B, C, H, W = t.shape
    pitch = (W + 7) // 8 * 8
    if pitch == W:
        return t.to(dt).contiguous()
    out = torch.zeros(B, C, H, pitch, dtype=dt, device=t.device)
    out[..., :W] = t
    return out


def _nchw16_shifted(t: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """[B, C, H, W] (fp32 or 16-bit) -> 16-bit [3, B, C, H, pitch]: copy d holds X[.., x + d - 1] (zero outside the row): the three x-shifted
    views of the contraction operand of savsr_conv_wgrad (TMA boxes cannot start off a 16-byte granule)."""
    B, C, H, W = t.shape
    pitch = (W + 7) // 8 * 8
    out = torch.zeros(3, B, C, H, pitch, dtype=dt, device=t.device)
    t16 = t.to(dt)
    out[1, ..., :W] = t16
    n = min(W, pitch - 1)
    out[0, ..., 1:1 + n] = t16[..., :n]
    out[2, ..., :W - 1] = t16[..., 1:]
    return out


class _Conv3x3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        c = _ctx_of(x)
        B, Ci, H, W = x.shape
        per_sample = weight.dim() == 5
        if Ci % 64 or Ci > 64 * K.MAX_SRC or weight.shape[-3] != Ci or weight.shape[-4] != 64 or tuple(weight.shape[-2:]) != (3, 3):
            raise ValueError(f"conv3x3: unsupported shapes x {tuple(x.shape)}, weight {tuple(weight.shape)}")
        if per_sample and weight.shape[0] != B:
            raise ValueError("conv3x3: per-sample weights need one kernel per sample")
        nsrc = Ci // 64
        dt = _h16(c)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            arena_t = torch.empty((nsrc + 1) * B, H, W, 64, dtype=dt, device=x.device)
            arena = K.Arena(c, arena_t.data_ptr(), nsrc + 1, B, H, W)
            xs = x.detach().float()
            for s in range(nsrc):
                chunk = xs[:, 64 * s:64 * s + 64].contiguous()
                K.check(c.lib.savsr_arena_import(arena.handle, s, chunk.data_ptr(), st))
            wp = _pack(c, weight.reshape(-1, Ci, 3, 3), st)
            b32 = bias.detach().float().contiguous() if bias is not None else None
            g = _group(list(range(nsrc)), nsrc, wp, b32, 64 * Ci * 9 * 2 if per_sample else 0)
            arr = (K.ConvGroup * 1)(g)
            K.check(c.lib.savsr_conv(c.handle, arena.handle, arr, 1, 3, 64, K.DST_ARENA, K.IMPL_HALO, st))
            y = torch.empty(B, 64, H, W, dtype=torch.float32, device=x.device)
            K.check(c.lib.savsr_arena_export(arena.handle, nsrc, y.data_ptr(), st))
        ctx.save_for_backward(xs.to(dt), weight)       # 16-bit copy of the input: all the weight gradient needs
        ctx.has_bias = bias is not None
        ctx.width = W
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        x16, weight = ctx.saved_tensors
        c = _ctx_of(dy)
        B, Ci, H, W = x16.shape
        pitch = (W + 7) // 8 * 8
        per_sample = weight.dim() == 5
        nsrc = Ci // 64
        dt = x16.dtype
        dy = dy.detach().float().contiguous()
        dx = dw = db = None
        with torch.cuda.device(dy.device):
            st = torch.cuda.current_stream().cuda_stream
            if ctx.needs_input_grad[0]:
                # dX_s = conv3x3(dY; Wt_s),  Wt_s[i][o][ky][kx] = W[o][64 s + i][2-ky][2-kx]: one group per source slot, dY is the only source
                arena_t = torch.empty((nsrc + 1) * B, H, W, 64, dtype=dt, device=dy.device)
                arena = K.Arena(c, arena_t.data_ptr(), nsrc + 1, B, H, W)
                K.check(c.lib.savsr_arena_import(arena.handle, nsrc, dy.data_ptr(), st))
                w5 = weight.detach().float().reshape(-1, 64, Ci, 3, 3)                       # [1 or B, o, i, ky, kx]
                wt = w5.flip(-1, -2).permute(0, 2, 1, 3, 4)                                  # [1 or B, i, o, ky, kx]
                keep, groups = [], []
                for s in range(nsrc):
                    wp = _pack(c, wt[:, 64 * s:64 * s + 64].reshape(-1, 64, 3, 3), st)
                    keep.append(wp)
                    groups.append(_group([nsrc], s, wp, None, 64 * 64 * 9 * 2 if per_sample else 0))
                arr = (K.ConvGroup * nsrc)(*groups)
                K.check(c.lib.savsr_conv(c.handle, arena.handle, arr, nsrc, 3, 64, K.DST_ARENA, K.IMPL_HALO, st))
                dx = torch.empty(B, Ci, H, W, dtype=torch.float32, device=dy.device)
                for s in range(nsrc):
                    part = torch.empty(B, 64, H, W, dtype=torch.float32, device=dy.device)
                    K.check(c.lib.savsr_arena_export(arena.handle, s, part.data_ptr(), st))
                    dx[:, 64 * s:64 * s + 64] = part
            if ctx.needs_input_grad[1]:
                dy16 = _nchw16(dy, dt)
                x3 = _nchw16_shifted(x16, dt)                # the three x-shifted views, built only now (transient)
                dw = torch.zeros(weight.shape, dtype=torch.float32, device=dy.device)
                K.check(c.lib.savsr_conv_wgrad(c.handle, x3.data_ptr(), dy16.data_ptr(), B, Ci, H, W, pitch, 1 if per_sample else 0,
                                               dw.data_ptr(), st))
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = dy.sum(dim=(0, 2, 3))
        return dx, dw, db


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable 3x3 convolution, stride 1, zero padding 1, 64 output channels, on the tcgen05 kernels.
    x: fp32 [B, Ci, H, W] on a CUDA device with Ci in {64, 128, 192, 256, 320};
    weight: [64, Ci, 3, 3], or [B, 64, Ci, 3, 3] for per-sample kernels (OSA-Conv); bias: [64] or None (shared weights only)."""
    if weight.dim() == 5 and bias is not None:
        raise ValueError("conv3x3: per-sample kernels take no bias (OSA-Conv has none, savsr_arch.py:166)")
    return _Conv3x3.apply(x, weight, bias)


class _StaLrelu(torch.autograd.Function):
    """sta_conv of STAUpsample fused with kernel_conv's LeakyReLU (savsr_arch.py:297-313, 226-228) on libsavsr_sm100:
    out[b,c,p] = sum_t x[b,c,clamp(p + d_t)] * lrelu(kpre[b, c*25 + t, p]).  The 25-tap unfold of x and the activated per-pixel kernels are
    never materialised, in either direction."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, kpre: torch.Tensor, slope: float):
        c = _ctx_of(x)
        B, C, H, W = x.shape
        if kpre.shape != (B, C * 25, H, W):
            raise ValueError(f"sta_lrelu: kpre {tuple(kpre.shape)} does not match x {tuple(x.shape)} (5x5 kernels per channel)")
        x32, k32 = x.detach().float().contiguous(), kpre.detach().float().contiguous()
        out = torch.empty_like(x32)
        with torch.cuda.device(x.device):
            K.check(c.lib.savsr_sta_lrelu_forward(c.handle, x32.data_ptr(), k32.data_ptr(), out.data_ptr(), B, C, H, W, float(slope),
                                                  torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(x32, k32)
        ctx.slope = float(slope)
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        x32, k32 = ctx.saved_tensors
        c = _ctx_of(dout)
        B, C, H, W = x32.shape
        d32 = dout.detach().float().contiguous()
        dx, dk = torch.empty_like(x32), torch.empty_like(k32)
        with torch.cuda.device(dout.device):
            K.check(c.lib.savsr_sta_lrelu_backward(c.handle, x32.data_ptr(), k32.data_ptr(), d32.data_ptr(), dx.data_ptr(), dk.data_ptr(), B, C, H, W,
                                                   ctx.slope, torch.cuda.current_stream().cuda_stream))
        return dx, dk, None


def sta_lrelu(x: torch.Tensor, kpre: torch.Tensor, slope: float = 0.1) -> torch.Tensor:
    """Differentiable per-pixel 5x5 dynamic filtering with replicate padding; kpre = kernel_conv's output before its LeakyReLU."""
    return _StaLrelu.apply(x, kpre, slope)
