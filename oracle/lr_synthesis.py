"""CPU restatement of the reference's LR-synthesis data path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(savsr_b200/datapath.py) runs CUDA kernels and has no CPU fallback.

What the reference does per output frame (lbasicsr/data/video_test_dataset.py:297-328):
  cv2.imread -> float32 / 255 (data_util.py:41) -> as_mod_crop (transforms.py:47-69) -> BGR->RGB, HWC->CHW
  (img_util.py img2tensor) -> T.Resize(size=(round(h/s_h), round(w/s_w)), BICUBIC, antialias=True) (data_util.py:396-412).
torchvision's Resize on a float tensor is torch.nn.functional.interpolate(mode="bicubic", antialias=True,
align_corners=False), i.e. ATen's `_upsample_bicubic2d_aa` (third-party dependency `torch`, un-pinned by the reference;
this image's 2.11.0 is the de-facto pin).  Its CPU kernel is restated here operation by operation:

  * separable: the WIDTH pass runs first into a float32 intermediate, then the HEIGHT pass; a dimension whose size does
    not change is skipped;
  * per output index i: scale = float(in) / float(out); support = 2 * scale (scale >= 1) else 2; center = scale * (i + 0.5)
    evaluated in double and rounded to float; xmin = max(int(center - support + 0.5), 0); xsize = min(int(center + support
    + 0.5), in) - xmin; tap j weighs cubic(|(j + xmin - center + 0.5) * invscale|) with the a = -0.5 (PIL) bicubic kernel,
    normalised by the float32 running total;
  * the cubic polynomials are compiled with fused multiply-adds: cubic1(x) = fma(fma(1.5, x, -2.5) * x, x, 1),
    cubic2(x) = fma(fma(fma(-0.5, x, 2.5), x, -4), x, 2);
  * the accumulation `t = src[0] * w[0]; for j in 1..n-1: t += src[j] * w[j]` is compiled as a 4-way unrolled
    multiply + add main loop with a fused-multiply-add remainder (taps j > 4 * ((n - 1) // 4)).
The last two points are properties of the x86-64 build of ATen 2.11.0, found by bisection against its outputs
(scripts/make_golden.py pins them in tests/golden/lr_kat.npz); any other association changes results by at most 1 ulp.
"""
from __future__ import annotations

from math import floor
from typing import List, Sequence, Tuple

import numpy as np

f32, f64 = np.float32, np.float64


# ------------------------------------------------------------------------------------------------ arbitrary-scale mod crop
def cal_step(scale: float) -> int:
    """transforms.py:31-44: the smallest of 1, 2, 5, 10, 20, 50 for which scale * step is (nearly) an integer."""
    for step in (1, 2, 5, 10, 20, 50):
        if abs(scale * step - round(scale * step)) < 0.001:
            return step
    raise ValueError(f"scale {scale} has no supported step (the reference raises UnboundLocalError here)")


def as_mod_crop_size(h: int, w: int, scale: Sequence[float]) -> Tuple[int, int]:
    """as_mod_crop (transforms.py:47-69): size of the top-left crop that keeps round(floor(h / step / s) * step * s) rows."""
    sh, sw = scale
    step_h, step_w = cal_step(sh), cal_step(sw)
    return round(floor(h / step_h / sh) * step_h * sh), round(floor(w / step_w / sw) * step_w * sw)


def lr_size(hc: int, wc: int, scale: Sequence[float]) -> Tuple[int, int]:
    """data_util.py:398: (round(h / scale_h), round(w / scale_w)) with Python's round (half to even)."""
    return round(hc / scale[0]), round(wc / scale[1])


# ------------------------------------------------------------------------------------------------ antialiased bicubic
def _fma(a, b, c):
    """fused multiply-add in float32: the float32 product is exact in float64, the sum is rounded once to float64 and
    once more to float32 (double rounding differs from a true fma only in ~2^-29 of the cases)."""
    return (np.asarray(a, dtype=f64) * np.asarray(b, dtype=f64) + np.asarray(c, dtype=f64)).astype(f32)


def _cubic(x: np.float32) -> np.float32:
    x = f32(x)
    if x < 1.0:
        t = _fma(f32(1.5), x, f32(-2.5))
        t = f32(t * x)
        return f32(_fma(t, x, f32(1.0)))
    if x < 2.0:
        t = _fma(f32(-0.5), x, f32(2.5))
        t = _fma(t, x, f32(-4.0))
        return f32(_fma(t, x, f32(2.0)))
    return f32(0.0)


def aa_bicubic_weights(in_size: int, out_size: int) -> List[Tuple[int, np.ndarray]]:
    """[(xmin, float32 weights[xsize])] per output index (ATen _compute_indices_min_size_weights_aa, bicubic a = -0.5)."""
    scale = f32(f32(in_size) / f32(out_size))
    support = f32(f32(2.0) * scale) if scale >= 1.0 else f32(2.0)
    invscale = f32(f64(1.0) / f64(scale)) if scale >= 1.0 else f32(1.0)
    out = []
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        xmin = max(int(f64(f32(center - support)) + 0.5), 0)
        xsize = min(int(f64(f32(center + support)) + 0.5), in_size) - xmin
        ws, total = [], f32(0.0)
        for j in range(xsize):
            x = f32((f64(f32(f32(j + xmin) - center)) + 0.5) * f64(invscale))
            w = _cubic(abs(x))
            ws.append(w)
            total = f32(total + w)
        if total != 0:
            ws = [f32(w / total) for w in ws]
        out.append((xmin, np.array(ws, dtype=f32)))
    return out


def _resize_last(x: np.ndarray, out_size: int) -> np.ndarray:
    if x.shape[-1] == out_size:
        return x
    res = np.empty(x.shape[:-1] + (out_size,), dtype=f32)
    for i, (xmin, ws) in enumerate(aa_bicubic_weights(x.shape[-1], out_size)):
        n = len(ws)
        main = ((n - 1) // 4) * 4
        t = x[..., xmin] * ws[0]
        for j in range(1, n):
            t = _fma(x[..., xmin + j], ws[j], t) if j > main else t + x[..., xmin + j] * ws[j]
        res[..., i] = t
    return res


def resize_aa_bicubic(x: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """x float32 [..., H, W] -> [..., oh, ow]: width pass, then height pass."""
    x = np.ascontiguousarray(x, dtype=f32)
    y = _resize_last(x, out_hw[1])
    y = np.swapaxes(_resize_last(np.ascontiguousarray(np.swapaxes(y, -1, -2)), out_hw[0]), -1, -2)
    return np.ascontiguousarray(y)


# ------------------------------------------------------------------------------------------------ frames -> LR / GT tensors
def frames_to_rgb(frames_bgr_u8: np.ndarray, crop_hw: Tuple[int, int]) -> np.ndarray:
    """uint8 [T,H,W,3] BGR (cv2.imread) -> float32 [T,3,hc,wc] RGB in [0,1] (data_util.py:41, img2tensor)."""
    hc, wc = crop_hw
    x = frames_bgr_u8[:, :hc, :wc, ::-1].astype(f32) / f32(255.0)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def synthesize_lr(frames_bgr_u8: np.ndarray, scale: Sequence[float]) -> Tuple[np.ndarray, np.ndarray]:
    """uint8 GT frames [T,H,W,3] BGR -> (LR float32 [T,3,h,w], mod-cropped GT float32 [T,3,Hc,Wc]).  The reference resizes
    the 7 frames of every window again for every output frame; frames are independent, so once per frame is identical."""
    T, H, W, _ = frames_bgr_u8.shape
    hc, wc = as_mod_crop_size(H, W, scale)
    gt = frames_to_rgb(frames_bgr_u8, (hc, wc))
    return resize_aa_bicubic(gt, lr_size(hc, wc, scale)), gt
