#!/usr/bin/env python
"""Per conv shape of the SAVSR trunk at the Vid4 size (17 windows): savsr_conv (tcgen05, 16-bit NHWC arena) vs the library
kernel the reference runs (F.conv2d through cuDNN, NCHW fp32 tensors, cudnn.benchmark = True, TF32 on and off; also
channels_last bf16 as the best case a PyTorch user could configure).  Prints a markdown table.

    python scripts/conv_vs_cudnn.py [B]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402
from gpu_checks import K  # noqa: E402


def t_ms(fn, warm=5, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    H, W = 144, 180
    torch.backends.cudnn.benchmark = True
    rows = []
    # (label, nsrc, groups in one of our launches, per-sample weights)
    shapes = [("64->64 (RCAB / conv0), 1 conv", 1, 1, False), ("64->64, 6 convs per launch (conv0 x 2 dirs x 3 streams)", 1, 6, False),
              ("128->64 (conv2), 6 convs per launch", 2, 6, False), ("192->64 OSA-Conv (grouped, per-sample weights), 2 convs", 3, 2, True),
              ("192->64 merge, 2 convs", 3, 2, False), ("320->64 OSA-Conv, 1 conv", 5, 1, True)]
    for label, nsrc, ng, per_sample in shapes:
        ci = 64 * nsrc
        flop = 2.0 * ng * B * H * W * 64 * ci * 9
        # ---- ours
        ab = G.ArenaBox(ng * (nsrc + 1), B, H, W)
        ab.t.normal_()
        groups, keep = [], []
        for g in range(ng):
            if per_sample:
                w = G.pack_weight(torch.randn(B * 64, ci, 3, 3, device=G.DEV) * 0.05)
                grp = G.group([g * (nsrc + 1) + i for i in range(nsrc)], g * (nsrc + 1) + nsrc, w, None, act=K.ACT_LRELU, wstride=64 * ci * 9 * 2)
            else:
                w = G.pack_weight(torch.randn(64, ci, 3, 3, device=G.DEV) * 0.05)
                bias = torch.randn(64, device=G.DEV)
                keep.append(bias)
                grp = G.group([g * (nsrc + 1) + i for i in range(nsrc)], g * (nsrc + 1) + nsrc, w, bias, act=K.ACT_LRELU)
            keep.append(w); groups.append(grp)
        arr = (K.ConvGroup * len(groups))(*groups)
        lib, ctx, st = K.load(), G.ctx().handle, G._stream()
        ours = t_ms(lambda: K.check(lib.savsr_conv(ctx, ab.a.handle, arr, len(groups), 3, 64, K.DST_ARENA, K.IMPL_HALO, st)))
        del ab
        # ---- cuDNN, the reference's call: conv2d (+ bias) then leaky_relu_, one call per conv; OSA = grouped conv with groups = B
        res = {}
        for tag, tf32, dt, cl in (("tf32", True, torch.float32, False), ("fp32", False, torch.float32, False), ("bf16_cl", True, torch.bfloat16, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            if per_sample:
                x = torch.randn(1, B * ci, H, W, device=G.DEV, dtype=dt)
                w = torch.randn(B * 64, ci, 3, 3, device=G.DEV, dtype=dt) * 0.05
                if cl:
                    x, w = x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
                fn = lambda: [F.leaky_relu_(F.conv2d(x, w, None, 1, 1, 1, groups=B), 0.2) for _ in range(ng)]   # noqa: E731
            else:
                x = torch.randn(B, ci, H, W, device=G.DEV, dtype=dt)
                w = torch.randn(64, ci, 3, 3, device=G.DEV, dtype=dt) * 0.05
                b = torch.randn(64, device=G.DEV, dtype=dt)
                if cl:
                    x, w = x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
                fn = lambda: [F.leaky_relu_(F.conv2d(x, w, b, 1, 1), 0.2) for _ in range(ng)]   # noqa: E731
            with torch.no_grad():
                try:
                    res[tag] = t_ms(fn, warm=3, reps=10)
                except Exception as e:  # noqa: BLE001
                    res[tag] = float("nan")
            del x, w
        rows.append((label, flop, ours, res))
    print(f"| conv shape (Vid4 144x180, {B} windows) | ours us | ours TFLOP/s | cuDNN TF32 us (x) | cuDNN fp32 us (x) | cuDNN bf16 channels_last us (x) |")
    print("|---|---|---|---|---|---|")
    for label, flop, ours, res in rows:
        f = lambda t: f"{t * 1e3:.0f} ({t / ours:.1f}x)"   # noqa: E731
        print(f"| {label} | {ours * 1e3:.0f} | {flop / ours / 1e9:.0f} | {f(res['tf32'])} | {f(res['fp32'])} | {f(res['bf16_cl'])} |")
    print("\n(x) = time relative to savsr_conv for the same convolutions; cuDNN numbers do not include the `torch.cat` that feeds multi-source convs "
          "in the reference, nor the x*ca / out*fa elementwise passes around OSA-Conv, which savsr_conv + the prologue absorb.")


if __name__ == "__main__":
    main()
