#!/usr/bin/env python
"""savsr_osadapt_mask at the Vid4 shape (B = 17, 144x180) for ncu / event timing (bring-up aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402
from gpu_checks import K  # noqa: E402

B, H, W = 17, 144, 180
dev = G.DEV
in16 = torch.relu(torch.randn(B, H * W, 16, device=dev))
wa = torch.randn(16, 16, 3, 3, device=dev) * 0.1; ba = torch.randn(16, device=dev) * 0.1
wb = torch.randn(16, 16, 3, 3, device=dev) * 0.1; bb = torch.randn(16, device=dev) * 0.1
wc = torch.randn(1, 16, 3, 3, device=dev) * 0.2; bc = torch.randn(1, device=dev) * 0.1
h0 = torch.empty(B, (H // 2) * (W // 2), 16, device=dev); h1 = torch.empty_like(h0)
mask = torch.empty(B, H * W, device=dev)


def run():
    K.check(K.load().savsr_osadapt_mask(G.ctx().handle, in16.data_ptr(), B, H, W, wa.data_ptr(), ba.data_ptr(), wb.data_ptr(), bb.data_ptr(),
                                        wc.data_ptr(), bc.data_ptr(), h0.data_ptr(), h1.data_ptr(), mask.data_ptr(), G._stream()))


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
print(f"osadapt_mask B={B} {H}x{W}: {e0.elapsed_time(e1) * 100:.1f} us per call (3 kernels)")
