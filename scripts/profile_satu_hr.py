#!/usr/bin/env python
"""Launch savsr_satu_hr alone at a BASELINE shape (for ncu / timing):  python scripts/profile_satu_hr.py [B] [h] [w] [s_h] [s_w] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402
from savsr_b200 import _capi as K, engine  # noqa: E402
from oracle.state_dict_fixture import make_state_dict  # noqa: E402


def main():
    a = sys.argv[1:]
    B, h, w = (int(a[0]) if a else 17), (int(a[1]) if len(a) > 1 else 144), (int(a[2]) if len(a) > 2 else 180)
    scale = (float(a[3]) if len(a) > 3 else 4.0, float(a[4]) if len(a) > 4 else 4.0)
    reps = int(a[5]) if len(a) > 5 else 5
    dev = G.DEV
    sd = make_state_dict(0)
    res, H, W = G.satu_index(h, w, scale, sd)
    hp, wp = h + (h & 1), w + (w & 1)
    lr = G.ArenaBox(2, B, hp, wp)
    lr.t.copy_(torch.randn_like(lr.t, dtype=torch.float32).to(lr.t.dtype))
    xin = torch.rand(B, 7, 3, h, w, device=dev)
    out = torch.empty(B, 3, H, W, device=dev)
    table = res["table"].to(dev); by = torch.from_numpy(res["base_y"]).to(dev); bx = torch.from_numpy(res["base_x"]).to(dev)
    parts = engine.satu_hr_compose(sd, dev)
    wts = engine.satu_hr_pack(parts, K.FMT_BF16, dev)
    zb = parts[4].contiguous(); tb = sd["tail.bias"].to(dev).contiguous()
    lib, ctx = K.load(), G.ctx()
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: K.check(lib.savsr_satu_hr(ctx.handle, lr.a.handle, 0, 1, h, w, H, W, table.data_ptr(), by.data_ptr(), bx.data_ptr(),  # noqa: E731
                                            wts.data_ptr(), zb.data_ptr(), tb.data_ptr(), xin.data_ptr(), 7, 3, out.data_ptr(), st))
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"satu_hr B={B} {h}x{w} x{scale} -> {H}x{W}: {ms * 1e3:.1f} us per launch, {ms * 1e3 / B:.1f} us per frame")


if __name__ == "__main__":
    main()
