#!/usr/bin/env python
"""One batched conv launch at the Vid4 shape for ncu / timing (bring-up aid).
usage: profile_conv.py [nsrc] [ngroups] [batch] [impl] [ksize] [reps] [pool 0/1] [res1 0/1]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402
from gpu_checks import K  # noqa: E402


def main():
    a = sys.argv[1:]
    nsrc = int(a[0]) if len(a) > 0 else 1
    ng = int(a[1]) if len(a) > 1 else 6
    B = int(a[2]) if len(a) > 2 else 4
    impl = K.IMPL_NAMES[a[3]] if len(a) > 3 else K.IMPL_HALO
    ks = int(a[4]) if len(a) > 4 else 3
    reps = int(a[5]) if len(a) > 5 else 5
    pool = len(a) > 6 and a[6] == "1"
    res = len(a) > 7 and a[7] == "1"
    H, W = 144, 180
    ab = G.ArenaBox(ng * (nsrc + 1), B, H, W)
    ab.t.normal_()
    groups, keep = [], []
    for g in range(ng):
        w = G.pack_weight(torch.randn(64, 64 * nsrc, ks, ks, device=G.DEV) * 0.05)
        bias = torch.randn(64, device=G.DEV)
        pb = torch.zeros(B, ab.a.tiles * 4, 64, device=G.DEV) if pool else None
        keep += [w, bias, pb]
        groups.append(G.group([g * (nsrc + 1) + i for i in range(nsrc)], g * (nsrc + 1) + nsrc, w, bias, act=K.ACT_LRELU, pool=pb,
                              res1=(g * (nsrc + 1) if res else -1)))
    for _ in range(2):
        G.run_conv(ab, groups, ksize=ks, impl=impl)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    arr = (K.ConvGroup * len(groups))(*groups)
    e0.record()
    for _ in range(reps):
        K.check(K.load().savsr_conv(G.ctx().handle, ab.a.handle, arr, len(groups), ks, 64, K.DST_ARENA, impl, G._stream()))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    flop = 2.0 * ng * B * H * W * 64 * 64 * nsrc * ks * ks
    if os.environ.get("CONV_DBG") and not hasattr(K.load(), "savsr_debug_conv_counters"):
        print("  (cycle counters need a library built with -DSAVSR_DEBUG_COUNTERS)")
        os.environ.pop("CONV_DBG")
    if os.environ.get("CONV_DBG"):
        import ctypes
        dbg = torch.zeros(148, 8, dtype=torch.int64, device=G.DEV)
        K.load().savsr_debug_conv_counters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        K.load().savsr_debug_conv_counters(G.ctx().handle, dbg.data_ptr())
        K.check(K.load().savsr_conv(G.ctx().handle, ab.a.handle, arr, len(groups), ks, 64, K.DST_ARENA, impl, G._stream()))
        torch.cuda.synchronize()
        K.load().savsr_debug_conv_counters(G.ctx().handle, None)
        d = dbg.float().cpu()
        d = d[d[:, 3] > 0]
        m = d.mean(0)
        print(f"  per CTA (avg over {len(d)}): tiles {m[3]:.1f} | MMA warp total {m[0]:.0f} cyc ({m[0]/m[3]:.0f}/tile), wait t_empty {m[1]:.0f} ({100*m[1]/m[0]:.0f}%), "
              f"wait a_full {m[2]:.0f} ({100*m[2]/m[0]:.0f}%), wait set_full {m[7]:.0f} | epilogue total {m[4]:.0f}, wait t_full {m[5]:.0f} ({100*m[5]/max(m[4],1):.0f}%) | producer wait a_empty {m[6]:.0f}")
    if os.environ.get("CONV_DBG"):
        print(f"  SM clock from clock64 / event time: {m[0] / us / 1e3:.2f} GHz; cycles per MMA (issuer 0 view): {m[0] / (m[3] * 36 * nsrc):.1f}")
    print(f"pool={int(pool)} res1={int(res)} ", end="")
    print(f"nsrc={nsrc} groups={ng} B={B} impl={a[3] if len(a) > 3 else 'halo'} k={ks}: {us:.1f} us/launch, {flop / us / 1e6:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
