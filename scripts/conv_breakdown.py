#!/usr/bin/env python
"""Per conv configuration (groups per launch x sources per group) time and TFLOP/s inside a Vid4-shaped forward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import savsr_b200  # noqa: E402
from oracle.state_dict_fixture import make_state_dict  # noqa: E402


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    dev = torch.device("cuda", 0)
    net = savsr_b200.SAVSR().to(dev).eval()
    net.load_state_dict(make_state_dict(0))
    net.set_scale((4, 4))
    net.conv_impl = "halo"
    x = torch.rand(b, 7, 3, 144, 180, device=dev)
    with torch.no_grad():
        plan = net.plan_for(x)
        plan.x_in.copy_(x)
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        acc = {}
        reps = 5
        for _ in range(reps):
            for k, v in plan.run_profiled(detail=True).items():
                d = acc.setdefault(k, dict(ms=0.0, flops=v["flops"], ops=v["ops"]))
                d["ms"] += v["ms"] / reps
    tot = sum(v["ms"] for v in acc.values())
    print(f"batch {b}: {tot:.2f} ms per forward (eager, per-op events)")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]["ms"]):
        tf = v["flops"] / v["ms"] / 1e9 if v["flops"] else 0.0
        print(f"  {k:24s} ops {v['ops']:4d}  {v['ms']:7.3f} ms ({100 * v['ms'] / tot:4.1f}%)  {1e3 * v['ms'] / v['ops']:7.1f} us/op  {tf:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
