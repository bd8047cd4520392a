#!/usr/bin/env python
"""Where does the time of building a plan for a NEW scale go (weights already packed)?  cProfile of SAVSR.plan_for."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import savsr_b200  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = savsr_b200.SAVSR().to(dev).eval()
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    x = torch.rand(b, 7, 3, 144, 180, device=dev)
    with torch.no_grad():
        for s in [(4, 4), (3.9, 3.9), (3.8, 3.8), (1.5, 4), (2.7, 2.7)]:
            net.set_scale(s)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            net.plan_for(x)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            y = net(x); torch.cuda.synchronize(); t2 = time.perf_counter()
            y = net(x); torch.cuda.synchronize(); t3 = time.perf_counter()
            print(f"scale {s}: plan_for {1e3 * (t1 - t0):.1f} ms (reported {net.last_plan_build_ms:.1f}), first forward incl. graph capture "
                  f"{1e3 * (t2 - t1):.1f} ms, next forward {1e3 * (t3 - t2):.2f} ms")
        net.set_scale((3.7, 3.7))
        pr = cProfile.Profile()
        pr.enable()
        net.plan_for(x)
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
