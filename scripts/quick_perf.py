#!/usr/bin/env python
"""Quick device timing of the forward at the Vid4 shape (bring-up aid; bench.py is the real harness)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import savsr_b200  # noqa: E402
from oracle.state_dict_fixture import make_state_dict  # noqa: E402


def main():
    impls = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tap"]
    batches = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 4]
    h, w, scale = 144, 180, (4, 4)
    dev = torch.device("cuda", 0)
    net = savsr_b200.SAVSR().to(dev).eval()
    net.load_state_dict(make_state_dict(0))
    net.set_scale(scale)
    for impl in impls:
        for b in batches:
            net.conv_impl = impl
            x = torch.rand(b, 7, 3, h, w, device=dev)
            try:
                with torch.no_grad():
                    plan = net.plan_for(x)
                    plan.x_in.copy_(x)
                    t0 = time.time(); plan.run(); torch.cuda.synchronize(); eager_ms = (time.time() - t0) * 1e3
                    plan.capture()
                    n = int(os.environ.get("QP_ITERS", "10"))
                    for _ in range(int(os.environ.get("QP_WARMUP", "3"))):
                        plan.run_graph()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(n):
                        plan.run_graph()
                    e1.record(); torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / n
                mpix = b * plan.H * plan.W / ms / 1e3
                print(json.dumps(dict(impl=impl, batch=b, ms_per_forward=round(ms, 3), ms_per_frame=round(ms / b, 3),
                                      hr_mpix_s=round(mpix, 2), first_eager_ms=round(eager_ms, 1), launches=plan.n_launches,
                                      tflops=round(b * 1208.4 / ms, 1))), flush=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps(dict(impl=impl, batch=b, error=str(e)[:300])), flush=True)
                return
            net.release_plans()


if __name__ == "__main__":
    main()
