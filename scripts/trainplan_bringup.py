"""GPU bring-up of the native training plan: gradient parity against the CPU oracle, then step timing."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_checks as G  # noqa: E402

if __name__ == "__main__":
    what = sys.argv[1:] or ["parity"]
    if "parity" in what:
        try:
            print("trainplan parity", json.dumps(G.check_trainplan(), default=str))
        except AssertionError as e:
            print("PARITY ASSERT", str(e)[:3000])
    if "profile" in what:
        import savsr_b200
        from savsr_b200 import trainplan as TP
        from savsr_b200.engine import get_hw
        dev = torch.device("cuda", 0)
        torch.manual_seed(0)
        net = savsr_b200.SAVSR().to(dev)
        tr = TP.NativeTrainer(net, use_graph=False)
        lq = torch.rand(4, 7, 3, 64, 64, device=dev)
        gt = torch.rand(4, 3, 256, 256, device=dev)
        for _ in range(2):
            tr.step(lq, gt, (4, 4))
        plan = tr.plan_for(lq, (4, 4))
        tr.flat.g.zero_(); tr.weights.pack(torch.cuda.current_stream().cuda_stream)
        prof = plan.run_profiled()
        tot = sum(v["ms"] for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"{k:28s} {v['ms']:8.3f} ms  {v['ops']:5d} ops")
        print("total eager ms", round(tot, 2))
    if "kprof" in what:
        import savsr_b200
        from savsr_b200 import trainplan as TP
        from torch.profiler import ProfilerActivity, profile
        dev = torch.device("cuda", 0)
        torch.manual_seed(0)
        net = savsr_b200.SAVSR().to(dev)
        tr = TP.NativeTrainer(net, use_graph=True)
        lq = torch.rand(4, 7, 3, 64, 64, device=dev)
        gt = torch.rand(4, 3, 256, 256, device=dev)
        for _ in range(3):
            tr.step(lq, gt, (4, 4))
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                tr.step(lq, gt, (4, 4))
            torch.cuda.synchronize()
        rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0]
        rows.sort(key=lambda r: -r[1])
        tot = sum(r[1] for r in rows)
        print(f"kernel time per step {tot:.2f} ms in {sum(r[2] for r in rows)} kernels")
        for k, ms, n in rows[:45]:
            print(f"{ms:8.3f} ms {n:6d}x  {k[:110]}")
    if "time" in what:
        import savsr_b200
        from savsr_b200 import trainplan as TP
        from savsr_b200.engine import get_hw
        dev = torch.device("cuda", 0)
        torch.manual_seed(0)
        net = savsr_b200.SAVSR().to(dev)
        for graph in (False, True):
            tr = TP.NativeTrainer(net, use_graph=graph)
            scales = [(2, 2), (4, 4), (1.5, 4)]
            lq = torch.rand(4, 7, 3, 64, 64, device=dev)
            gts = {s: torch.rand(4, 3, *get_hw(64, 64, s), device=dev) for s in scales}
            for s in scales:
                tr.step(lq, gts[s], s)
            torch.cuda.synchronize()
            t0 = time.time()
            n = 9
            for i in range(n):
                loss = tr.step(lq, gts[scales[i % 3]], scales[i % 3])
            torch.cuda.synchronize()
            p = tr.plans[next(iter(tr.plans))]
            print(f"graph={graph}: {(time.time() - t0) / n * 1e3:.2f} ms/step, loss {float(loss):.5f}, launches {p.launches}, slots {p.n_slots}, tslots {p.n_tslots}, "
                  f"plan GB {p.nbytes / 2**30:.2f}")
            del tr
            net = savsr_b200.SAVSR().to(dev)
