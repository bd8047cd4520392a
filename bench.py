#!/usr/bin/env python
"""SAVSR forward hot-path benchmark (contract: see task brief / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vid4_x4] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" = super-resolving one synthetic Vid4-shaped clip (34 frames, 144x180 LR -> 576x720 HR at x4;
every output frame is an independent 7-frame window, lbasicsr/models/video_base_model.py:50-59), in batches
of `--batch` windows.  With N GPUs every rank processes its own clip (frames shard with no data-path
collective; weak scaling).  Rank 0 prints ONE JSON line:

  value      HR Mpix/s, whole job, windows already resident in HBM when the timed region starts
  e2e        same metric through the public API (savsr_b200.sharding.infer_clip on the savsr_b200.SAVSR module) with
             the LR clip in pinned HOST memory and the HR result copied back to pinned host memory every step
  roofline   dominant kernel (tcgen05 implicit-GEMM conv, N = 64): algorithmic FLOPs / CUDA-event time of those launches
             measured live in an eager pass, against the measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle (CPU restatement of the reference, kind "port") on the host cores, bounded sample
  pipeline   (informational) the reference's whole test loop minus the PNG codec through savsr_b200.datapath.evaluate_clip:
             uint8 ground-truth frames in pinned host memory -> device LR synthesis -> net -> uint8 images + PSNR-Y + SSIM-Y

--impl reference times the reference's own CPU path (the pinned oracle port; the reference itself is a Python
package that is not present on the GPU box) on this arm's workload/metric, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name: (frames, h, w, scale)
    "vid4_x4": (34, 144, 180, (4, 4)),
    "vid4_x1.5x4": (34, 144, 180, (1.5, 4)),
    "vid4_x2.7": (34, 144, 180, (2.7, 2.7)),
    "udm10_x4": (32, 180, 318, (4, 4)),
    "cfg1_x2": (7, 64, 64, (2, 2)),
}


def flops_per_frame(h, w, H, W):
    """F = 2 * [22 888 128 hp wp + 104 000 h w + 19 904 H W]."""
    hp, wp = h + (h & 1), w + (w & 1)
    return 2.0 * (22888128.0 * hp * wp + 104000.0 * h * w + 19904.0 * H * W)


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons sampled while the timed region runs: NVML every 20 ms when `pynvml` imports,
    else one nvidia-smi call per 200 ms."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()   # rows: (sm MHz, max MHz, watts, {reasons})
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; CUDA_VISIBLE_DEVICES may remap torch's index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        watts = n.nvmlDeviceGetPowerUsage(h) / 1e3
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.rows.append((float(sm), float(mx), watts, {name for name, bit in self.BITS if mask & bit}))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [c.strip() for c in out.split(",")]
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.rows.append((float(r[0]), float(r[1]), float(r[2]), {nm for nm, v in zip(names, r[3:7]) if v.lower().startswith("active")}))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self.nvml is not None else 0.2)

    def stop(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [r[0] for r in self.rows]
        reasons = set().union(*[r[3] for r in self.rows]) if self.rows else set()
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(r[1] for r in self.rows) if self.rows else None,
                    reasons=sorted(reasons), samples=len(sm), power_w_avg=round(statistics.median(r[2] for r in self.rows), 1) if self.rows else None,
                    source="nvml" if self.nvml is not None else "nvidia-smi")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained", 1418.2), bf16_burst=d.get("bf16_tflops", 1675.7),
                    hbm_gbs=d.get("hbm_gbs", 6543.7), source="measured")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm_gbs=6650.0, source="fallback")


def cpu_oracle_time(sd_cpu, frames, h, w, scale, threads=None):
    """Time the CPU oracle (restatement of the reference) on `frames` windows; returns (HR Mpix/s, seconds, threads)."""
    from oracle import savsr_oracle as O           # only this leg of bench.py may touch oracle/
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    H, W = O.get_hw(h, w, scale)
    x = torch.rand(1, 7, 3, h, w, generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        O.forward(sd_cpu, x[:, :, :, : min(h, 32), : min(w, 32)].contiguous(), scale)   # warm-up (thread pool, oneDNN primitives)
        t0 = time.perf_counter()
        for _ in range(frames):
            O.forward(sd_cpu, x, scale)
        dt = time.perf_counter() - t0
    return frames * H * W / dt / 1e6, dt, threads


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores, rank 0 only."""
    if rank != 0:
        return
    import savsr_b200
    frames, h, w, scale = WORKLOADS[args.workload]
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in savsr_b200.SAVSR().state_dict().items()}
    H, W = savsr_b200.get_HW(h, w, scale)
    per_step = 1                                     # bounded sample: one output frame (one 7-frame window) per step
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_time(sd, 1, h, w, scale)
    t_total, n = 0.0, 0
    for _ in range(args.steps):
        _, dt, threads = cpu_oracle_time(sd, per_step, h, w, scale)
        t_total += dt; n += per_step
    val = n * H * W / t_total / 1e6
    sample = f"{per_step} output frame(s) of the {args.workload} clip per step ({h}x{w} LR -> {H}x{W}), fp32, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": "hr_mpix_per_s", "value": round(val, 5), "unit": "HR Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_total / args.steps, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "frames_per_clip": frames, "lr": [h, w], "hr": [H, W], "scale": list(scale),
                   "note": "reference CPU path = oracle port of lbasicsr/archs/savsr_arch.py (pinned to the reference by tests/golden)"},
        "cpu_baseline": {"value": round(val, 5), "unit": "HR Mpix/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 5), "unit": "HR Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vid4_x4", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=17, help="windows per forward")
    ap.add_argument("--conv-impl", default=os.environ.get("SAVSR_CONV_IMPL", "halo"), choices=["halo", "tap"])
    ap.add_argument("--precision", default=os.environ.get("SAVSR_PRECISION", "bf16"), choices=["bf16", "fp16"],
                    help="16-bit operand format (fp32 accumulate): bf16 = throughput path, fp16 = <=1e-3 max-abs path, same speed")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import savsr_b200
    from savsr_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"

    frames, h, w, scale = WORKLOADS[args.workload]
    H, W = savsr_b200.get_HW(h, w, scale)
    torch.manual_seed(0)                                   # random-init weights of the shipped architecture
    net = savsr_b200.SAVSR().to(dev).eval()
    net.conv_impl = args.conv_impl
    net.precision = args.precision
    net.set_scale(scale)
    B = min(args.batch, frames)

    # synthetic clip (one per rank), U[0,1) fp32; pinned host copy for the e2e leg
    gen = torch.Generator().manual_seed(1234 + rank)
    clip_host = torch.rand(frames, 3, h, w, generator=gen).pin_memory()
    out_host = torch.empty(frames, 3, H, W).pin_memory()
    windows = sharding.gather_windows(clip_host.to(dev), list(range(frames)))   # [frames, 7, 3, h, w] resident in HBM
    out_dev = torch.empty(frames, 3, H, W, device=dev)
    batches = [(i, min(i + B, frames)) for i in range(0, frames, B)]
    plans = {}
    with torch.no_grad():
        for (a, b) in batches:
            if b - a not in plans:
                plan = net.plan_for(windows[a:b])
                plan.x_in.copy_(windows[a:b]); plan.capture()
                plans[b - a] = plan
    launches_per_step = sum(plans[b - a].n_launches for a, b in batches)

    def step_resident():
        for (a, b) in batches:
            p = plans[b - a]
            p.x_in.copy_(windows[a:b], non_blocking=True)
            p.run_graph()
            out_dev[a:b].copy_(p.out, non_blocking=True)

    def step_e2e():
        # public API, host buffers: H2D of the LR clip, window gather + forward per batch, D2H of the HR frames
        clip = clip_host.to(dev, non_blocking=True)
        with torch.no_grad():
            sharding.infer_clip(net, clip, batch=B, out=out_host)      # D2H of batch i overlaps the forward of batch i+1
        torch.cuda.current_stream().synchronize()

    # reference test loop on the device (rows f2 + f3): uint8 ground-truth frames in pinned host memory -> LR synthesis ->
    # windows -> net -> uint8 BGR images + PSNR-Y + SSIM-Y back in host memory (PNG decode / encode stay outside)
    from savsr_b200 import datapath
    pipeline_ok = datapath.as_mod_crop_size(H, W, scale) == (H, W) and datapath.lr_size(H, W, scale) == (h, w)
    if pipeline_ok:
        gt_u8_host = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, generator=gen).pin_memory()
        img_host = torch.empty(frames, H, W, 3, dtype=torch.uint8).pin_memory()
        met_host = torch.empty(2, frames, dtype=torch.float64).pin_memory()

    def step_pipeline():
        gt_u8 = gt_u8_host.to(dev, non_blocking=True)
        with torch.no_grad():
            res = datapath.evaluate_clip(net, gt_u8, scale, batch=B)
        img_host.copy_(res["images"], non_blocking=True)
        met_host[0].copy_(res["psnr_y"], non_blocking=True)
        met_host[1].copy_(res["ssim_y"], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank); sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    ms_pipe = None
    if pipeline_ok:
        for _ in range(2):
            step_pipeline()
        ms_pipe = timed(step_pipeline, args.steps)

    mpix_step = frames * H * W / 1e6 * world
    value = mpix_step * args.steps / (ms_total / 1e3)
    e2e = mpix_step * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel, measured live (eager pass, CUDA events around every op)
    peaks = measured_peaks()
    big = plans[max(plans)]
    prof = big.run_profiled()
    prof = big.run_profiled()
    conv = prof.get("conv3x3_n64", dict(ms=0.0, flops=0.0, launches=0))
    total_ms = sum(d["ms"] for d in prof.values())
    achieved = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] else 0.0
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_note = tj.get("dram_bytes_per_launch"), f"ncu --set full, one launch: {tj.get('launch')} ({tj.get('source')})"
    roofline = {"bound": "tensor", "kernel": "conv_igemm_kernel<64,3,halo> (tcgen05 implicit-GEMM 3x3 conv, all launches of one forward)",
                "achieved": round(achieved, 1), "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["bf16_sustained"], 4), "peak_source": peaks["source"] + " bf16 sustained",
                "traffic": traffic, "traffic_note": traffic_note, "share_of_step": round(conv["ms"] / total_ms, 3) if total_ms else None,
                "launches": conv["launches"], "avg_launch_us": round(1e3 * conv["ms"] / max(conv["launches"], 1), 1),
                "per_kind_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}}
    whole = flops_per_frame(h, w, H, W) * frames * world * args.steps / (ms_total / 1e3) / 1e12

    line = {
        "metric": "hr_mpix_per_s", "value": round(value, 2), "unit": "HR Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": args.workload, "frames_per_clip": frames, "clips": world, "lr": [h, w], "hr": [H, W],
                   "scale": list(scale), "windows_per_forward": B, "conv_impl": args.conv_impl, "weights": "random init (seed 0)",
                   "l2": "per-step working set (activation arenas, several GB) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"frame-sharded x{world}, no data-path collective"},
        "frames_per_s": round(frames * world * args.steps / (ms_total / 1e3), 2),
        "whole_forward_tflops": round(whole, 1),
        "e2e": {"value": round(e2e, 2), "unit": "HR Mpix/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": clip_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                "api": "savsr_b200.sharding.infer_clip(savsr_b200.SAVSR, clip)"},
        "pipeline": None if ms_pipe is None else {
            "value": round(mpix_step * args.steps / (ms_pipe / 1e3), 2), "unit": "HR Mpix/s", "ms_per_step": round(ms_pipe / args.steps, 3),
            "h2d_bytes_per_step": frames * H * W * 3, "d2h_bytes_per_step": frames * H * W * 3 + 16 * frames,
            "api": "savsr_b200.datapath.evaluate_clip(net, uint8 GT frames, scale)",
            "what": "the reference test loop minus the PNG codec: uint8 GT frames in pinned host memory -> as_mod_crop + antialiased "
                    "bicubic LR synthesis -> windows -> net -> uint8 BGR images, PSNR-Y and SSIM-Y back in host memory"},
        "gpu_launches": launches_per_step * args.steps * world,
        "clocks": clocks,
        "roofline": roofline,
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        n_cpu = 3
        val, dt, threads = cpu_oracle_time(sd, n_cpu, h, w, scale)
        line["cpu_baseline"] = {"value": round(val, 5), "unit": "HR Mpix/s", "cores": threads, "kind": "port",
                                "sample": f"{n_cpu} output frames of the same clip shape ({h}x{w} -> {H}x{W}), oracle fp32, {dt:.1f} s"}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
