#!/usr/bin/env python
"""Run the reference's own test pipeline (lbasicsr/test.py + a YAML in the reference's schema) twice on a synthetic PNG clip:
once with savsr_b200.SAVSR served through savsr_b200.overlay, once with the unmodified reference arch, same checkpoint,
and compare the PSNR-Y / SSIM-Y the harness logs.  Evidence for "lbasicsr/test.py and the YAML options run unchanged".

    python scripts/run_reference_harness.py [out_dir] [n_frames]

Needs a GPU and the offline install of the reference under baseline/_ref (git-ignored).
The YAML below is written for this test (keys per options/test/SAVSR/test_SAVSR_Vid4_asBI.yml: datasets.*.downsampling_scale,
network_g, path, val.metrics); nothing of the reference tree is modified.
"""
import os
import re
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")

YAML = """# written by scripts/run_reference_harness.py
name: {name}
model_type: ASVSRModel
num_gpu: 1
manual_seed: 0
datasets:
  test_01:
    name: Vid4_x4
    type: ASVideoTestDataset
    dataroot_gt: {gt}
    dataroot_lq: {gt}
    io_backend:
      type: disk
    cache_data: false
    num_frame: 7
    padding: reflection
    use_arbitrary_scale_downsampling: true
    downsampling_scale: !!python/tuple [4, 4]
    downsampling_mode: torch
  test_02:
    name: Vid4_x1.5x4
    type: ASVideoTestDataset
    dataroot_gt: {gt}
    dataroot_lq: {gt}
    io_backend:
      type: disk
    cache_data: false
    num_frame: 7
    padding: reflection
    use_arbitrary_scale_downsampling: true
    downsampling_scale: !!python/tuple [1.5, 4]
    downsampling_mode: torch
  test_03:
    name: Vid4_x2.7
    type: ASVideoTestDataset
    dataroot_gt: {gt}
    dataroot_lq: {gt}
    io_backend:
      type: disk
    cache_data: false
    num_frame: 7
    padding: reflection
    use_arbitrary_scale_downsampling: true
    downsampling_scale: !!python/tuple [2.7, 2.7]
    downsampling_mode: torch
network_g:
  type: SAVSR
  num_in_ch: 3
  num_feat: 64
  num_frame: 7
  slid_win: 3
  fusion_win: 5
  interval: 0
  w1_num_block: 4
  w2_num_block: 2
  n_resgroups: 4
  n_resblocks: 8
  center_frame_idx: ~
path:
  pretrain_network_g: {ckpt}
  strict_load_g: true
  resume_state: ~
  results_root: {results}
val:
  save_img: true
  suffix: ~
  metrics:
    psnr_y:
      type: calculate_psnr
      crop_border: 0
      test_y_channel: true
    ssim_y:
      type: calculate_ssim
      crop_border: 0
      test_y_channel: true
"""


def make_clip(gt_dir, n_frames, H, W):
    import cv2
    rng = np.random.default_rng(7)
    base = rng.random((H // 8 + 3, W // 8 + 3, 3)).astype(np.float32)
    for clip in ("synth_a", "synth_b"):
        os.makedirs(os.path.join(gt_dir, clip), exist_ok=True)
        for i in range(n_frames):
            shifted = np.roll(base, (i, 2 * i), axis=(0, 1))
            img = cv2.resize(shifted, (W + 16, H + 16), interpolation=cv2.INTER_CUBIC)[8:8 + H, 8:8 + W]
            img = np.clip(img + 0.03 * rng.standard_normal(img.shape).astype(np.float32), 0, 1)
            cv2.imwrite(os.path.join(gt_dir, clip, f"{i:08d}.png"), (img * 255).round().astype(np.uint8))
        base = rng.random(base.shape).astype(np.float32)


def metrics_from_log(text):
    """'# psnr_y: 27.1234' style lines of VideoBaseModel._log_validation_metric_values, per dataset."""
    out = {}
    cur = None
    for line in text.splitlines():
        m = re.search(r"Validation (\S+)", line)
        if m:
            cur = m.group(1)
        m = re.search(r"#\s*(psnr_y|ssim_y):\s*([0-9.]+)", line)
        if m and cur:
            out.setdefault(cur, {})[m.group(1)] = float(m.group(2))
    return out


def main():
    report_dir = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "harness"))
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    H, W = 288, 360                                      # GT size: x4 -> 72x90 LR, x1.5/x4 -> 192x90, x2.7 -> mod-cropped
    import tempfile
    out_dir = tempfile.mkdtemp(prefix="savsr_harness_")   # PNG clips and results stay out of the report directory
    os.makedirs(report_dir, exist_ok=True)
    if not os.path.isdir(os.path.join(REF, "lbasicsr")):
        raise SystemExit("baseline/_ref/lbasicsr missing: install the reference offline first (DESIGN.md)")
    gt_dir = os.path.join(out_dir, "data", "GT")
    make_clip(gt_dir, n_frames, H, W)
    import savsr_b200
    torch.manual_seed(0)
    ckpt = os.path.join(out_dir, "random_init_seed0.pth")
    torch.save({"params": savsr_b200.SAVSR().state_dict()}, ckpt)             # reference checkpoint layout (base_model.py:211-256)
    summary = {}
    for arm in ("savsr_b200", "reference"):
        yml = os.path.join(out_dir, f"test_{arm}.yml")
        open(yml, "w").write(YAML.format(name=f"harness_{arm}", gt=gt_dir, ckpt=ckpt, results=os.path.join(out_dir, "results")))
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, REF]))
        if arm == "savsr_b200":
            cmd = [sys.executable, "-m", "savsr_b200.overlay", REF, "lbasicsr/test.py", "-opt", yml]
        else:
            cmd = [sys.executable, os.path.join(REF, "lbasicsr", "test.py"), "-opt", yml]
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=1500)
        dt = time.time() - t0
        log = r.stdout + r.stderr
        open(os.path.join(report_dir, f"{arm}.log"), "w").write(log.replace(out_dir, "<work>"))
        if r.returncode != 0:
            print(log[-3000:])
            raise SystemExit(f"{arm}: lbasicsr/test.py failed with exit code {r.returncode}")
        summary[arm] = dict(seconds=round(dt, 1), metrics=metrics_from_log(log),
                            arch_module=("savsr_b200 via overlay" if arm == "savsr_b200" else "lbasicsr.archs.savsr_arch (unmodified)"))
    lines = [f"# lbasicsr/test.py on a synthetic clip (2 clips x {n_frames} frames, GT {H}x{W}), same random-init checkpoint", ""]
    lines.append("| dataset | metric | reference arch | savsr_b200 (overlay) | diff |")
    lines.append("|---|---|---|---|---|")
    worst = 0.0
    for ds, m in sorted(summary["reference"]["metrics"].items()):
        for k, v in sorted(m.items()):
            o = summary["savsr_b200"]["metrics"].get(ds, {}).get(k, float("nan"))
            lines.append(f"| {ds} | {k} | {v:.4f} | {o:.4f} | {o - v:+.4f} |")
            if k == "psnr_y":
                worst = max(worst, abs(o - v))
    lines.append("")
    lines.append(f"wall time of the whole pipeline (PNG read, CPU LR synthesis, net, PNG write, metrics): reference arch "
                 f"{summary['reference']['seconds']} s, savsr_b200 {summary['savsr_b200']['seconds']} s (includes plan builds for 3 scales)")
    lines.append(f"largest |PSNR-Y difference| = {worst:.4f} dB (gate: 0.05 dB on the bf16 path)")
    text = "\n".join(lines)
    n_png = sum(len(f) for _, _, f in os.walk(os.path.join(out_dir, "results")))
    text += f"\nPNG files written by the two runs: {n_png}"
    open(os.path.join(report_dir, "summary.md"), "w").write(text + "\n")
    print(text)
    if not summary["savsr_b200"]["metrics"] or worst > 0.05:
        raise SystemExit("harness comparison failed")


if __name__ == "__main__":
    main()
