#!/usr/bin/env python
"""GPU bring-up: run every parity check in its own subprocess (a trapping kernel cannot poison the rest)
and log one JSON line per check to gpurun_out/bringup.log.  Usage on the GPU box:
    python scripts/gpu_bringup.py            # all checks
    python scripts/gpu_bringup.py --one NAME # one check, in-process
"""
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHECKS = {
    "pack": "check_pack_roundtrip()",
    "conv_check_k3": "check_conv(impl=K.IMPL_CHECK, ksize=3, nsrc=2, B=2, H=21, W=19)",
    "conv_tap_k3_s1": "check_conv(impl=K.IMPL_TAP, ksize=3, nsrc=1, B=1, H=32, W=24)",
    "conv_tap_k3_s3_ragged": "check_conv(impl=K.IMPL_TAP, ksize=3, nsrc=3, B=2, H=37, W=45)",
    "conv_tap_k3_s5": "check_conv(impl=K.IMPL_TAP, ksize=3, nsrc=5, B=1, H=48, W=40)",
    "conv_tap_k1_s3": "check_conv(impl=K.IMPL_TAP, ksize=1, nsrc=3, B=2, H=37, W=45)",
    "conv_tap_big": "check_conv(impl=K.IMPL_TAP, ksize=3, nsrc=2, B=2, H=144, W=180)",
    "conv_halo_s2_ragged": "check_conv(impl=K.IMPL_HALO, ksize=3, nsrc=2, B=2, H=37, W=45)",
    "conv_halo_s3_big": "check_conv(impl=K.IMPL_HALO, ksize=3, nsrc=3, B=2, H=144, W=180)",
    "conv_halo_s5": "check_conv(impl=K.IMPL_HALO, ksize=3, nsrc=5, B=1, H=48, W=40)",
    "conv_aux16": "check_conv_aux16()",
    "conv_aux16_halo": "check_conv_aux16(impl=K.IMPL_HALO)",
    "conv_per_sample_halo": "check_osa_conv_per_sample(impl=K.IMPL_HALO)",
    "conv_per_sample": "check_osa_conv_per_sample()",
    "pack_frames": "check_pack_frames()",
    "osa_prologue_192": "check_osa_prologue(ci=192)",
    "osa_prologue_320": "check_osa_prologue(ci=320, B=1)",
    "osa_prologue_64": "check_osa_prologue(ci=64, B=3)",
    "ca": "check_ca()",
    "mask": "check_mask()",
    "satu_index_1p5x4": "check_satu_index(144, 180, (1.5, 4))",
    "satu_index_2p7": "check_satu_index(144, 180, (2.7, 2.7))",
    "satu_index_x4": "check_satu_index(144, 180, (4, 4))",
    "satu_table": "check_satu_table()",
    "satu_kconv_sta": "check_satu_kconv_sta()",
    "satu_hr": "check_satu_hr()",
    "satu_hr_x4": "check_satu_hr(B=1, h=16, w=20, scale=(4, 4), seed=2)",
    "img_metrics": "check_img_metrics()",
    "lr_synthesis": "check_lr_synthesis()",
    "evaluate_clip": "check_evaluate_clip()",
    "conv_repeat_s1": "check_conv_repeatability(nsrc=1, ngroups=6, B=3)",
    "conv_repeat_s3": "check_conv_repeatability(nsrc=3, ngroups=2, B=4)",
    "forward_check_impl": "check_forward(b=1, h=16, w=20, scale=(2, 2), impl='check')",
    "forward_tap": "check_forward(b=1, h=16, w=20, scale=(2, 2), impl='tap')",
    "forward_tap_odd_b2": "check_forward(b=2, h=13, w=15, scale=(1.5, 4), sd_seed=1, in_seed=1236, impl='tap')",
    "forward_tap_graph": "check_forward(b=1, h=16, w=20, scale=(2.7, 2.7), impl='tap', graph=True)",
    "forward_halo": "check_forward(b=1, h=16, w=20, scale=(2, 2), impl='halo')",
    "forward_halo_fp16": "check_forward(b=1, h=16, w=20, scale=(2, 2), impl='halo', precision='fp16', tol=1e-3, stage_tol=0.01)",
    "forward_fp16_odd_b2": "check_forward(b=2, h=13, w=15, scale=(1.5, 4), sd_seed=1, in_seed=1236, impl='halo', precision='fp16', tol=1e-3, stage_tol=0.01)",
    "forward_fp16_x2p7_graph": "check_forward(b=1, h=16, w=20, scale=(2.7, 2.7), impl='halo', graph=True, precision='fp16', tol=1e-3, stage_tol=0.01)",
}


def run_one(name: str) -> int:
    import gpu_checks as G                      # noqa: F401
    from gpu_checks import K                    # noqa: F401
    t0 = time.time()
    try:
        res = eval("G." + CHECKS[name], {"G": G, "K": K})
        print(json.dumps(dict(check=name, ok=True, sec=round(time.time() - t0, 2), result=res), default=str))
        return 0
    except BaseException as e:                  # noqa: BLE001
        tb = traceback.format_exc().strip().splitlines()
        print(json.dumps(dict(check=name, ok=False, sec=round(time.time() - t0, 2), error=f"{type(e).__name__}: {e}"[:1500],
                              where=tb[-4:-1])))
        return 1


def main():
    if "--one" in sys.argv:
        sys.exit(run_one(sys.argv[sys.argv.index("--one") + 1]))
    names = [a for a in sys.argv[1:] if a in CHECKS] or list(CHECKS)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "a")
    nfail = 0
    for n in names:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True, timeout=300)
            lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
            line = lines[-1] if lines else json.dumps(dict(check=n, ok=False, error="no result", rc=p.returncode,
                                                         stdout=p.stdout[-800:], stderr=p.stderr[-1500:]))
        except subprocess.TimeoutExpired:
            line = json.dumps(dict(check=n, ok=False, error="timeout 300 s"))
        ok = json.loads(line).get("ok", False)
        nfail += 0 if ok else 1
        log.write(line + "\n"); log.flush()
        print(("PASS " if ok else "FAIL ") + line[:600], flush=True)
    print(f"bringup: {len(names) - nfail}/{len(names)} checks passed")


if __name__ == "__main__":
    main()
