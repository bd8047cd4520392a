#!/usr/bin/env python
"""Drive the reference's own training step -- `ASVSRModel.optimize_parameters` (lbasicsr/models/asvsr_model.py:21-29 ->
sr_model.py:101-128: net_g(lq), CharbonnierLoss, backward, torch.optim.Adam, model_ema) -- on synthetic Vimeo90K-shaped batches, twice:
once with savsr_b200.SAVSR served through savsr_b200.overlay (its train-mode forward is the native launch list behind one autograd
node), once with the unmodified reference arch; same initial weights, same batches.  Prints the loss trajectories and the time per
iteration.  Evidence for "the reference's training code runs the native path unchanged".

    python scripts/run_reference_training.py [iterations]

Needs a GPU and the offline install of the reference under baseline/_ref (git-ignored).
"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
SCALES = [(2, 2), (4, 4), (1.5, 4), (2.7, 2.7)]


def arm(name: str, iters: int, ckpt: str) -> None:
    sys.path.insert(0, ROOT)
    if name == "overlay":
        from savsr_b200 import overlay
        overlay.install(REF)
    else:
        sys.path.insert(0, REF)
    from lbasicsr.models import build_model
    torch.backends.cudnn.benchmark = True                      # lbasicsr/train.py sets it
    opt = {
        "name": "savsr_b200_train_check", "model_type": "ASVSRModel", "num_gpu": 1, "dist": False, "rank": 0, "world_size": 1, "is_train": True,
        "scale": (4, 4),
        "network_g": dict(type="SAVSR", num_in_ch=3, num_feat=64, num_frame=7, slid_win=3, fusion_win=5, interval=0, w1_num_block=4,
                          w2_num_block=2, n_resgroups=4, n_resblocks=8, center_frame_idx=None),
        "path": {"pretrain_network_g": None, "strict_load_g": True},
        "train": {"ema_decay": 0.999,
                  "optim_g": {"type": "Adam", "lr": 2e-4, "weight_decay": 0, "betas": [0.9, 0.99]},
                  "scheduler": {"type": "CosineAnnealingRestartLR", "periods": [300000], "restart_weights": [1], "eta_min": 1e-7},
                  "pixel_opt": {"type": "CharbonnierLoss", "loss_weight": 1.0, "reduction": "mean"}},
    }
    torch.manual_seed(0)
    model = build_model(opt)
    net = model.get_bare_model(model.net_g)
    if os.path.exists(ckpt):
        sd = torch.load(ckpt, map_location="cuda")
        net.load_state_dict(sd, strict=True)
        model.net_g_ema.load_state_dict(sd, strict=True)
    else:
        torch.save(net.state_dict(), ckpt)
    gen = torch.Generator().manual_seed(7)
    lq = torch.rand(4, 7, 3, 64, 64, generator=gen)
    gts = {s: torch.rand(4, 3, round(64 * s[0]), round(64 * s[1]), generator=gen) for s in SCALES}
    losses, times = [], []
    for it in range(iters):
        s = SCALES[it % len(SCALES)]
        torch.cuda.synchronize()
        t0 = time.time()
        model.feed_data({"lq": lq, "gt": gts[s], "scale": s})
        model.optimize_parameters(it)
        torch.cuda.synchronize()
        times.append((time.time() - t0) * 1e3)
        losses.append(float(model.log_dict["l_pix"]))
    warm = times[2 * len(SCALES):] or times
    print("ARM " + json.dumps(dict(arm=name, arch=type(net).__module__, native=bool(getattr(net, "__dict__", {}).get("_train_state")), losses=losses,
                                   ms_per_iter_after_warmup=round(sum(warm) / len(warm), 2))))


def main() -> None:
    if len(sys.argv) > 2 and sys.argv[1] == "--arm":
        arm(sys.argv[2], int(sys.argv[3]), sys.argv[4])
        return
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    if not os.path.isdir(os.path.join(REF, "lbasicsr")):
        raise SystemExit("baseline/_ref not installed (DESIGN.md section 2)")
    import tempfile
    ckpt = os.path.join(tempfile.gettempdir(), "savsr_b200_train_check_init.pth")
    if os.path.exists(ckpt):
        os.remove(ckpt)
    res = {}
    for name in ("reference", "overlay"):                      # the reference arm goes first and writes the initial weights
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--arm", name, str(iters), ckpt], capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("ARM ")]
        if not line:
            print(out.stdout[-3000:], out.stderr[-3000:])
            raise SystemExit(f"arm {name} failed")
        res[name] = json.loads(line[0][4:])
    os.remove(ckpt)
    a, b = res["reference"]["losses"], res["overlay"]["losses"]
    rel = max(abs(x - y) / max(abs(x), 1e-9) for x, y in zip(a, b))
    print(json.dumps(dict(iterations=iters, scales=SCALES, reference=res["reference"], overlay=res["overlay"], max_rel_loss_difference=rel), indent=1))
    assert res["overlay"]["native"], "the overlay arm did not take the native training path"
    assert rel < 2e-2, rel


if __name__ == "__main__":
    main()
