#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Imports lbasicsr from /root/reference (read-only; needs the version stub of SURVEY.md appendix E),
loads oracle/state_dict_fixture.make_state_dict(seed) with strict=True (pins the 791-key layout),
runs the reference forward on seeded inputs with hooks on the stage modules, checks the oracle
restatement against it, and writes compact golden vectors that travel to the GPU box.

    python scripts/make_golden.py            # rewrites tests/golden/
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import savsr_oracle as O                      # noqa: E402
from oracle.state_dict_fixture import make_input, make_state_dict, state_dict_spec  # noqa: E402

NETWORK_G = dict(type="SAVSR", num_in_ch=3, num_feat=64, num_frame=7, slid_win=3, fusion_win=5,
                 interval=0, w1_num_block=4, w2_num_block=2, n_resgroups=4, n_resblocks=8,
                 center_frame_idx=None)

CASES = [  # name, b, h, w, scale, sd_seed, in_seed
    ("x2_16x20", 1, 16, 20, (2, 2), 0, 1234),
    ("x4_16x16", 1, 16, 16, (4, 4), 0, 1235),
    ("x1p5x4_13x15", 1, 13, 15, (1.5, 4), 1, 1236),       # odd sizes -> pad_spatial + crop
    ("x2p7_b2_12x14", 2, 12, 14, (2.7, 2.7), 1, 1237),    # b > 1 -> grouped OSA-Conv
    ("x3_10x12", 1, 10, 12, (3, 3), 2, 1238),             # odd integer scale (floor ambiguity, A.3)
]


def load_reference():
    sys.path.insert(0, REF)
    v = types.ModuleType("lbasicsr.version")
    v.__version__, v.__gitsha__, v.version_info = "0.1.1", "unknown", (0, 1, 1)
    sys.modules["lbasicsr.version"] = v
    import lbasicsr  # noqa: F401
    from lbasicsr.archs import build_network
    import lbasicsr.archs.savsr_arch as ref_arch
    return build_network, ref_arch


def probe_summary(t: torch.Tensor) -> dict:
    t = t.detach().float().contiguous()
    flat = t.flatten()
    idx = torch.linspace(0, flat.numel() - 1, steps=min(257, flat.numel())).long()
    return dict(shape=np.array(t.shape, dtype=np.int64), mean=np.float64(flat.double().mean()),
                std=np.float64(flat.double().std()), absmax=np.float64(flat.abs().max()),
                sample=flat[idx].numpy().copy())


def metric_kat(outdir):
    import hashlib
    # --- metric chain: tensor2img + calculate_psnr(test_y_channel=True) of the reference itself
    from lbasicsr.metrics.psnr_ssim import calculate_psnr, calculate_ssim
    from lbasicsr.utils.img_util import tensor2img
    g = torch.Generator().manual_seed(99)
    gt = torch.rand(4, 3, 24, 31, generator=g)
    sr = gt + 0.04 * torch.randn(4, 3, 24, 31, generator=g)
    sr[0, :, :3] = 1.5; sr[0, :, 3:6] = -0.2
    sr[1, 0, 0, :31] = (torch.arange(31, dtype=torch.float32) * 2 + 0.5) / 255.0
    psnr, ssim, shas = [], [], []
    for i in range(4):
        a, b = tensor2img(sr[i]), tensor2img(gt[i])
        psnr.append(calculate_psnr(a, b, crop_border=0, test_y_channel=True))
        ssim.append(calculate_ssim(a, b, crop_border=0, test_y_channel=True))     # cv2.filter2D path of the reference
        shas.append(hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest())
        assert np.array_equal(a, O.tensor2img(sr[i])), "oracle tensor2img differs from the reference"
        assert abs(psnr[-1] - O.psnr_y(sr[i], gt[i])) < 1e-9, (psnr[-1], O.psnr_y(sr[i], gt[i]))
        assert abs(ssim[-1] - O.ssim_y(sr[i], gt[i])) < 1e-12, (ssim[-1], O.ssim_y(sr[i], gt[i]))
    np.savez_compressed(os.path.join(outdir, "metrics_kat.npz"), sr=sr.numpy(), gt=gt.numpy(), psnr_y=np.array(psnr, dtype=np.float64),
                        ssim_y=np.array(ssim, dtype=np.float64), img_sha1=np.array(shas))
    print("metric KAT: reference PSNR-Y", [round(p, 4) for p in psnr], "SSIM-Y", [round(p, 6) for p in ssim],
          "== oracle (1e-9 / 1e-12), uint8 images bit-identical")


def lr_kat(outdir):
    """Row f2: the reference's own as_mod_crop + img2tensor + arbitrary_scale_downsample (torchvision Resize, bicubic,
    antialias) on synthetic uint8 frames, stored as known answers for oracle/lr_synthesis.py and the CUDA data path."""
    from lbasicsr.data.data_util import arbitrary_scale_downsample
    from lbasicsr.data.transforms import as_mod_crop
    from lbasicsr.utils import img2tensor
    from oracle import lr_synthesis as L
    rng = np.random.default_rng(7)
    rec = {}
    cases = [("x4", 3, 50, 66, (4, 4)), ("x2p7", 2, 45, 61, (2.7, 2.7)), ("x1p5x4", 2, 47, 63, (1.5, 4)),
             ("x3p9", 2, 64, 50, (3.9, 3.9)), ("x1p2x1p7", 1, 37, 53, (1.2, 1.7)), ("x7p3x5p1", 1, 160, 120, (7.3, 5.1))]
    for name, t, h, w, scale in cases:
        frames = rng.integers(0, 256, size=(t, h, w, 3), dtype=np.uint8)
        frames[0, : h // 3] = np.where(rng.random((h // 3, w, 1)) < 0.5, 0, 255)        # hard edges: overshoot outside [0, 1]
        imgs = [as_mod_crop(f.astype(np.float32) / 255., scale) for f in frames]        # data_util.py:41-46
        gt = torch.stack(img2tensor(imgs, bgr2rgb=True, float32=True), dim=0)           # [t,3,hc,wc]
        lr = arbitrary_scale_downsample(gt, scale=scale, mode="torch")                  # video_test_dataset.py:312
        lr_o, gt_o = L.synthesize_lr(frames, scale)
        assert gt_o.shape == tuple(gt.shape) and np.array_equal(gt_o, gt.numpy()), name
        assert lr_o.shape == tuple(lr.shape), (name, lr_o.shape, lr.shape)
        assert np.array_equal(lr_o, lr.numpy()), (name, float(np.abs(lr_o - lr.numpy()).max()))
        rec[f"{name}.frames"] = frames
        rec[f"{name}.scale"] = np.array(scale, dtype=np.float64)
        rec[f"{name}.lr"] = lr.numpy()
        rec[f"{name}.crop"] = np.array(gt.shape[-2:], dtype=np.int64)
        print(f"lr KAT {name}: {h}x{w} -> crop {tuple(gt.shape[-2:])} -> lr {tuple(lr.shape[-2:])}, min {float(lr.min()):.3f} "
              f"max {float(lr.max()):.3f}; oracle bit-identical")
    # as_mod_crop sizes over a sweep of the YAML scales (transforms.py:47-69)
    sweep = []
    for sc in [(4, 4), (3.9, 3.9), (2.7, 2.7), (1.5, 4), (1.2, 1.2), (3.5, 2.5), (1.05, 1.95), (6.25, 6.25), (2, 3.14)]:
        for (h, w) in [(576, 720), (480, 704), (101, 67)]:
            got = as_mod_crop(np.zeros((h, w, 3), np.float32), sc).shape[:2]
            assert tuple(got) == L.as_mod_crop_size(h, w, sc), (sc, h, w, got, L.as_mod_crop_size(h, w, sc))
            sweep.append([h, w, sc[0], sc[1], got[0], got[1]])
    rec["crop_sweep"] = np.array(sweep, dtype=np.float64)
    np.savez_compressed(os.path.join(outdir, "lr_kat.npz"), **rec)


TRAIN_CASES = [  # name, b, h, w, scale, sd_seed, in_seed  (row f1: the reference's optimisation step, sr_model.py:101-128)
    ("train_x2_b2_12x14", 2, 12, 14, (2, 2), 0, 1240),
    ("train_x1p5x4_b3_9x11", 3, 9, 11, (1.5, 4), 1, 1241),    # odd sizes -> pad_spatial + crop; asymmetric scale; b = 3
]


def train_kat(outdir):
    """Row f1: the UNMODIFIED reference module in train() mode (BatchNorm on batch statistics) + its own CharbonnierLoss
    (losses/basic_loss.py:84-123, loss_weight 1, mean) + backward, on CPU.  Stored per case: the loss, probes of the output, the L2 norm of
    every parameter's gradient, the projection of the whole gradient on a seeded random direction, every gradient tensor of at most 4096
    elements in full, and the BatchNorm buffers after the step.  The oracle's train-mode restatement (oracle.savsr_oracle.BN_TRAIN + autograd)
    is checked against them here and in tests/test_oracle_golden.py -- it is the checker of the native training path."""
    build_network, _ = load_reference()
    from lbasicsr.losses.basic_loss import CharbonnierLoss
    crit = CharbonnierLoss(loss_weight=1.0, reduction="mean")
    torch.set_num_threads(os.cpu_count())
    net = build_network(dict(NETWORK_G))
    for name, b, h, w, scale, sd_seed, in_seed in TRAIN_CASES:
        sd = make_state_dict(sd_seed)
        net.load_state_dict(sd, strict=True)
        net.train()
        net.set_scale(scale)
        x = make_input(b, h, w, in_seed)
        H, W = O.get_hw(h, w, scale)
        gt = torch.rand(b, 3, H, W, generator=torch.Generator().manual_seed(in_seed + 7))
        net.zero_grad(set_to_none=True)
        out = net(x)
        loss = crit(out, gt)
        loss.backward()
        names = [k for k, p in net.named_parameters()]
        grads = {k: p.grad for k, p in net.named_parameters()}
        present = np.array([grads[k] is not None for k in names])
        norms = np.array([float(grads[k].double().norm()) if grads[k] is not None else 0.0 for k in names], dtype=np.float64)
        gen = torch.Generator().manual_seed(4242)
        proj = 0.0
        for k in names:
            r = torch.randn(dict(net.named_parameters())[k].shape, generator=gen, dtype=torch.float64)
            if grads[k] is not None:
                proj += float((grads[k].double() * r).sum())
        rec = dict(loss=np.float64(loss.detach().double()), gt=gt.numpy(), names=np.array(names), grad_present=present, grad_norm=norms,
                   grad_proj=np.float64(proj), scale=np.array(scale, dtype=np.float64), dims=np.array([b, h, w, sd_seed, in_seed], dtype=np.int64))
        for k, v in probe_summary(out).items():
            rec["out." + k] = v
        for k in names:
            if grads[k] is not None and grads[k].numel() <= 4096:
                rec["grad." + k] = grads[k].detach().numpy().copy()
        for k, v in net.state_dict().items():
            if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                rec["buf." + k] = v.detach().numpy().copy()
        # the oracle's train-mode restatement against it
        sd_o = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        O.BN_TRAIN = True
        try:
            out_o = O.forward(sd_o, x, scale)
            loss_o = torch.sqrt((out_o - gt) ** 2 + 1e-12).mean()
            loss_o.backward()
        finally:
            O.BN_TRAIN = False
        worst, worst_k = 0.0, None
        big = norms.max()
        for k in names:
            if grads[k] is None:
                assert sd_o[k].grad is None or float(sd_o[k].grad.abs().max()) == 0.0, k
                continue
            e = float((sd_o[k].grad - grads[k]).double().norm()) / max(float(grads[k].double().norm()), 1e-3 * big)
            if e > worst:
                worst, worst_k = e, k
        print(f"[{name}] reference train-mode loss {float(loss):.8f}; oracle loss diff {abs(float(loss_o) - float(loss)):.2e}, out max-abs "
              f"{float((out_o - out).abs().max()):.2e}; worst gradient tensor {worst_k}: {worst:.2e} (of |g| floored at 0.1 % of the largest); "
              f"{int(present.sum())}/{len(names)} parameters receive a gradient")
        assert abs(float(loss_o) - float(loss)) < 1e-6 and worst < 2e-3, (worst_k, worst)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **rec)


def window_kat(outdir):
    """Row f2: the reference's generate_frame_indices (data_util.py:63-112, padding='reflection' as in every SAVSR test YAML) for every
    frame of clips of 4..41 frames, 7- and 5-frame windows -- known answers for savsr_b200.sharding.frame_window_indices."""
    from lbasicsr.data.data_util import generate_frame_indices
    rec = {}
    for nf in (7, 5):
        for T in range(nf // 2 + 1, 42):
            rec[f"nf{nf}.T{T}"] = np.array([generate_frame_indices(i, T, nf, padding="reflection") for i in range(T)], dtype=np.int64)
    np.savez_compressed(os.path.join(outdir, "window_kat.npz"), **rec)
    print(f"window KAT: {len(rec)} (window, clip length) tables from the reference's generate_frame_indices")


def main():
    if "--metrics-only" in sys.argv or "--lr-only" in sys.argv or "--train-only" in sys.argv or "--window-only" in sys.argv:
        load_reference()
        if "--window-only" in sys.argv:
            window_kat(os.path.join(ROOT, "tests", "golden"))
        if "--metrics-only" in sys.argv:
            metric_kat(os.path.join(ROOT, "tests", "golden"))
        if "--lr-only" in sys.argv:
            lr_kat(os.path.join(ROOT, "tests", "golden"))
        if "--train-only" in sys.argv:
            train_kat(os.path.join(ROOT, "tests", "golden"))
        return
    build_network, ref_arch = load_reference()
    torch.set_num_threads(os.cpu_count())
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)

    # --- survey fingerprint of the default-init reference -------------
    torch.manual_seed(0)
    net0 = build_network(dict(NETWORK_G)).eval()
    psum = sum(float(p.double().sum()) for p in net0.parameters())
    print(f"default-init fingerprint: sum(params) = {psum:.6f}  (survey: 466.810371)")
    sd0 = {k: v.detach().clone() for k, v in net0.state_dict().items()}
    spec = state_dict_spec()
    assert [k for k, _, _ in spec] == list(sd0.keys()), "fixture key order differs from the reference"
    for k, shp, _ in spec:
        assert tuple(sd0[k].shape) == tuple(shp), (k, sd0[k].shape, shp)
    print(f"state_dict layout pinned: {len(spec)} keys, order and shapes identical to the reference")
    x0 = make_input(1, 64, 64, 1234)
    net0.set_scale((2, 2))
    with torch.no_grad():
        y_ref0 = net0(x0)
        y_or0 = O.forward(sd0, x0, (2, 2))
    print(f"cfg1 default init: ref mean {float(y_ref0.mean()):.8f} (survey 0.45701084)  "
          f"oracle-vs-ref max-abs {float((y_ref0 - y_or0).abs().max()):.3e}")
    cfg1 = dict(out_mean=np.float64(y_ref0.double().mean()), out_std=np.float64(y_ref0.double().std()),
                y0000=np.float32(y_ref0[0, 0, 0, 0]), param_sum=np.float64(psum),
                oracle_maxabs=np.float64((y_ref0 - y_or0).abs().max()))
    np.savez_compressed(os.path.join(outdir, "cfg1_default_init_fingerprint.npz"), **cfg1)

    net = build_network(dict(NETWORK_G)).eval()
    for name, b, h, w, scale, sd_seed, in_seed in CASES:
        sd = make_state_dict(sd_seed)
        net.load_state_dict(sd, strict=True)
        net.set_scale(scale)
        x = make_input(b, h, w, in_seed)
        probes = {}
        hooks = []

        def keep(key, pick=lambda o: o):
            def fn(_m, _i, o):
                probes[key] = pick(o).detach().clone()
            return fn
        hooks.append(net.f2p_win.register_forward_hook(keep("f2p_last")))
        hooks.append(net.p2f_win.register_forward_hook(keep("p2f_last")))
        hooks.append(net.h_win.register_forward_hook(keep("h_win", lambda o: o[0][0])))
        hooks.append(net.h_win_act.register_forward_hook(keep("align")))
        for i in range(4):
            hooks.append(net.RG[i].register_forward_hook(keep(f"rg{i}")))
            hooks.append(net.adapt[i].register_forward_hook(keep(f"adaptmod{i}")))
        hooks.append(net.conv_last.register_forward_hook(keep("conv_last")))
        hooks.append(net.upsample.register_forward_hook(keep("satu_out")))
        hooks.append(net.tail.register_forward_hook(keep("tail")))
        hooks.append(net.upsample.body.register_forward_pre_hook(
            lambda _m, i: probes.__setitem__("satu_mlp_input", i[0].detach().clone())))
        hooks.append(net.upsample.offset.register_forward_hook(keep("satu_offset")))
        hooks.append(net.upsample.st_offset.register_forward_hook(keep("satu_st_offset")))
        hooks.append(net.upsample.routing.register_forward_hook(keep("satu_routing")))
        grids = []
        orig_gs = ref_arch.F.grid_sample

        def spy(inp, grid, *a, **k):
            grids.append(grid.detach().clone())
            return orig_gs(inp, grid, *a, **k)
        ref_arch.F.grid_sample = spy
        try:
            with torch.no_grad():
                y_ref = net(x)
        finally:
            ref_arch.F.grid_sample = orig_gs
            for hk in hooks:
                hk.remove()

        oprobes = {}
        y_or = O.forward(sd, x, scale, oprobes)
        err = float((y_ref - y_or).abs().max())
        # bit-level checks of the index path against the tensors the reference actually built
        H, W = O.get_hw(h, w, scale)
        s = O.normalize_scale(scale)
        inp = probes["satu_mlp_input"]
        assert tuple(inp.shape) == (1, 4, H, W)
        assert torch.equal(inp, O.satu_mlp_input(h, w, scale)), "SATU coordinate features not bit-exact"
        g_or = O.satu_grid(h, w, scale, oprobes["satu_offset"])
        grid_bits = torch.equal(grids[0][0:1], g_or)
        gmax = float((grids[0][0:1] - g_or).abs().max())
        base_x = torch.from_numpy(O.satu_base_norm(W, w, s[1]))
        base_y = torch.from_numpy(O.satu_base_norm(H, h, s[0]))
        zero_grid = O.satu_grid(h, w, scale, torch.zeros(1, 2, H, W))
        assert torch.equal(zero_grid[0, 0, :, 0], base_x) and torch.equal(zero_grid[0, :, 0, 1], base_y)
        stage_err = {k: float((probes[k] - oprobes[k]).abs().max()) for k in probes if k in oprobes}
        print(f"[{name}] out {tuple(y_ref.shape)} oracle-vs-ref max-abs {err:.3e}; grid bit-exact={grid_bits} "
              f"(max diff {gmax:.2e}); worst stage {max(stage_err, key=stage_err.get)}={max(stage_err.values()):.3e}")
        assert err < 2e-5, err
        assert max(stage_err.values()) < 2e-4, stage_err

        rec = dict(b=np.int64(b), h=np.int64(h), w=np.int64(w), scale=np.array(s, dtype=np.float64),
                   sd_seed=np.int64(sd_seed), in_seed=np.int64(in_seed), out=y_ref.numpy(),
                   satu_mlp_input_sha1=np.array(hashlib.sha1(inp.numpy().tobytes()).hexdigest()),
                   grid0=grids[0][0].numpy(), grid1=grids[1][0].numpy())
        for k, t in probes.items():
            if k == "satu_mlp_input":
                continue
            for f, val in probe_summary(t).items():
                rec[f"probe.{k}.{f}"] = val
        np.savez_compressed(os.path.join(outdir, f"{name}.npz"), **rec)
    metric_kat(outdir)
    lr_kat(outdir)
    print("golden vectors written to", outdir)


if __name__ == "__main__":
    main()
