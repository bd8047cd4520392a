"""CPU oracle for the SAVSR forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``savsr_b200``) never imports it and fails loudly without its CUDA library.

What this is
    A from-scratch *functional* restatement (plain fp32 PyTorch ops on a flat
    ``state_dict``) of the algorithm in the reference file
    ``lbasicsr/archs/savsr_arch.py``.  Each function cites the reference lines
    it restates.  All arithmetic primitives (conv2d, grid_sample, matmul,
    interpolate) are ATen library calls, exactly the third-party dependency the
    reference itself sits on (torch, un-pinned by the reference:
    ``requirements.txt:18`` says ``torch>=1.9``; this image's torch 2.11.0 is the
    de-facto pin).

Parity pin
    The reference ships no tests or golden vectors.  The
    oracle is therefore pinned against *outputs of the reference itself*:
    ``scripts/make_golden.py`` imports the unmodified reference from
    ``/root/reference`` (build container only), loads the seeded state_dict of
    ``oracle/state_dict_fixture.py`` with ``strict=True``, runs it, and commits
    outputs + per-stage probes under ``tests/golden/``.  ``tests/test_oracle_golden.py``
    checks this file against those vectors, and the SATU index vectors against
    the known-answer hashes of SURVEY.md appendix A.3.
    The train-mode restatement (``BN_TRAIN``: BatchNorm on batch statistics, used
    with autograd as the checker of the native training step, row f1) is pinned
    the same way: ``scripts/make_golden.py --train-only`` runs the unmodified
    reference in ``train()`` mode with its own ``CharbonnierLoss`` and commits the
    loss, every parameter's gradient norm, the small gradient tensors and a random
    projection of the whole gradient (``tests/golden/train_*.npz``); at generation
    time the oracle's loss agrees to 3e-8 and its gradients to 3e-6 / 1.7e-3 per
    tensor.

The oracle deliberately uses a different formulation from the reference where
that makes the maths explicit (OSA-Conv is evaluated with the four attentions
folded into a per-sample kernel; the SATU expert mix is evaluated matrix-free),
so agreement with the golden vectors is a meaningful check of the restatement.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]
BN_EPS = 1e-5


# --------------------------------------------------------------------------- sizes / scale
def normalize_scale(scale) -> Tuple[float, float]:
    """(s_h, s_w) as the reference indexes it: scale[0] -> height, scale[1] -> width."""
    if isinstance(scale, (int, float)):
        return (scale, scale)
    s = tuple(scale)
    if len(s) != 2:
        raise ValueError(f"scale must be a number or a (s_h, s_w) pair, got {scale!r}")
    return (s[0], s[1])


def get_hw(h: int, w: int, scale) -> Tuple[int, int]:
    """savsr_arch.py:745-751 (get_HW = get_HW_round): python round() = half-to-even."""
    s = normalize_scale(scale)
    return round(h * s[0]), round(w * s[1])


# --------------------------------------------------------------------------- SATU index path (numpy, bit-level)
def satu_rel_coord(n_out: int, s) -> np.ndarray:
    """R(.) of savsr_arch.py:331/333 in fp32 with CPU true-division semantics.

    q_i = (i + 0.5) / s ; R_i = (q_i - floor(q_i + 1e-3)) - 0.5   (all fp32, left to right)
    """
    i = np.arange(n_out, dtype=np.float32)
    q = (i + np.float32(0.5)) / np.float32(s)
    r = (q - np.floor(q + np.float32(1e-3))) - np.float32(0.5)
    return r.astype(np.float32)


def satu_cell(n_out: int, s) -> np.ndarray:
    """Integer source LR cell floor(q_i + 1e-3) (the integer part of savsr_arch.py:331)."""
    i = np.arange(n_out, dtype=np.float32)
    q = (i + np.float32(0.5)) / np.float32(s)
    return np.floor(q + np.float32(1e-3)).astype(np.int32)


def satu_base_norm(n_out: int, n_lr: int, s) -> np.ndarray:
    """Normalised base sampling coordinate of savsr_arch.py:270-280 (zero offset), fp32.

    The meshgrid is float64 but is cast to fp32 by ``torch.Tensor(grid)`` *before* any
    arithmetic, so every op below is fp32, evaluated left to right as written there.
    """
    g = np.arange(n_out, dtype=np.float64).astype(np.float32)
    g = (g + np.float32(0.5)) / np.float32(s) - np.float32(0.5)
    g = g * np.float32(2) / np.float32(n_lr - 1) - np.float32(1)
    return g.astype(np.float32)


def satu_unnormalize(g: np.ndarray, n_lr: int) -> np.ndarray:
    """ATen grid_sampler align_corners=True un-normalisation: ((g + 1) / 2) * (n - 1), fp32."""
    return ((g.astype(np.float32) + np.float32(1)) / np.float32(2)) * np.float32(n_lr - 1)


def satu_base_corner(n_out: int, n_lr: int, s) -> np.ndarray:
    """floor of the un-normalised base coordinate -> int32 top/left corner index (may be -1)."""
    return np.floor(satu_unnormalize(satu_base_norm(n_out, n_lr, s), n_lr)).astype(np.int32)


# --------------------------------------------------------------------------- small helpers
def _conv(sd: SD, name: str, x: Tensor, pad: Optional[int] = None) -> Tensor:
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    if pad is None:
        pad = w.shape[-1] // 2
    return F.conv2d(x, w, b, stride=1, padding=pad)


def _lrelu(x: Tensor, slope: float = 0.2) -> Tensor:
    return F.leaky_relu(x, slope)


BN_TRAIN = False   # tests of the training row (8f1) set this: BatchNorm then normalises with the batch statistics (nn.BatchNorm2d.train())


def _bn_eval(sd: SD, name: str, x: Tensor) -> Tensor:
    """BatchNorm2d (savsr_arch.py:26, 191-204).  Eval mode (default): affine with running stats.  With BN_TRAIN: batch mean and
    biased batch variance over (n, h, w), as nn.BatchNorm2d computes in train mode (running-stat updates are not modelled)."""
    g, b = sd[name + ".weight"], sd[name + ".bias"]
    shp = (1, -1, 1, 1)
    if BN_TRAIN:
        rm = x.mean(dim=(0, 2, 3))
        rv = x.var(dim=(0, 2, 3), unbiased=False)
    else:
        rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
    return (x - rm.view(shp)) / torch.sqrt(rv.view(shp) + BN_EPS) * g.view(shp) + b.view(shp)


# --------------------------------------------------------------------------- OSA-Conv
def osa_attention(sd: SD, prefix: str, pooled: Tensor, scale) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """scale_routing + ScaleAttention (savsr_arch.py:143-151, 91-96, 69-89).

    pooled: [b, Ci] spatial mean of the OSA-Conv input.  Returns
    ca [b,Ci], fa [b,Co], sa [b,9], ka [b,K].
    """
    s = normalize_scale(scale)
    b = pooled.shape[0]
    # NOTE order: (1/s_h, 1/s_w) -- ones/scale[0], ones/scale[1]  (143-145)
    inv = torch.ones(1, 1, device=pooled.device) / s[0], torch.ones(1, 1, device=pooled.device) / s[1]
    info = torch.cat(inv, 1).repeat(b, 1)
    v = torch.cat([info, pooled], dim=1)
    v = F.relu(F.linear(v, sd[prefix + ".scale_routing.0.weight"], sd[prefix + ".scale_routing.0.bias"]))
    v = F.relu(F.linear(v, sd[prefix + ".scale_routing.2.weight"], sd[prefix + ".scale_routing.2.bias"]))
    a = prefix + ".attention"
    z = F.conv2d(v.view(b, -1, 1, 1), sd[a + ".fc.weight"])
    z = F.relu(_bn_eval(sd, a + ".bn", z))
    ca = torch.sigmoid(_conv(sd, a + ".channel_fc", z)).flatten(1)
    fa = torch.sigmoid(_conv(sd, a + ".filter_fc", z)).flatten(1)
    sa = torch.sigmoid(_conv(sd, a + ".spatial_fc", z)).flatten(1)
    ka = torch.softmax(_conv(sd, a + ".kernel_fc", z).flatten(1), dim=1)
    return ca, fa, sa, ka


def osa_fold_weight(bank: Tensor, ca: Tensor, fa: Tensor, sa: Tensor, ka: Tensor) -> Tensor:
    """W'[b,o,i,u,v] = fa[b,o] ca[b,i] sa[b,u,v] sum_k ka[b,k] W[k,o,i,u,v]   (savsr_arch.py:156-171;
    the reference's own comment at 148-149 asserts the equivalence)."""
    b = ca.shape[0]
    k, co, ci, kh, kw = bank.shape
    w = torch.einsum("bk,koiuv->boiuv", ka, bank)
    w = w * sa.view(b, 1, 1, kh, kw) * ca.view(b, 1, ci, 1, 1) * fa.view(b, co, 1, 1, 1)
    return w


def osconv(sd: SD, prefix: str, x: Tensor, scale) -> Tensor:
    """OSConv2d._forward_impl_common (savsr_arch.py:139-172), folded-weight formulation."""
    b, ci, h, w = x.shape
    pooled = x.mean(dim=(2, 3))
    ca, fa, sa, ka = osa_attention(sd, prefix, pooled, scale)
    wf = osa_fold_weight(sd[prefix + ".weight"], ca, fa, sa, ka)  # [b,Co,Ci,3,3]
    co = wf.shape[1]
    out = F.conv2d(x.reshape(1, b * ci, h, w), wf.reshape(b * co, ci, 3, 3), None, 1, 1, 1, groups=b)
    return out.view(b, co, h, w)


# --------------------------------------------------------------------------- trunk blocks
def residual_block(sd: SD, prefix: str, xs: List[Tensor], scale) -> List[Tensor]:
    """ResidualBlock.forward (savsr_arch.py:399-415)."""
    n = len(xs)
    x1 = [_lrelu(_conv(sd, f"{prefix}.conv0.{i}", xs[i])) for i in range(n)]
    merged = torch.cat(x1, dim=1)
    if prefix + ".osconv.weight" in sd:
        base = _lrelu(osconv(sd, prefix + ".osconv", merged, scale))
    else:
        base = _lrelu(_conv(sd, prefix + ".conv1", merged))
    out = []
    for i in range(n):
        x2 = _lrelu(_conv(sd, f"{prefix}.conv2.{i}", torch.cat([base, x1[i]], dim=1)))
        out.append(xs[i] + x2)
    return out


def _count(sd: SD, pattern: str) -> int:
    """Number of consecutive indices i for which pattern.format(i) is a key."""
    n = 0
    while pattern.format(n) in sd:
        n += 1
    return n


def window_unit_l1(sd: SD, prefix: str, frames: Tensor, h_past: Tensor, scale) -> Tensor:
    """WindowUnit_l1.forward (savsr_arch.py:444-464). frames: [b,3,c,h,w] (prev, centre, next)."""
    b, t, c, h, w = frames.shape
    mid = t // 2
    x_c = frames[:, mid]
    x_sup = torch.cat([frames[:, i] for i in range(t) if i != mid], dim=1)
    h_sup = _lrelu(_conv(sd, prefix + ".conv_sup", x_sup))
    h_c = _lrelu(_conv(sd, prefix + ".conv_c", x_c))
    feats = [h_c, h_sup, h_past]
    for j in range(_count(sd, prefix + ".blocks.{}.conv0.0.weight")):
        feats = residual_block(sd, f"{prefix}.blocks.{j}", feats, scale)
    return _conv(sd, prefix + ".merge", torch.cat(feats, dim=1))


def window_unit_l2(sd: SD, prefix: str, xs: List[Tensor], scale) -> List[Tensor]:
    """WindowUnit_l2.forward (savsr_arch.py:485-501)."""
    ws = _count(sd, prefix + ".conv_h.{}.weight")
    sw = _count(sd, prefix + ".blocks.0.conv0.{}.weight")
    nb = _count(sd, prefix + ".blocks.{}.conv0.0.weight")
    hf = [_lrelu(_conv(sd, f"{prefix}.conv_h.{i}", xs[i])) for i in range(ws)]
    out = hf if len(hf) == 1 else []
    for i in range(ws - sw + 1):
        f = hf[i:i + sw]
        for j in range(nb):
            f = residual_block(sd, f"{prefix}.blocks.{j}", f, scale)
        out.append(_conv(sd, prefix + ".merge", torch.cat(f, dim=1)))
    return out


def rcab(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """RCAB + ChannelAttention (savsr_arch.py:504-549), res_scale = 1."""
    t = F.relu(_conv(sd, prefix + ".rcab.0", x))
    t = _conv(sd, prefix + ".rcab.2", t)
    y = t.mean(dim=(2, 3), keepdim=True)
    y = F.relu(_conv(sd, prefix + ".rcab.3.attention.1", y))
    y = torch.sigmoid(_conv(sd, prefix + ".rcab.3.attention.3", y))
    return t * y + x


def residual_group(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """ResidualGroup.forward (savsr_arch.py:552-571)."""
    t = x
    for j in range(_count(sd, prefix + ".residual_group.{}.rcab.0.weight")):
        t = rcab(sd, f"{prefix}.residual_group.{j}", t)
    return _conv(sd, prefix + ".conv", t) + x


def osadapt_mask(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """OSAdapt.mask (savsr_arch.py:189-206), BN in eval mode."""
    m = prefix + ".mask"
    t = F.relu(_bn_eval(sd, m + ".1", _conv(sd, m + ".0", x)))
    t = F.avg_pool2d(t, 2)
    t = F.relu(_bn_eval(sd, m + ".5", _conv(sd, m + ".4", t)))
    t = F.relu(_bn_eval(sd, m + ".8", _conv(sd, m + ".7", t)))
    t = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    return torch.sigmoid(_bn_eval(sd, m + ".12", _conv(sd, m + ".11", t)))


def osadapt(sd: SD, prefix: str, x: Tensor, scale) -> Tensor:
    """OSAdapt.forward (savsr_arch.py:210-214)."""
    return x + osconv(sd, prefix + ".adapt", x, scale) * osadapt_mask(sd, prefix, x)


# --------------------------------------------------------------------------- SATU
def satu_sta_conv(x: Tensor, kern: Tensor, ks: int = 5) -> Tensor:
    """sta_conv (savsr_arch.py:297-313):
    out[b,c,y,x] = sum_{u,v} xpad[b,c,y+u,x+v] * K[b, c*ks*ks + u*ks + v, y, x], replicate pad."""
    b, c, h, w = x.shape
    p = (ks - 1) // 2
    xp = F.pad(x, (p, p, p, p), mode="replicate")
    k = kern.view(b, c, ks, ks, h, w)
    out = torch.zeros_like(x)
    for u in range(ks):
        for v in range(ks):
            out = out + xp[:, :, u:u + h, v:v + w] * k[:, :, u, v]
    return out


def satu_mlp_input(h: int, w: int, scale, device=None) -> Tensor:
    """4-channel MLP input (savsr_arch.py:326-340): [1/s_w, 1/s_h, R_y, R_x] (note channel 0 = 1/s_w).
    Always evaluated with CPU (true-division) semantics, then moved to `device` (the device-aware mode only serves the
    GPU timing leg of bench.py; parity is defined on the CPU)."""
    s = normalize_scale(scale)
    H, W = get_hw(h, w, s)
    ry = torch.from_numpy(satu_rel_coord(H, s[0])).view(H, 1).expand(H, W)
    rx = torch.from_numpy(satu_rel_coord(W, s[1])).view(1, W).expand(H, W)
    c0 = torch.ones(H, W) / s[1]
    c1 = torch.ones(H, W) / s[0]
    return torch.stack([c0, c1, ry, rx], 0).unsqueeze(0).to(device)


def satu_heads(sd: SD, prefix: str, h: int, w: int, scale) -> Tuple[Tensor, Tensor, Tensor]:
    """body + offset / st_offset / routing heads (savsr_arch.py:344-351).
    Depends only on (scale, h, w).  Returns offset [1,2,H,W] (x then y), st_offset, routing [1,4,H,W]."""
    inp = satu_mlp_input(h, w, scale, sd[prefix + ".body.0.weight"].device)
    e = F.relu(_conv(sd, prefix + ".body.0", inp))
    e = F.relu(_conv(sd, prefix + ".body.2", e))
    off = _conv(sd, prefix + ".offset", e)
    st_off = _conv(sd, prefix + ".st_offset", e)
    r = torch.sigmoid(_conv(sd, prefix + ".routing.0", e))
    return off, st_off, r


def satu_grid(h: int, w: int, scale, offset: Tensor) -> Tensor:
    """Normalised sampling grid of STAUpsample.grid_sample (savsr_arch.py:262-288): [1,H,W,2] (x,y)."""
    s = normalize_scale(scale)
    H, W = get_hw(h, w, s)
    gx = torch.from_numpy(satu_base_norm(W, w, s[1])).to(offset.device).view(1, 1, W).expand(1, H, W)
    gy = torch.from_numpy(satu_base_norm(H, h, s[0])).to(offset.device).view(1, H, 1).expand(1, H, W)
    ox = offset[:, 0] * 2 / (w - 1)
    oy = offset[:, 1] * 2 / (h - 1)
    return torch.stack([gx + ox, gy + oy], dim=-1)


def satu_gather(x: Tensor, scale, offset: Tensor) -> Tensor:
    """Bilinear gather, zeros padding, align_corners=True (savsr_arch.py:291)."""
    b, _, h, w = x.shape
    grid = satu_grid(h, w, scale, offset).expand(b, -1, -1, -1)
    return F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def satu_expert_mix(sd: SD, prefix: str, fea0: Tensor, routing: Tensor) -> Tensor:
    """Spatially varying compress/expand (savsr_arch.py:353-370), matrix-free two-stage form:
    t = sum_e r_e (Wc_e f) ; out = sum_e r_e (We_e t) + f.   (r is a per-expert sigmoid, not softmax.)"""
    wc = sd[prefix + ".weight_compress"].flatten(2)  # [E, 8, 64]
    we = sd[prefix + ".weight_expand"].flatten(2)    # [E, 64, 8]
    r = routing[0]                                    # [E, H, W]
    u = torch.einsum("ekc,bchw->bekhw", wc, fea0)
    t = (u * r.unsqueeze(0).unsqueeze(2)).sum(1)      # [b, 8, H, W]
    v = torch.einsum("eck,bkhw->bechw", we, t)
    out = (v * r.unsqueeze(0).unsqueeze(2)).sum(1)
    return out + fea0


def satu(sd: SD, prefix: str, x: Tensor, scale, st_feat: Tensor, probes: Optional[dict] = None) -> Tensor:
    """STAUpsample.forward (savsr_arch.py:315-376)."""
    b, c, h, w = x.shape
    kern = _lrelu(_conv(sd, prefix + ".kernel_conv.0", st_feat), 0.1)
    sta = satu_sta_conv(x, kern)
    off, st_off, routing = satu_heads(sd, prefix, h, w, scale)
    fea0 = satu_gather(x, scale, off)
    fea = satu_expert_mix(sd, prefix, fea0, routing)
    sta_s = satu_gather(sta, scale, st_off)
    out = _conv(sd, prefix + ".fusion", torch.cat([sta_s, fea], dim=1))
    if probes is not None:
        probes.update(satu_sta=sta, satu_offset=off, satu_st_offset=st_off, satu_routing=routing,
                      satu_fea=fea, satu_sta_sampled=sta_s, satu_out=out)
    return out


# --------------------------------------------------------------------------- whole forward
def pad_spatial(x: Tensor, multiple: int = 2) -> Tensor:
    """SAVSR.pad_spatial (savsr_arch.py:670-690): reflect-pad bottom/right to a multiple of 2."""
    n, t, c, h, w = x.shape
    ph = (multiple - h % multiple) % multiple
    pw = (multiple - w % multiple) % multiple
    if ph == 0 and pw == 0:
        return x
    y = F.pad(x.reshape(-1, c, h, w), [0, pw, 0, ph], mode="reflect")
    return y.view(n, t, c, h + ph, w + pw)


def forward(sd: SD, x: Tensor, scale, probes: Optional[dict] = None) -> Tensor:
    """SAVSR.forward (savsr_arch.py:692-742) for the shipped configuration (interval=0).

    sd: flat reference-layout state_dict, fp32 tensors; x: [b, 7, 3, h, w] fp32 on the same device
    (CPU for every parity check; bench.py's `gpu_reference` leg may place both on a CUDA device to time the same ATen /
    cuDNN calls the reference makes there).  Returns [b, 3, H, W] fp32 (not clamped).
    """
    scale = normalize_scale(scale)
    b, t, c, h_in, w_in = x.shape
    H, W = get_hw(h_in, w_in, scale)
    nf = sd["conv_last.weight"].shape[0]
    centre = t // 2
    x_center = x[:, centre].contiguous()
    xp = pad_spatial(x)
    hp, wp = xp.shape[-2:]
    slid = 3
    n_it = t - slid + 1

    ht_f2p = torch.zeros(b, nf, hp, wp, device=x.device)
    ht_p2f = torch.zeros(b, nf, hp, wp, device=x.device)
    f2p: List[Tensor] = []
    p2f: List[Tensor] = []
    for idx in range(n_it):                                   # savsr_arch.py:708-719
        ct = t - 1 - slid // 2 - idx
        ht_f2p = window_unit_l1(sd, "f2p_win", xp[:, ct - 1:ct + 2], ht_f2p, scale)
        f2p.insert(0, ht_f2p)
        ct = idx + slid // 2
        ht_p2f = window_unit_l1(sd, "p2f_win", xp[:, ct - 1:ct + 2], ht_p2f, scale)
        p2f.append(ht_p2f)
    feats = [torch.cat([f2p[i], p2f[i]], dim=1) for i in range(n_it)]   # :721
    feats = window_unit_l2(sd, "h_win.0", feats, scale)                  # :722 (single l2 unit for t=7, fusion_win=5)
    hf = _lrelu(_conv(sd, "h_win_conv_h", feats[0]))                     # :723
    align = hf
    share = hf
    if probes is not None:
        probes.update(f2p_last=f2p[0], p2f_last=p2f[-1], h_win=feats[0], align=align)
    n_rg = _count(sd, "RG.{}.conv.weight")
    for i in range(n_rg):                                                # :727-732  (K = 1)
        hf = residual_group(sd, f"RG.{i}", hf)
        if probes is not None:
            probes[f"rg{i}"] = hf
        hf = osadapt(sd, f"adapt.{i}", hf, scale)
        if probes is not None:
            probes[f"adaptmod{i}"] = hf
        hf = hf + sd["gamma"] * share
    hf = _conv(sd, "conv_last", hf)                                      # :733
    if probes is not None:
        probes["conv_last"] = hf
    hf = hf + share                                                      # :734
    if probes is not None:
        probes["trunk"] = hf
    sr = satu(sd, "upsample", hf[..., :h_in, :w_in], scale, align[..., :h_in, :w_in], probes)   # :737
    sr = _conv(sd, "tail", sr)                                           # :738
    skip = F.interpolate(x_center, size=(H, W), mode="bilinear", align_corners=False)          # :739
    if probes is not None:
        probes.update(tail=sr, skip=skip)
    return sr + skip


# --------------------------------------------------------------------------- post-processing / metrics
def tensor2img(t: Tensor) -> np.ndarray:
    """tensor2img (img_util.py:38-94) for one RGB frame [3,H,W] in [0,1]: clamp -> HWC BGR -> (x * 255).round() -> uint8.
    numpy round = half to even."""
    x = t.detach().float().cpu().clamp(0, 1).numpy().transpose(1, 2, 0)[..., ::-1]
    return (x * 255.0).round().astype(np.uint8)


def y_channel(img_bgr_u8: np.ndarray) -> np.ndarray:
    """to_y_channel (metric_util.py:32-45) + bgr2ycbcr(y_only) (color_util.py:38-68): float32 in [16,235], not rounded."""
    img = img_bgr_u8.astype(np.float32) / 255.
    out = np.dot(img, [24.966, 128.553, 65.481]) + 16.0          # float64, as numpy promotes against the python list
    out = (out / 255.).astype(np.float32)                        # _convert_output_type_range for a float32 input
    return out * 255.


def psnr_y(a: Tensor, b: Tensor) -> float:
    """calculate_psnr(test_y_channel=True, crop_border=0) (psnr_ssim.py:11-48) on the reference's uint8 images.
    a, b: [3,H,W] or [n,3,H,W] RGB in [0,1].  Returns the mean PSNR over n (inf when identical)."""
    if a.dim() == 3:
        a, b = a[None], b[None]
    vals = []
    for p, q in zip(a, b):
        y1, y2 = y_channel(tensor2img(p)).astype(np.float64), y_channel(tensor2img(q)).astype(np.float64)
        mse = np.mean((y1 - y2) ** 2)
        vals.append(float("inf") if mse == 0 else 10. * np.log10(255. * 255. / mse))
    return float(np.mean(vals))


def _gaussian_window_11() -> np.ndarray:
    """cv2.getGaussianKernel(11, 1.5) outer itself (psnr_ssim.py:186-187): exp(-(i-5)^2 / (2 sigma^2)), normalised, float64."""
    x = np.arange(11, dtype=np.float64) - 5.0
    k = np.exp(-(x * x) / (2.0 * 1.5 * 1.5))
    k = k / k.sum()
    return np.outer(k, k)


def _filter_valid(img: np.ndarray, window: np.ndarray) -> np.ndarray:
    """cv2.filter2D(img, -1, window)[5:-5, 5:-5] (psnr_ssim.py:189): correlation, only the fully covered positions."""
    v = np.lib.stride_tricks.sliding_window_view(img, window.shape)      # [H-10, W-10, 11, 11]
    return np.einsum("ijkl,kl->ij", v, window)


def ssim_y(a: Tensor, b: Tensor) -> float:
    """calculate_ssim(test_y_channel=True, crop_border=0) (psnr_ssim.py:85-129, _ssim 172-200) on the reference's uint8
    images.  a, b: [3,H,W] or [n,3,H,W] RGB in [0,1].  Returns the mean SSIM over n."""
    if a.dim() == 3:
        a, b = a[None], b[None]
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    w = _gaussian_window_11()
    vals = []
    for p, q in zip(a, b):
        x, y = y_channel(tensor2img(p)).astype(np.float64), y_channel(tensor2img(q)).astype(np.float64)
        mu1, mu2 = _filter_valid(x, w), _filter_valid(y, w)
        mu1_sq, mu2_sq, mu12 = mu1 ** 2, mu2 ** 2, mu1 * mu2
        s1 = _filter_valid(x ** 2, w) - mu1_sq
        s2 = _filter_valid(y ** 2, w) - mu2_sq
        s12 = _filter_valid(x * y, w) - mu12
        m = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
        vals.append(float(m.mean()))
    return float(np.mean(vals))
