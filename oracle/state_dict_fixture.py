"""Seeded, randomised SAVSR state_dict fixture -- TEST INFRASTRUCTURE (see savsr_oracle.py header).

The reference's default init leaves BatchNorm running stats at (0, 1), gamma at 1 and every
attention-head bias at 0, which is too benign to catch BN-folding or bias bugs (SURVEY.md
section 4, item 3).  This fixture perturbs all of them.  It is generated from (seed, key name)
alone, so the reference (in the build container), the oracle and the CUDA path (on the GPU box)
can all be fed bit-identical weights without shipping a 75 MB checkpoint.

Key names / shapes restate SURVEY.md appendix B (the reference's 791-key layout for the shipped
YAML: num_feat=64, num_frame=7, slid_win=3, fusion_win=5, w1_num_block=4, w2_num_block=2,
n_resgroups=4, n_resblocks=8); ``scripts/make_golden.py`` asserts ``strict=True`` loading into
the unmodified reference, which pins the layout.
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from typing import List, Tuple

import torch

NF = 64
KERNEL_NUM = 8


def _conv(spec, name, co, ci, k, bias=True):
    spec.append((name + ".weight", (co, ci, k, k), "w"))
    if bias:
        spec.append((name + ".bias", (co,), "b"))


def _bn(spec, name, c):
    spec.append((name + ".weight", (c,), "bn_w"))
    spec.append((name + ".bias", (c,), "bn_b"))
    spec.append((name + ".running_mean", (c,), "bn_rm"))
    spec.append((name + ".running_var", (c,), "bn_rv"))
    spec.append((name + ".num_batches_tracked", (), "bn_n"))


def _osconv(spec, name, ci, co):
    a = max(int(ci * 0.0625), 16)
    spec.append((name + ".weight", (KERNEL_NUM, co, ci, 3, 3), "bank"))
    _conv(spec, name + ".attention.fc", a, ci, 1, bias=False)
    _bn(spec, name + ".attention.bn", a)
    _conv(spec, name + ".attention.channel_fc", ci, a, 1)
    _conv(spec, name + ".attention.filter_fc", co, a, 1)
    _conv(spec, name + ".attention.spatial_fc", 9, a, 1)
    _conv(spec, name + ".attention.kernel_fc", KERNEL_NUM, a, 1)
    spec.append((name + ".scale_routing.0.weight", (2 * ci, ci + 2), "lin_w"))
    spec.append((name + ".scale_routing.0.bias", (2 * ci,), "b"))
    spec.append((name + ".scale_routing.2.weight", (ci, 2 * ci), "lin_w"))
    spec.append((name + ".scale_routing.2.bias", (ci,), "b"))


def _res_block(spec, name, nfr, use_os):
    for i in range(nfr):
        _conv(spec, f"{name}.conv0.{i}", NF, NF, 3)
    if use_os:
        _osconv(spec, name + ".osconv", NF * nfr, NF)
    else:
        _conv(spec, name + ".conv1", NF, NF * nfr, 1)
    for i in range(nfr):
        _conv(spec, f"{name}.conv2.{i}", NF, 2 * NF, 3)


def state_dict_spec() -> List[Tuple[str, tuple, str]]:
    spec: list = [("gamma", (1,), "gamma")]
    for win in ("f2p_win", "p2f_win"):
        _conv(spec, win + ".conv_c", NF, 3, 3)
        _conv(spec, win + ".conv_sup", NF, 6, 3)
        for j in range(4):
            _res_block(spec, f"{win}.blocks.{j}", 3, use_os=j >= 1)
        _conv(spec, win + ".merge", NF, 3 * NF, 3)
    for i in range(5):
        _conv(spec, f"h_win.0.conv_h.{i}", NF, 2 * NF, 3)
    for j in range(2):
        _res_block(spec, f"h_win.0.blocks.{j}", 5, use_os=True)
    _conv(spec, "h_win.0.merge", 2 * NF, 5 * NF, 3)
    _conv(spec, "h_win_conv_h", NF, 2 * NF, 3)
    for g in range(4):
        for r in range(8):
            p = f"RG.{g}.residual_group.{r}.rcab"
            _conv(spec, p + ".0", NF, NF, 3)
            _conv(spec, p + ".2", NF, NF, 3)
            _conv(spec, p + ".3.attention.1", NF // 16, NF, 1)
            _conv(spec, p + ".3.attention.3", NF, NF // 16, 1)
        _conv(spec, f"RG.{g}.conv", NF, NF, 3)
    for g in range(4):
        m = f"adapt.{g}.mask"
        _conv(spec, m + ".0", 16, NF, 3)
        _bn(spec, m + ".1", 16)
        _conv(spec, m + ".4", 16, 16, 3)
        _bn(spec, m + ".5", 16)
        _conv(spec, m + ".7", 16, 16, 3)
        _bn(spec, m + ".8", 16)
        _conv(spec, m + ".11", 1, 16, 3)
        _bn(spec, m + ".12", 1)
        _osconv(spec, f"adapt.{g}.adapt", NF, NF)
    _conv(spec, "conv_last", NF, NF, 3)
    spec.append(("upsample.weight_compress", (4, NF // 8, NF, 1, 1), "expert"))
    spec.append(("upsample.weight_expand", (4, NF, NF // 8, 1, 1), "expert"))
    _conv(spec, "upsample.kernel_conv.0", NF * 25, NF, 1)
    _conv(spec, "upsample.body.0", 64, 4, 1)
    _conv(spec, "upsample.body.2", 64, 64, 1)
    _conv(spec, "upsample.routing.0", 4, 64, 1)
    _conv(spec, "upsample.offset", 2, 64, 1)
    _conv(spec, "upsample.st_offset", 2, 64, 1)
    _conv(spec, "upsample.fusion", NF, 2 * NF, 1)
    _conv(spec, "tail", 3, NF, 3)
    return spec


def _gen(seed: int, name: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:7], "little"))


def make_state_dict(seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 state_dict.  Weight std = gain / sqrt(3 * fan_in) (the contractive
    PyTorch-default scale), biases and all BN statistics perturbed away from their defaults."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape, kind in state_dict_spec():
        g = _gen(seed, name)
        if kind == "bn_n":
            sd[name] = torch.tensor(100, dtype=torch.long)
            continue
        if kind in ("w", "lin_w", "expert"):
            fan_in = int(torch.Size(shape[1:]).numel()) if kind != "expert" else int(torch.Size(shape[2:]).numel())
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(3.0 * fan_in))
        elif kind == "bank":
            fan_in = shape[2] * 9
            t = torch.randn(shape, generator=g) * (gain * math.sqrt(2.0 / fan_in) * 0.6)
        elif kind == "b":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "bn_w":
            t = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif kind == "bn_b":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_rm":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_rv":
            t = 0.5 + torch.rand(shape, generator=g)
        elif kind == "gamma":
            t = torch.full(shape, 0.8)
        else:
            raise AssertionError(kind)
        sd[name] = t.float().contiguous()
    return sd


def make_input(b: int, h: int, w: int, seed: int = 1234, t: int = 7) -> torch.Tensor:
    """Synthetic LR window, U[0,1) fp32 (seed 1234 + clip id)."""
    return torch.rand(b, t, 3, h, w, generator=torch.Generator().manual_seed(seed))
