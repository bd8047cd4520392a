"""Train-mode forward of SAVSR with the 3x3 convolutions differentiable on the tcgen05 kernels.

Staged design: every 3x3 convolution with 64-multiple channel counts -- 98 % of the FLOPs of a training
step, forward, data gradient and weight gradient -- runs through ``savsr_b200.autograd.conv3x3``; OSA-Conv's per-sample folded
kernels go through the same op, so its gradient reaches the weight bank, the four attention heads, scale_routing and the pooled
mean (savsr_arch.py:139-172).  The small glue of the graph (1x1 convs, the 3/6->64 first layer, the 64->16->1 mask net, BatchNorm in
train mode with batch statistics, channel attention, grid_sample, the SATU MLP) is recorded on the autograd tape with ATen ops on
the device: functionally complete, not yet hand-written.  The step mirrors lbasicsr/models/sr_model.py:101-128 (fp32 master
weights, Charbonnier loss basic_loss.py:22-24, Adam, EMA base_model.py:75-82).

Written as a function of the module's named parameters / buffers (same names as the reference's state_dict).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import autograd as _A
from .autograd import conv3x3
from .engine import get_hw, normalize_scale

Tensor = torch.Tensor
BN_MOMENTUM, BN_EPS = 0.1, 1e-5


class _Net:
    """Parameters, buffers and mode of one forward."""

    def __init__(self, module: torch.nn.Module, training: bool):
        self.P: Dict[str, Tensor] = dict(module.named_parameters())
        self.B: Dict[str, Tensor] = dict(module.named_buffers())
        self.training = training

    def has(self, name: str) -> bool:
        return name in self.P

    def count(self, pattern: str) -> int:
        n = 0
        while pattern.format(n) in self.P:
            n += 1
        return n

    # ---- layers
    def conv(self, name: str, x: Tensor) -> Tensor:
        """nn.Conv2d(stride 1, padding k // 2).  3x3 with 64-multiple channels -> tensor-core op (64 output channels per call)."""
        w, b = self.P[name + ".weight"], self.P.get(name + ".bias")
        co, ci, k, _ = w.shape
        if k == 3 and ci % 64 == 0 and co % 64 == 0:
            outs = [conv3x3(x, w[o:o + 64], b[o:o + 64] if b is not None else None) for o in range(0, co, 64)]
            return outs[0] if len(outs) == 1 else torch.cat(outs, 1)
        return F.conv2d(x, w, b, 1, k // 2)

    def bn(self, name: str, x: Tensor) -> Tensor:
        """nn.BatchNorm2d: batch statistics + running-stat update in train mode (savsr_arch.py:26, 191-204), affine with running stats in eval."""
        if self.training and name + ".num_batches_tracked" in self.B:
            self.B[name + ".num_batches_tracked"].add_(1)
        return F.batch_norm(x, self.B[name + ".running_mean"], self.B[name + ".running_var"], self.P[name + ".weight"],
                            self.P[name + ".bias"], self.training, BN_MOMENTUM, BN_EPS)


def _lrelu(x: Tensor, slope: float = 0.2) -> Tensor:
    return F.leaky_relu(x, slope)


# ------------------------------------------------------------------------------------------------ OSA-Conv (savsr_arch.py:139-172)
def osa_fold(net: _Net, prefix: str, pooled: Tensor, scale) -> Tensor:
    """pooled [b, ci] = mean(x) -> the per-sample folded kernels [b, co, ci, 3, 3] of OSA-Conv: scale_routing (savsr_arch.py:123-128),
    ScaleAttention with train-mode BatchNorm (91-96), bank mix and the four attentions folded into the kernel (158-163)."""
    s = normalize_scale(scale)
    b, ci = pooled.shape
    P = net.P
    info = torch.cat([pooled.new_full((b, 1), 1.0 / s[0]), pooled.new_full((b, 1), 1.0 / s[1])], 1)   # (1/s_h, 1/s_w), savsr_arch.py:143-145; fill kernels: graph-capturable
    v = torch.cat([info, pooled], 1)
    v = F.relu(F.linear(v, P[prefix + ".scale_routing.0.weight"], P[prefix + ".scale_routing.0.bias"]))
    v = F.relu(F.linear(v, P[prefix + ".scale_routing.2.weight"], P[prefix + ".scale_routing.2.bias"]))
    a = prefix + ".attention"
    z = F.relu(net.bn(a + ".bn", F.conv2d(v.view(b, ci, 1, 1), P[a + ".fc.weight"])))
    ca = torch.sigmoid(net.conv(a + ".channel_fc", z)).flatten(1)
    fa = torch.sigmoid(net.conv(a + ".filter_fc", z)).flatten(1)
    sa = torch.sigmoid(net.conv(a + ".spatial_fc", z)).flatten(1)
    ka = torch.softmax(net.conv(a + ".kernel_fc", z).flatten(1), 1)
    bank = P[prefix + ".weight"]                                                          # [8, 64, ci, 3, 3]
    co = bank.shape[1]
    w = torch.einsum("bk,koiuv->boiuv", ka, bank)
    return w * sa.view(b, 1, 1, 3, 3) * ca.view(b, 1, ci, 1, 1) * fa.view(b, co, 1, 1, 1)  # all four attentions folded into the kernel


def osconv(net: _Net, prefix: str, x: Tensor, scale) -> Tensor:
    return conv3x3(x, osa_fold(net, prefix, x.mean(dim=(2, 3)), scale))                   # per-sample kernels, no bias (savsr_arch.py:166)


# ------------------------------------------------------------------------------------------------ trunk blocks
def residual_block(net: _Net, prefix: str, xs: List[Tensor], scale) -> List[Tensor]:
    """savsr_arch.py:399-415"""
    n = len(xs)
    x1 = [_lrelu(net.conv(f"{prefix}.conv0.{i}", xs[i])) for i in range(n)]
    merged = torch.cat(x1, 1)
    if net.has(prefix + ".osconv.weight"):
        base = _lrelu(osconv(net, prefix + ".osconv", merged, scale))
    else:
        base = _lrelu(net.conv(prefix + ".conv1", merged))
    return [xs[i] + _lrelu(net.conv(f"{prefix}.conv2.{i}", torch.cat([base, x1[i]], 1))) for i in range(n)]


def window_unit_l1(net: _Net, prefix: str, frames: Tensor, h_past: Tensor, scale) -> Tensor:
    """savsr_arch.py:444-464; frames [b, 3, c, h, w] = (previous, centre, next)"""
    h_sup = _lrelu(net.conv(prefix + ".conv_sup", torch.cat([frames[:, 0], frames[:, 2]], 1)))
    h_c = _lrelu(net.conv(prefix + ".conv_c", frames[:, 1]))
    feats = [h_c, h_sup, h_past]
    for j in range(net.count(prefix + ".blocks.{}.conv0.0.weight")):
        feats = residual_block(net, f"{prefix}.blocks.{j}", feats, scale)
    return net.conv(prefix + ".merge", torch.cat(feats, 1))


def window_unit_l2(net: _Net, prefix: str, xs: List[Tensor], scale) -> Tensor:
    """savsr_arch.py:485-501 for the shipped configuration (5 streams, one output)"""
    f = [_lrelu(net.conv(f"{prefix}.conv_h.{i}", xs[i])) for i in range(len(xs))]
    for j in range(net.count(prefix + ".blocks.{}.conv0.0.weight")):
        f = residual_block(net, f"{prefix}.blocks.{j}", f, scale)
    return net.conv(prefix + ".merge", torch.cat(f, 1))


def rcab(net: _Net, prefix: str, x: Tensor) -> Tensor:
    """savsr_arch.py:504-549"""
    t = net.conv(prefix + ".rcab.2", F.relu(net.conv(prefix + ".rcab.0", x)))
    y = t.mean(dim=(2, 3), keepdim=True)
    y = torch.sigmoid(net.conv(prefix + ".rcab.3.attention.3", F.relu(net.conv(prefix + ".rcab.3.attention.1", y))))
    return t * y + x


def residual_group(net: _Net, prefix: str, x: Tensor) -> Tensor:
    """savsr_arch.py:552-571"""
    t = x
    for j in range(net.count(prefix + ".residual_group.{}.rcab.0.weight")):
        t = rcab(net, f"{prefix}.residual_group.{j}", t)
    return net.conv(prefix + ".conv", t) + x


def osadapt_mask(net: _Net, prefix: str, x: Tensor) -> Tensor:
    """The mask net of OSAdapt (savsr_arch.py:189-206) -> [b, 1, h, w]"""
    m = prefix + ".mask"
    t = F.relu(net.bn(m + ".1", net.conv(m + ".0", x)))
    t = F.avg_pool2d(t, 2)
    t = F.relu(net.bn(m + ".5", net.conv(m + ".4", t)))
    t = F.relu(net.bn(m + ".8", net.conv(m + ".7", t)))
    t = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    return torch.sigmoid(net.bn(m + ".12", net.conv(m + ".11", t)))


def osadapt(net: _Net, prefix: str, x: Tensor, scale) -> Tensor:
    """savsr_arch.py:186-214"""
    return x + osconv(net, prefix + ".adapt", x, scale) * osadapt_mask(net, prefix, x)


# ------------------------------------------------------------------------------------------------ SATU (savsr_arch.py:315-376)
def _rel_coord(n_out: int, s: float, device) -> Tensor:
    q = (torch.arange(n_out, dtype=torch.float32, device=device) + 0.5) / s
    return q - torch.floor(q + 1e-3) - 0.5


def satu(net: _Net, prefix: str, x: Tensor, scale, st_feat: Tensor) -> Tensor:
    s = normalize_scale(scale)
    b, c, h, w = x.shape
    H, W = get_hw(h, w, s)
    dev = x.device
    sta = _A.sta_lrelu(x, net.conv(prefix + ".kernel_conv.0", st_feat), 0.1)                               # kernel_conv's LeakyReLU + sta_conv (297-313), fused natively
    ry = _rel_coord(H, s[0], dev).view(H, 1).expand(H, W)
    rx = _rel_coord(W, s[1], dev).view(1, W).expand(H, W)
    inp = torch.stack([torch.full((H, W), 1.0 / s[1], device=dev), torch.full((H, W), 1.0 / s[0], device=dev), ry, rx], 0)[None]
    e = F.relu(net.conv(prefix + ".body.2", F.relu(net.conv(prefix + ".body.0", inp))))
    off, st_off = net.conv(prefix + ".offset", e), net.conv(prefix + ".st_offset", e)
    r = torch.sigmoid(net.conv(prefix + ".routing.0", e))[0]                                               # [4, H, W]

    def gather(t: Tensor, o: Tensor) -> Tensor:                                                            # 262-295
        gx = ((torch.arange(W, dtype=torch.float32, device=dev) + 0.5) / s[1] - 0.5) * 2 / (w - 1) - 1
        gy = ((torch.arange(H, dtype=torch.float32, device=dev) + 0.5) / s[0] - 0.5) * 2 / (h - 1) - 1
        grid = torch.stack([gx.view(1, 1, W).expand(1, H, W) + o[:, 0] * 2 / (w - 1),
                            gy.view(1, H, 1).expand(1, H, W) + o[:, 1] * 2 / (h - 1)], -1)
        return F.grid_sample(t, grid.expand(b, -1, -1, -1), mode="bilinear", padding_mode="zeros", align_corners=True)
    fea0 = gather(x, off)
    wc = net.P[prefix + ".weight_compress"].flatten(2)                                                     # [4, 8, 64]
    we = net.P[prefix + ".weight_expand"].flatten(2)                                                       # [4, 64, 8]
    # two-stage routed mix, 353-370, as two 1x1 convolutions over the 32 = 4 experts x 8 compressed channels (the literal form
    # materialises a [b, 4, 64, H, W] tensor: 268 MB at 4 x 256 x 256)
    u = F.conv2d(fea0, wc.reshape(32, c, 1, 1))                                                            # rows e * 8 + k
    t = (u.view(b, 4, 8, H, W) * r[None, :, None]).sum(1)                                                  # t_k = sum_e r_e (Wc_e fea0)_k
    v = (r[None, :, None] * t[:, None]).reshape(b, 32, H, W)                                               # v[e * 8 + k] = r_e t_k
    fea = F.conv2d(v, we.permute(1, 0, 2).reshape(c, 32, 1, 1)) + fea0                                     # sum_e r_e We_e t + fea0
    return net.conv(prefix + ".fusion", torch.cat([gather(sta, st_off), fea], 1))


# ------------------------------------------------------------------------------------------------ whole forward (savsr_arch.py:692-742)
def forward(module: torch.nn.Module, x: Tensor, scale, training: Optional[bool] = None) -> Tensor:
    """x [b, 7, 3, h, w] fp32 on a CUDA device -> [b, 3, H, W], differentiable with respect to every parameter."""
    if not x.is_cuda:
        raise RuntimeError("savsr_b200.train runs on CUDA (sm_100a) only; there is no CPU fallback")
    net = _Net(module, module.training if training is None else training)
    scale = normalize_scale(scale)
    b, t, c, h, w = x.shape
    H, W = get_hw(h, w, scale)
    xc = x[:, t // 2]
    ph, pw = h & 1, w & 1
    xp = F.pad(x.reshape(-1, c, h, w), [0, pw, 0, ph], mode="reflect").view(b, t, c, h + ph, w + pw) if (ph or pw) else x
    hf = hb = x.new_zeros(b, 64, h + ph, w + pw)
    f2p: List[Tensor] = []
    p2f: List[Tensor] = []
    for idx in range(t - 2):                                                                               # 708-719
        ct = t - 2 - idx
        hf = window_unit_l1(net, "f2p_win", xp[:, ct - 1:ct + 2], hf, scale)
        f2p.insert(0, hf)
        ct = idx + 1
        hb = window_unit_l1(net, "p2f_win", xp[:, ct - 1:ct + 2], hb, scale)
        p2f.append(hb)
    feats = [torch.cat([f2p[i], p2f[i]], 1) for i in range(t - 2)]
    share = _lrelu(net.conv("h_win_conv_h", window_unit_l2(net, "h_win.0", feats, scale)))
    y = share
    for i in range(net.count("RG.{}.conv.weight")):                                                        # 727-732
        y = osadapt(net, f"adapt.{i}", residual_group(net, f"RG.{i}", y), scale) + net.P["gamma"] * share
    y = net.conv("conv_last", y) + share
    sr = net.conv("tail", satu(net, "upsample", y[..., :h, :w], scale, share[..., :h, :w]))
    return sr + F.interpolate(xc, size=(H, W), mode="bilinear", align_corners=False)


# ------------------------------------------------------------------------------------------------ the training step
def charbonnier(pred: Tensor, target: Tensor, eps: float = 1e-12) -> Tensor:
    """CharbonnierLoss(loss_weight 1, reduction mean): mean(sqrt((pred - target)^2 + eps)), lbasicsr/losses/basic_loss.py:22-24, 83-114."""
    return torch.sqrt((pred - target) ** 2 + eps).mean()


class Trainer:
    """One optimisation step as lbasicsr/models/sr_model.py:101-128 + asvsr_model.py:21-29 run it: set the batch's scale, forward,
    Charbonnier, backward, Adam (train YAML: lr 2e-4, betas 0.9 / 0.99), EMA of the weights (base_model.py:75-82, decay 0.999).
    `net` may be wrapped in DistributedDataParallel by the caller (base_model.py:98-99): gradients then all-reduce over NCCL.

    use_graph: capture the whole step (forward, backward, Adam, EMA) into one CUDA graph per (scale, batch shape) and replay it.
    At the training shape (4 x 7 x 3 x 64 x 64 per GPU) a step is ~20 000 small kernels and purely launch-bound when issued from
    Python; the graph removes that.  Single-process only (not combined with DistributedDataParallel here)."""

    def __init__(self, net: torch.nn.Module, lr: float = 2e-4, betas=(0.9, 0.99), ema_decay: float = 0.999, use_graph: bool = False):
        self.net = net
        self.core = net.module if hasattr(net, "module") else net
        if use_graph and self.core is not net:
            raise ValueError("use_graph is for a single process; with DistributedDataParallel run the eager step")
        self.use_graph = use_graph
        self.opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=lr, betas=betas, capturable=use_graph)
        self.ema_decay = ema_decay
        self.ema = {k: v.detach().clone() for k, v in self.core.named_parameters()} if ema_decay > 0 else None
        self._graphs = {}

    def _step_body(self, lq: Tensor, gt: Tensor) -> Tensor:
        loss = charbonnier(self.net(lq), gt)           # the module's train-mode forward: native launch list unless core.native_training is False
        loss.backward()
        self.opt.step()
        if self.ema is not None:
            with torch.no_grad():
                for k, v in self.core.named_parameters():
                    self.ema[k].mul_(self.ema_decay).add_(v.detach(), alpha=1 - self.ema_decay)
        return loss.detach()

    def step(self, lq: Tensor, gt: Tensor, scale) -> Tensor:
        self.core.set_scale(scale)
        self.net.train()
        if not self.use_graph:
            self.opt.zero_grad(set_to_none=True)
            return self._step_body(lq, gt)
        key = (tuple(normalize_scale(scale)), tuple(lq.shape), tuple(gt.shape))
        ent = self._graphs.get(key)
        if ent is None:
            s_lq, s_gt = lq.clone(), gt.clone()
            side = torch.cuda.Stream(lq.device)
            side.wait_stream(torch.cuda.current_stream(lq.device))
            with torch.cuda.stream(side):                       # warm-up outside the graph: kernel attributes, optimizer state, cuDNN plans
                for _ in range(2):
                    self.opt.zero_grad(set_to_none=True)
                    self._step_body(s_lq, s_gt)
            torch.cuda.current_stream(lq.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(g, stream=torch.cuda.Stream(lq.device)):        # capture stream on the data's device (cf. engine.Plan.capture)
                s_loss = self._step_body(s_lq, s_gt)
            ent = self._graphs[key] = (g, s_lq, s_gt, s_loss)
        g, s_lq, s_gt, s_loss = ent
        s_lq.copy_(lq, non_blocking=True)
        s_gt.copy_(gt, non_blocking=True)
        g.replay()
        return s_loss
