"""Parity checks of every libsavsr_sm100 entry point against fp32 PyTorch / the CPU oracle.

Shared by ``tests/test_gpu_*.py`` (pytest, ``-m gpu``) and ``scripts/gpu_bringup.py`` (one subprocess per
check, so a trapping kernel cannot poison the other checks).  Every check calls through the C ABI.
Each function returns a dict of error metrics and raises AssertionError when out of tolerance.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from savsr_b200 import _capi as K          # noqa: E402
from savsr_b200 import engine              # noqa: E402

DEV = torch.device("cuda", 0)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ctx() -> K.Context:
    return engine.context(0)


def set_precision(name: str) -> None:
    """Switch the context's 16-bit operand format ("bf16" / "fp16") for the kernel-level checks."""
    ctx().set_format(K.FMT_NAMES[name])


def _h16() -> torch.dtype:
    return torch.float16 if K.load().savsr_ctx_get_format(ctx().handle) == K.FMT_FP16 else torch.bfloat16


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round to the context's 16-bit operand format (bf16 by default) and back to fp32."""
    return t.to(_h16()).to(torch.float32)


class ArenaBox:
    def __init__(self, nslots: int, batch: int, height: int, width: int):
        self.t = torch.zeros(nslots * batch, height, width, 64, dtype=_h16(), device=DEV)
        self.a = K.Arena(ctx(), self.t.data_ptr(), nslots, batch, height, width)
        self.B, self.H, self.W = batch, height, width

    def put(self, slot: int, nchw: torch.Tensor) -> None:
        nchw = nchw.to(DEV, torch.float32).contiguous()
        K.check(K.load().savsr_arena_import(self.a.handle, slot, nchw.data_ptr(), _stream()))

    def get(self, slot: int) -> torch.Tensor:
        out = torch.empty(self.B, 64, self.H, self.W, device=DEV)
        K.check(K.load().savsr_arena_export(self.a.handle, slot, out.data_ptr(), _stream()))
        torch.cuda.synchronize()
        return out


def quad_row(n):
    """SAVSR_ROWS_QUAD: packed row n of an N = 64 block holds channel quad_row(n) (bit fields [2:1] and [4:3] swapped)."""
    return (n & 0x21) | (((n >> 1) & 3) << 3) | (((n >> 3) & 3) << 1)


def pack_weight(w: torch.Tensor, n_tile: int = 64, co_pad=None, rows: int = K.ROWS_QUAD) -> torch.Tensor:
    lib = K.load()
    w = w.to(DEV, torch.float32).contiguous()
    co_real, ci, ks, _ = w.shape
    co = co_pad or co_real
    out = torch.empty(lib.savsr_packed_weight_bytes(co, ci, ks), dtype=torch.uint8, device=DEV)
    K.check(lib.savsr_pack_conv_weight(w.data_ptr(), co_real, co, ci, ks, n_tile, lib.savsr_ctx_get_format(ctx().handle), rows,
                                       out.data_ptr(), _stream()))
    return out


def unpack_weight(packed: torch.Tensor, co: int, ci: int, ks: int, n_tile: int = 64, rows: int = K.ROWS_QUAD) -> torch.Tensor:
    """Inverse of the packed layout (python restatement of the swizzle and row order) -> fp32 [co][ci][ks][ks]."""
    v = packed.view(_h16()).float().cpu().numpy()
    taps = ks * ks
    nkb = (ci // 64) * taps
    v = v.reshape(co // n_tile, nkb, n_tile * 64)
    n = np.arange(n_tile)[:, None]
    k = np.arange(64)[None, :]
    pos = n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7))
    blk = v[:, :, pos]                                 # [ng, kb, n, k]
    if n_tile == 64 and rows == K.ROWS_QUAD:
        blk = blk[:, :, quad_row(np.arange(64)), :]    # channel o sits in row quad_row(o) (involution)
    blk = blk.reshape(co // n_tile, ci // 64, taps, n_tile, 64)
    w = blk.transpose(0, 3, 1, 4, 2).reshape(co, ci, ks, ks)
    return torch.from_numpy(np.ascontiguousarray(w))


def group(src, dst, weight, bias=None, act=K.ACT_NONE, slope=0.2, res1=-1, res2=-1, res2_scale=0.0, wstride=0,
          mask=None, pool=None, aux=None) -> K.ConvGroup:
    g = K.ConvGroup()
    for i, s in enumerate(src):
        g.src_slot[i] = s
    g.nsrc = len(src)
    g.dst_slot, g.res1_slot, g.res2_slot, g.res2_scale = dst, res1, res2, res2_scale
    g.act, g.slope = act, slope
    g.weight = weight.data_ptr()
    g.weight_sample_stride = wstride
    g.bias = bias.data_ptr() if bias is not None else None
    g.mask = mask.data_ptr() if mask is not None else None
    g.pool = pool.data_ptr() if pool is not None else None
    g.aux_dst = aux.data_ptr() if aux is not None else None
    return g


def run_conv(ab: ArenaBox, groups, ksize=3, n_tile=64, dst_mode=K.DST_ARENA, impl=K.IMPL_TAP):
    arr = (K.ConvGroup * len(groups))(*groups)
    K.check(K.load().savsr_conv(ctx().handle, ab.a.handle, arr, len(groups), ksize, n_tile, dst_mode, impl, _stream()))
    torch.cuda.synchronize()


def _tol(ref: torch.Tensor, rel=2.0 ** -7, abs_=2e-3):
    return rel * ref.abs() + abs_


def assert_close(name, got, ref, rel=2.0 ** -7, abs_=2e-3):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    bad = err > _tol(ref, rel, abs_)
    info = dict(max_abs=float(err.max()), ref_absmax=float(ref.abs().max()), bad=int(bad.sum()), numel=ref.numel())
    assert not bool(bad.any()), f"{name}: {info}"
    return info


# ------------------------------------------------------------------------------------------------ convolution
def check_conv(impl=K.IMPL_TAP, ksize=3, nsrc=1, B=1, H=32, W=24, full_epilogue=True, seed=0):
    """savsr_conv (N = 64) vs F.conv2d on the same bf16-rounded operands, with the whole epilogue menu."""
    torch.manual_seed(seed)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ngroups = 2
    nslots = ngroups * (nsrc + 3) + 1
    ab = ArenaBox(nslots, B, H, W)
    tiles = ab.a.tiles
    results = {}
    groups, refs, pools = [], [], []
    slot = 0
    for gi in range(ngroups):
        xs = [bf16_round(torch.randn(B, 64, H, W, device=DEV)) for _ in range(nsrc)]
        src = []
        for x in xs:
            ab.put(slot, x); src.append(slot); slot += 1
        w = bf16_round(torch.randn(64, 64 * nsrc, ksize, ksize, device=DEV) * (1.0 / (8.0 * ksize * nsrc ** 0.5)))
        bias = torch.randn(64, device=DEV) * 0.1
        r1 = bf16_round(torch.randn(B, 64, H, W, device=DEV))
        r2 = bf16_round(torch.randn(B, 64, H, W, device=DEV))
        mask = torch.rand(B, H * W, device=DEV)
        ab.put(slot, r1); s_r1 = slot; slot += 1
        ab.put(slot, r2); s_r2 = slot; slot += 1
        dst = slot; slot += 1
        pool = torch.full((B, tiles * 4, 64), float("nan"), device=DEV)
        wp = pack_weight(w)
        ref = F.conv2d(torch.cat(xs, 1), w, bias, padding=ksize // 2)
        if full_epilogue and gi == 0:
            ref = F.leaky_relu(ref, 0.2) * mask.view(B, 1, H, W) + r1 + 0.75 * r2
            g = group(src, dst, wp, bias, act=K.ACT_LRELU, slope=0.2, res1=s_r1, res2=s_r2, res2_scale=0.75, mask=mask, pool=pool)
        elif full_epilogue:
            ref = F.relu(ref)
            g = group(src, dst, wp, bias, act=K.ACT_RELU, pool=pool)
        else:
            g = group(src, dst, wp, bias)
        groups.append(g); refs.append((dst, ref, pool if full_epilogue else None)); pools.append((wp, bias, mask, pool))
    run_conv(ab, groups, ksize=ksize, impl=impl)
    for gi, (dst, ref, pool) in enumerate(refs):
        results[f"g{gi}"] = assert_close(f"conv g{gi}", ab.get(dst), ref)
        if pool is not None:
            psum = pool.sum(1)                                   # [B, 64]
            results[f"pool{gi}"] = assert_close(f"pool g{gi}", psum, ref.sum((2, 3)), rel=2e-3, abs_=0.05)
    return results


def check_conv_repeatability(nsrc=2, ngroups=3, B=3, H=144, W=180, reps=25, seed=21):
    """Back-to-back launches at the benchmark shape must be bit-identical (guards the multi-warp mbarrier protocol
    of the batched dual-issuer kernel against races: a lapped or aliased barrier phase shows up as a mismatch or a trap)."""
    torch.manual_seed(seed)
    ab = ArenaBox(ngroups * (nsrc + 1), B, H, W)
    ab.t.normal_()
    groups, keep = [], []
    for g in range(ngroups):
        w = pack_weight(torch.randn(64, 64 * nsrc, 3, 3, device=DEV) * 0.05)
        bias = torch.randn(64, device=DEV)
        keep += [w, bias]
        groups.append(group([g * (nsrc + 1) + i for i in range(nsrc)], g * (nsrc + 1) + nsrc, w, bias, act=K.ACT_LRELU))
    run_conv(ab, groups, impl=K.IMPL_HALO)
    dst = [g * (nsrc + 1) + nsrc for g in range(ngroups)]
    first = [ab.t[d * B:(d + 1) * B].clone() for d in dst]
    for _ in range(reps):
        for d in dst:
            ab.t[d * B:(d + 1) * B].zero_()
        run_conv(ab, groups, impl=K.IMPL_HALO)
        for d, f in zip(dst, first):
            assert torch.equal(ab.t[d * B:(d + 1) * B], f), "conv output changed between identical launches"
    chk = [group(g_.src_slot[:nsrc], dst[i], keep[2 * i], keep[2 * i + 1], act=K.ACT_LRELU) for i, g_ in enumerate(groups)]
    run_conv(ab, chk, impl=K.IMPL_CHECK)
    for d, f in zip(dst, first):
        chk_out = ab.t[d * B:(d + 1) * B].float()
        bad = (chk_out - f.float()).abs() > 2.0 ** -7 * f.float().abs() + 1e-3      # one bf16 ulp: accumulation order differs
        assert not bool(bad.any()), f"tensor-core kernel differs from the checker kernel in {int(bad.sum())} values"
    return dict(reps=reps, status="bit-identical")


def check_conv_aux16(impl=K.IMPL_TAP, B=2, H=20, W=28, seed=1):
    """N = 16 variant writing the fp32 [B][H*W][16] side buffer (OSAdapt mask conv, savsr_arch.py:190-192)."""
    torch.manual_seed(seed)
    ab = ArenaBox(2, B, H, W)
    x = bf16_round(torch.randn(B, 64, H, W, device=DEV))
    ab.put(0, x)
    w = bf16_round(torch.randn(16, 64, 3, 3, device=DEV) * 0.05)
    bias = torch.randn(16, device=DEV) * 0.1
    aux = torch.full((B, H * W, 16), float("nan"), device=DEV)
    run_conv(ab, [group([0], 0, pack_weight(w, n_tile=16), bias, act=K.ACT_RELU, aux=aux)], n_tile=16, dst_mode=K.DST_AUX16, impl=impl)
    ref = F.relu(F.conv2d(x, w, bias, padding=1)).permute(0, 2, 3, 1).reshape(B, H * W, 16)
    return dict(aux16=assert_close("conv aux16", aux, ref, rel=1e-3, abs_=1e-3))


def check_osa_conv_per_sample(impl=K.IMPL_TAP, B=3, H=18, W=26, seed=3):
    """Per-sample weights (weight_sample_stride): the groups=batch conv of savsr_arch.py:166."""
    torch.manual_seed(seed)
    ab = ArenaBox(4, B, H, W)
    xs = [bf16_round(torch.randn(B, 64, H, W, device=DEV)) for _ in range(3)]
    for i, x in enumerate(xs):
        ab.put(i, x)
    w = bf16_round(torch.randn(B, 64, 192, 3, 3, device=DEV) * 0.03)
    packs = torch.stack([pack_weight(w[n]) for n in range(B)])
    run_conv(ab, [group([0, 1, 2], 3, packs, wstride=packs.stride(0))], impl=impl)
    xin = torch.cat(xs, 1)
    ref = torch.cat([F.conv2d(xin[n:n + 1], w[n], None, padding=1) for n in range(B)])
    return dict(per_sample=assert_close("per-sample conv", ab.get(3), ref))


def check_pack_roundtrip():
    torch.manual_seed(4)
    out = {}
    for (co, ci, ks, nt) in ((64, 192, 3, 64), (128, 320, 3, 64), (64, 192, 1, 64), (16, 64, 3, 16), (1600, 64, 1, 64)):
        w = torch.randn(co, ci, ks, ks)
        for rows in (K.ROWS_QUAD, K.ROWS_LINEAR):
            back = unpack_weight(pack_weight(w, n_tile=nt, rows=rows), co, ci, ks, nt, rows=rows)
            assert torch.equal(back, bf16_round(w)), (co, ci, ks, nt, rows)
        out[f"{co}x{ci}x{ks}"] = "exact"
    return out


# ------------------------------------------------------------------------------------------------ small ops
def check_pack_frames(B=2, h=13, w=15, seed=11):
    """savsr_pack_frames + zero-expanded filters on the tensor-core conv == conv_c / conv_sup of the reference."""
    torch.manual_seed(seed)
    hp, wp = h + (h & 1), w + (w & 1)
    ab = ArenaBox(3, B, hp, wp)
    x = torch.rand(B, 7, 3, h, w, device=DEV)
    K.check(K.load().savsr_pack_frames(ctx().handle, ab.a.handle, x.data_ptr(), 7, h, w, 0, _stream()))
    xp = F.pad(x.reshape(-1, 3, h, w), [0, wp - w, 0, hp - h], mode="reflect").view(B, 7, 3, hp, wp) if (hp != h or wp != w) else x
    got = ab.get(0)
    ref = torch.zeros(B, 64, hp, wp, device=DEV)
    ref[:, :21] = xp.reshape(B, 21, hp, wp)
    res = dict(frames=assert_close("pack_frames", got, bf16_round(ref), rel=0, abs_=0))
    wc = torch.randn(64, 3, 3, 3, device=DEV) * 0.2
    ws = torch.randn(64, 6, 3, 3, device=DEV) * 0.2
    bc = torch.randn(64, device=DEV) * 0.1
    c = 4
    w64c = torch.zeros(64, 64, 3, 3, device=DEV); w64c[:, 3 * c:3 * c + 3] = wc
    w64s = torch.zeros(64, 64, 3, 3, device=DEV); w64s[:, 3 * (c - 1):3 * c] = ws[:, :3]; w64s[:, 3 * (c + 1):3 * (c + 2)] = ws[:, 3:]
    pc, ps = pack_weight(w64c), pack_weight(w64s)          # keep the packed tensors alive across the launch
    run_conv(ab, [group([0], 1, pc, bc, act=K.ACT_LRELU), group([0], 2, ps, bc, act=K.ACT_LRELU)], impl=K.IMPL_HALO)
    xb = bf16_round(xp)
    ref_c = F.leaky_relu(F.conv2d(xb[:, c], bf16_round(wc), bc, padding=1), 0.2)
    ref_s = F.leaky_relu(F.conv2d(torch.cat([xb[:, c - 1], xb[:, c + 1]], 1), bf16_round(ws), bc, padding=1), 0.2)
    res["conv_c"] = assert_close("conv_c via frames slot", ab.get(1), ref_c)
    res["conv_sup"] = assert_close("conv_sup via frames slot", ab.get(2), ref_s)
    return res


def _osa_state(ci, seed):
    g = torch.Generator().manual_seed(seed)
    att = max(int(ci * 0.0625), 16)
    r = lambda *s, k=1.0: torch.randn(*s, generator=g) * k
    sd = {
        "o.weight": r(8, 64, ci, 3, 3, k=0.05),
        "o.scale_routing.0.weight": r(2 * ci, ci + 2, k=(ci + 2) ** -0.5), "o.scale_routing.0.bias": r(2 * ci, k=0.1),
        "o.scale_routing.2.weight": r(ci, 2 * ci, k=(2 * ci) ** -0.5), "o.scale_routing.2.bias": r(ci, k=0.1),
        "o.attention.fc.weight": r(att, ci, 1, 1, k=ci ** -0.5),
        "o.attention.bn.weight": 0.5 + torch.rand(att, generator=g), "o.attention.bn.bias": r(att, k=0.1),
        "o.attention.bn.running_mean": r(att, k=0.1), "o.attention.bn.running_var": 0.5 + torch.rand(att, generator=g),
        "o.attention.channel_fc.weight": r(ci, att, 1, 1, k=0.3), "o.attention.channel_fc.bias": r(ci, k=0.1),
        "o.attention.filter_fc.weight": r(64, att, 1, 1, k=0.3), "o.attention.filter_fc.bias": r(64, k=0.1),
        "o.attention.spatial_fc.weight": r(9, att, 1, 1, k=0.3), "o.attention.spatial_fc.bias": r(9, k=0.1),
        "o.attention.kernel_fc.weight": r(8, att, 1, 1, k=0.5), "o.attention.kernel_fc.bias": r(8, k=0.1),
    }
    return sd, att


def check_osa_prologue(ci=192, B=2, npart=12, npix=300, scale=(1.5, 4), seed=6):
    """pool -> scale_routing -> ScaleAttention -> assembled per-sample kernel vs the oracle
    (savsr_arch.py:143-163, 91-96)."""
    from oracle import savsr_oracle as O
    sd, att = _osa_state(ci, seed)
    dsd = {k: v.to(DEV).contiguous() for k, v in sd.items()}
    nsrc = ci // 64
    g = torch.Generator().manual_seed(seed + 1)
    parts = [torch.randn(B, npart, 64, generator=g).to(DEV) for _ in range(nsrc)]
    pooled = torch.cat([p.sum(1) / npix for p in parts], 1).cpu()          # [B, ci]
    o = K.OsaParams()
    o.ci, o.co, o.att = ci, 64, att
    a = "o.attention"
    bn_s = dsd[a + ".bn.weight"] / torch.sqrt(dsd[a + ".bn.running_var"] + 1e-5)
    bn_b = dsd[a + ".bn.bias"] - dsd[a + ".bn.running_mean"] * bn_s
    o.bank = dsd["o.weight"].data_ptr()
    o.r0_w, o.r0_b = dsd["o.scale_routing.0.weight"].data_ptr(), dsd["o.scale_routing.0.bias"].data_ptr()
    o.r2_w, o.r2_b = dsd["o.scale_routing.2.weight"].data_ptr(), dsd["o.scale_routing.2.bias"].data_ptr()
    o.fc_w = dsd[a + ".fc.weight"].data_ptr()
    o.bn_scale, o.bn_shift = bn_s.data_ptr(), bn_b.data_ptr()
    o.ch_w, o.ch_b = dsd[a + ".channel_fc.weight"].data_ptr(), dsd[a + ".channel_fc.bias"].data_ptr()
    o.fl_w, o.fl_b = dsd[a + ".filter_fc.weight"].data_ptr(), dsd[a + ".filter_fc.bias"].data_ptr()
    o.sp_w, o.sp_b = dsd[a + ".spatial_fc.weight"].data_ptr(), dsd[a + ".spatial_fc.bias"].data_ptr()
    o.kn_w, o.kn_b = dsd[a + ".kernel_fc.weight"].data_ptr(), dsd[a + ".kernel_fc.bias"].data_ptr()
    for i, p in enumerate(parts):
        o.pool[i] = p.data_ptr()
    scratch = torch.zeros(B, 5 * ci + 192, device=DEV)
    packed = torch.zeros(B, 64 * ci * 9 * 2, dtype=torch.uint8, device=DEV)
    o.scratch, o.packed = scratch.data_ptr(), packed.data_ptr()
    arr = (K.OsaParams * 1)(o)
    inv_h = float(np.float32(1.0) / np.float32(scale[0])); inv_w = float(np.float32(1.0) / np.float32(scale[1]))
    K.check(K.load().savsr_osa_prologue(ctx().handle, arr, 1, B, npart, npix, inv_h, inv_w, _stream()))
    torch.cuda.synchronize()
    ca, fa, sa, ka = O.osa_attention(sd, "o", pooled, scale)
    got = scratch[:, 4 * ci + 8: 4 * ci + 8 + ci + 64 + 17].cpu()
    ref = torch.cat([ca, fa, sa, ka], 1)
    res = dict(attention=assert_close("osa attention", got, ref, rel=1e-4, abs_=1e-5))
    wref = O.osa_fold_weight(sd["o.weight"], ca, fa, sa, ka)                 # [B, 64, ci, 3, 3]
    for n in range(B):
        wgot = unpack_weight(packed[n], 64, ci, 3)
        res[f"weight{n}"] = assert_close(f"osa weight {n}", wgot, wref[n], rel=2.0 ** -8, abs_=1e-6)
    return res


def check_ca(B=2, H=20, W=24, seed=7):
    """RCAB channel attention: dst = x + t * sigmoid(W2 relu(W1 mean(t) + b1) + b2) (savsr_arch.py:514-549)."""
    torch.manual_seed(seed)
    ab = ArenaBox(3, B, H, W)
    t = bf16_round(torch.randn(B, 64, H, W, device=DEV)); x = bf16_round(torch.randn(B, 64, H, W, device=DEV))
    ab.put(0, t); ab.put(1, x)
    npart = 7
    parts = torch.randn(B, npart, 64, device=DEV)
    parts = parts - parts.sum(1, keepdim=True) / npart + t.sum((2, 3)).unsqueeze(1) / npart     # partials summing to sum(t)
    w1 = torch.randn(4, 64, device=DEV) * 0.3; b1 = torch.randn(4, device=DEV) * 0.1
    w2 = torch.randn(64, 4, device=DEV) * 0.5; b2 = torch.randn(64, device=DEV) * 0.1
    ysc = torch.empty(B, 64, device=DEV)
    K.check(K.load().savsr_ca_scale_residual(ctx().handle, ab.a.handle, 0, 1, 2, parts.data_ptr(), npart, w1.data_ptr(),
                                             b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), ysc.data_ptr(), _stream()))
    y = torch.sigmoid(F.relu(t.mean((2, 3)) @ w1.t() + b1) @ w2.t() + b2)
    ref = x + t * y.view(B, 64, 1, 1)
    return dict(ca=assert_close("ca_scale_residual", ab.get(2), ref))


def check_mask(B=2, H=12, W=18, seed=8):
    """OSAdapt mask tail: AvgPool2 -> 2 x (conv16+ReLU) -> bilinear x2 -> conv16->1 -> sigmoid (savsr_arch.py:193-205)."""
    torch.manual_seed(seed)
    in16 = torch.relu(torch.randn(B, 16, H, W, device=DEV))
    wa = torch.randn(16, 16, 3, 3, device=DEV) * 0.1; ba = torch.randn(16, device=DEV) * 0.1
    wb = torch.randn(16, 16, 3, 3, device=DEV) * 0.1; bb = torch.randn(16, device=DEV) * 0.1
    wc = torch.randn(1, 16, 3, 3, device=DEV) * 0.2; bc = torch.randn(1, device=DEV) * 0.1
    nhwc = in16.permute(0, 2, 3, 1).reshape(B, H * W, 16).contiguous()
    h0 = torch.empty(B, (H // 2) * (W // 2), 16, device=DEV); h1 = torch.empty_like(h0)
    mask = torch.empty(B, H * W, device=DEV)
    K.check(K.load().savsr_osadapt_mask(ctx().handle, nhwc.data_ptr(), B, H, W, wa.data_ptr(), ba.data_ptr(), wb.data_ptr(),
                                        bb.data_ptr(), wc.data_ptr(), bc.data_ptr(), h0.data_ptr(), h1.data_ptr(), mask.data_ptr(), _stream()))
    t = F.avg_pool2d(in16, 2)
    t = F.relu(F.conv2d(t, wa, ba, padding=1))
    t = F.relu(F.conv2d(t, wb, bb, padding=1))
    t = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    ref = torch.sigmoid(F.conv2d(t, wc, bc, padding=1)).view(B, H * W)
    return dict(mask=assert_close("osadapt mask", mask, ref, rel=1e-4, abs_=1e-5))


# ------------------------------------------------------------------------------------------------ SATU
def _satu_weights(sd):
    d = {k: v.to(DEV).contiguous() for k, v in sd.items() if k.startswith("upsample.")}
    sw = K.SatuWeights()
    sw.body0_w, sw.body0_b = d["upsample.body.0.weight"].data_ptr(), d["upsample.body.0.bias"].data_ptr()
    sw.body2_w, sw.body2_b = d["upsample.body.2.weight"].data_ptr(), d["upsample.body.2.bias"].data_ptr()
    sw.routing_w, sw.routing_b = d["upsample.routing.0.weight"].data_ptr(), d["upsample.routing.0.bias"].data_ptr()
    sw.offset_w, sw.offset_b = d["upsample.offset.weight"].data_ptr(), d["upsample.offset.bias"].data_ptr()
    sw.st_offset_w, sw.st_offset_b = d["upsample.st_offset.weight"].data_ptr(), d["upsample.st_offset.bias"].data_ptr()
    sw.compress, sw.expand = d["upsample.weight_compress"].data_ptr(), d["upsample.weight_expand"].data_ptr()
    return sw, d


def satu_index(h, w, scale, sd=None):
    """Run savsr_satu_index; returns dict of CPU arrays (and the table if weights are given)."""
    H, W = engine.get_hw(h, w, scale)
    s = engine.normalize_scale(scale)
    out = dict(rel_y=torch.empty(H, device=DEV), rel_x=torch.empty(W, device=DEV),
               cell_y=torch.empty(H, dtype=torch.int32, device=DEV), cell_x=torch.empty(W, dtype=torch.int32, device=DEV),
               base_y=torch.empty(H, device=DEV), base_x=torch.empty(W, device=DEV),
               corner_y=torch.empty(H, dtype=torch.int32, device=DEV), corner_x=torch.empty(W, dtype=torch.int32, device=DEV))
    table = None
    sw = keep = None
    if sd is not None:
        sw, keep = _satu_weights(sd)
        table = torch.empty(H * W, 8, device=DEV)
    K.check(K.load().savsr_satu_index(ctx().handle, C.byref(sw) if sw is not None else None, h, w, H, W, float(s[0]), float(s[1]),
                                      out["rel_y"].data_ptr(), out["rel_x"].data_ptr(), out["cell_y"].data_ptr(), out["cell_x"].data_ptr(),
                                      out["base_y"].data_ptr(), out["base_x"].data_ptr(), out["corner_y"].data_ptr(),
                                      out["corner_x"].data_ptr(), table.data_ptr() if table is not None else None, _stream()))
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if table is not None:
        res["table"] = table.cpu()
    return res, H, W


def check_satu_index(h=144, w=180, scale=(1.5, 4)):
    """Bit-exact index vectors vs the numpy oracle (savsr_arch.py:326-333, 270-280)."""
    from oracle import savsr_oracle as O
    res, H, W = satu_index(h, w, scale)
    s = engine.normalize_scale(scale)
    for ax, n_out, n_lr, sc in (("y", H, h, s[0]), ("x", W, w, s[1])):
        assert np.array_equal(res["rel_" + ax].view(np.uint32), O.satu_rel_coord(n_out, sc).view(np.uint32)), f"rel_{ax} not bit-exact"
        assert np.array_equal(res["cell_" + ax], O.satu_cell(n_out, sc)), f"cell_{ax}"
        assert np.array_equal(res["base_" + ax].view(np.uint32), O.satu_base_norm(n_out, n_lr, sc).view(np.uint32)), f"base_{ax} not bit-exact"
        assert np.array_equal(res["corner_" + ax], O.satu_base_corner(n_out, n_lr, sc)), f"corner_{ax}"
    return dict(H=H, W=W, status="bit-exact")


def check_satu_table(h=13, w=15, scale=(2.7, 1.5), seed=1):
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_state_dict
    sd = make_state_dict(seed)
    res, H, W = satu_index(h, w, scale, sd)
    off, st_off, r = O.satu_heads(sd, "upsample", h, w, scale)
    ref = torch.cat([off, st_off, r], 1)[0].permute(1, 2, 0).reshape(H * W, 8)
    return dict(table=assert_close("satu table", res["table"], ref, rel=1e-4, abs_=2e-6))


def check_satu_kconv_sta(B=2, h=13, w=15, seed=19):
    """kernel_conv (1x1 64 -> 1600, LeakyReLU 0.1) + sta_conv in one kernel, kernels consumed from TMEM
    (savsr_arch.py:297-313, 326), vs the oracle's two-step restatement on the same (bf16-rounded) operands."""
    import torch.nn.functional as F
    from oracle import savsr_oracle as O
    torch.manual_seed(seed)
    hp, wp = h + (h & 1), w + (w & 1)
    ab = ArenaBox(3, B, hp, wp)
    a = bf16_round(torch.randn(B, 64, hp, wp, device=DEV))
    x = bf16_round(torch.randn(B, 64, hp, wp, device=DEV))
    wk = bf16_round(torch.randn(1600, 64, 1, 1, device=DEV) * 0.1)           # reference layout: output channel c*25 + tap
    bk = torch.randn(1600, device=DEV) * 0.1
    ab.put(0, a); ab.put(1, x)
    wt = wk.view(64, 25, 64).permute(1, 0, 2).reshape(1600, 64, 1, 1).contiguous()      # tap-major rows t*64 + c
    bt = bk.view(64, 25).t().contiguous()
    wp_ = pack_weight(wt)
    K.check(K.load().savsr_satu_kconv_sta(ctx().handle, ab.a.handle, 0, 1, 2, h, w, wp_.data_ptr(), bt.data_ptr(), 0.1, _stream()))
    torch.cuda.synchronize()
    kern = F.leaky_relu(F.conv2d(a[..., :h, :w].cpu(), wk.cpu(), bk.cpu()), 0.1)
    ref = O.satu_sta_conv(x[..., :h, :w].cpu(), kern)
    full = ab.get(2)
    got = full[..., :h, :w]
    assert float(full[..., h:, :].abs().max() if hp > h else 0.0) == 0.0 and float(full[..., :, w:].abs().max() if wp > w else 0.0) == 0.0
    return dict(sta=assert_close("fused kernel_conv + sta", got, ref))


def check_satu_hr(B=2, h=13, w=15, scale=(2.7, 1.5), seed=1, offset_gain=8.0):
    """savsr_satu_hr: gathers + routed experts + fusion + 3x3 tail + bilinear skip in one kernel, fp32 RGB out, vs the oracle's
    step-by-step restatement (savsr_arch.py:291, 353-376, 738-739) on the same 16-bit-rounded LR features."""
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_state_dict
    sd = make_state_dict(seed)
    # exaggerate the learned offsets so that corners move and the zero-padding border is exercised
    sd["upsample.offset.weight"] = sd["upsample.offset.weight"] * offset_gain
    sd["upsample.st_offset.weight"] = sd["upsample.st_offset.weight"] * offset_gain
    torch.manual_seed(seed)
    hp, wp = h + (h & 1), w + (w & 1)
    res, H, W = satu_index(h, w, scale, sd)
    lr = ArenaBox(2, B, hp, wp)
    x = bf16_round(torch.randn(B, 64, hp, wp, device=DEV)); sta = bf16_round(torch.randn(B, 64, hp, wp, device=DEV))
    lr.put(0, x); lr.put(1, sta)
    xin = torch.rand(B, 7, 3, h, w, device=DEV)
    out = torch.full((B, 3, H, W), float("nan"), device=DEV)
    table = res["table"].to(DEV); by = torch.from_numpy(res["base_y"]).to(DEV); bx = torch.from_numpy(res["base_x"]).to(DEV)
    parts = engine.satu_hr_compose(sd, DEV)
    fmt = K.load().savsr_ctx_get_format(ctx().handle)
    wts = engine.satu_hr_pack(parts, fmt, DEV)
    zb = parts[4].contiguous(); tb = sd["tail.bias"].to(DEV).contiguous()
    K.check(K.load().savsr_satu_hr(ctx().handle, lr.a.handle, 0, 1, h, w, H, W, table.data_ptr(), by.data_ptr(), bx.data_ptr(),
                                   wts.data_ptr(), zb.data_ptr(), tb.data_ptr(), xin.data_ptr(), 7, 3, out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    off, st_off, r = O.satu_heads(sd, "upsample", h, w, scale)
    xc, sc = x[..., :h, :w].cpu(), sta[..., :h, :w].cpu()
    fea = O.satu_expert_mix(sd, "upsample", O.satu_gather(xc, scale, off), r)
    sta_s = O.satu_gather(sc, scale, st_off)
    y = F.conv2d(torch.cat([sta_s, fea], 1), sd["upsample.fusion.weight"], sd["upsample.fusion.bias"])
    tail = F.conv2d(y, sd["tail.weight"], sd["tail.bias"], padding=1)
    skip = F.interpolate(xin[:, 3].cpu(), size=(H, W), mode="bilinear", align_corners=False)
    got = out.cpu()
    assert not bool(torch.isnan(got).any()), "satu_hr left output pixels unwritten"
    err = float(((got - skip) - tail).abs().max())
    info = dict(max_abs=err, tail_absmax=float(tail.abs().max()), rel=err / float(tail.abs().max()))
    assert info["rel"] < 0.02, info          # 16-bit operands (composite weights rounded once), fp32 accumulation
    return info


def check_conv_autograd(B=2, nsrc=2, H=16, W=20, per_sample=False, bias=True, seed=0):
    """savsr_b200.autograd.conv3x3 (forward + tcgen05 dgrad + tcgen05 wgrad) vs fp32 autograd through F.conv2d on the same
    16-bit-rounded operands (row f1, stage A; the reference trains through cuDNN: lbasicsr/models/sr_model.py:101-128)."""
    from savsr_b200.autograd import conv3x3
    torch.manual_seed(seed)
    torch.backends.cudnn.allow_tf32 = False
    Ci = 64 * nsrc
    x = bf16_round(torch.randn(B, Ci, H, W, device=DEV)).requires_grad_(True)
    wshape = (B, 64, Ci, 3, 3) if per_sample else (64, Ci, 3, 3)
    w = bf16_round(torch.randn(*wshape, device=DEV) * 0.05).requires_grad_(True)
    b = (torch.randn(64, device=DEV) * 0.1).requires_grad_(True) if (bias and not per_sample) else None
    g = bf16_round(torch.randn(B, 64, H, W, device=DEV))
    y = conv3x3(x, w, b)
    (y * g).sum().backward()
    got = dict(y=y.detach(), dx=x.grad.clone(), dw=w.grad.clone(), db=b.grad.clone() if b is not None else None)
    x.grad = None; w.grad = None
    if b is not None:
        b.grad = None
    if per_sample:
        yr = F.conv2d(x.reshape(1, B * Ci, H, W), w.reshape(B * 64, Ci, 3, 3), None, 1, 1, 1, groups=B).view(B, 64, H, W)
    else:
        yr = F.conv2d(x, w, b, 1, 1)
    (yr * g).sum().backward()
    ref = dict(y=yr.detach(), dx=x.grad, dw=w.grad, db=b.grad if b is not None else None)
    info = {}
    for k in ("y", "dx", "dw", "db"):
        if ref[k] is None:
            continue
        scale = float(ref[k].abs().max())
        err = float((got[k] - ref[k]).abs().max())
        # y and dx leave through a 16-bit arena slot (one rounding: 2^-8 relative); dw / db are fp32 sums of exact products
        tol = (2.0 ** -7 if k in ("y", "dx") else 2e-4) * scale + 1e-6
        info[k] = dict(max_abs=err, ref_absmax=scale)
        assert err <= tol, (k, info[k], tol)
    return info


def _grad_report(named_gpu, sd_cpu, keys):
    """Error of the gradient per parameter tensor, GPU (16-bit operand convs) vs fp32 CPU autograd on the oracle:
    |g - r| / max(|r|, 1e-2 * largest tensor-gradient norm of the block).  The floor matters: some true gradients are zero or
    ill-conditioned (a conv bias in front of a train-mode BatchNorm; everything upstream of ScaleAttention's BatchNorm over a batch
    of two 1x1 maps, whose output is +-1 whatever its input: scale_routing), where a pure relative error would compare rounding
    noise with rounding noise (and the fp32 atomics of the weight gradient make that noise vary from run to run)."""
    big = max(float(sd_cpu[k].grad.norm()) for k in keys if sd_cpu[k].grad is not None)
    rep = {}
    for k in keys:
        g, r = named_gpu[k].grad, sd_cpu[k].grad
        assert (g is None) == (r is None), f"{k}: gradient present on one side only"
        if r is None:
            continue
        rep[k] = float((g.cpu() - r).norm()) / max(float(r.norm()), 1e-2 * big)
    return rep


def check_train_block(block="residual_group", b=2, h=16, w=20, scale=(2.7, 1.5), seed=1, tol=0.04):
    """Row f1 stage A: gradient parity of the train-mode blocks (3x3 convs on tcgen05 forward / dgrad / wgrad, train-mode BatchNorm,
    OSA fold, pooled means) against fp32 autograd through the CPU oracle, for one WindowUnit_l1 and one ResidualGroup."""
    import savsr_b200
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_state_dict
    from savsr_b200 import train as T
    torch.manual_seed(seed)
    sd = make_state_dict(seed)
    net = savsr_b200.SAVSR().to(DEV)
    net.load_state_dict(sd, strict=True)
    net.train()
    sd_cpu = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    tn = T._Net(net, training=True)
    O.BN_TRAIN = True
    try:
        if block == "residual_group":
            x = torch.randn(b, 64, h, w) * 0.3
            prefix = "RG.1"
            out = T.residual_group(tn, prefix, x.to(DEV))
            ref = O.residual_group(sd_cpu, prefix, x)
        elif block == "window_unit_l1":
            frames = torch.rand(b, 3, 3, h, w)
            hp = torch.randn(b, 64, h, w) * 0.2
            prefix = "f2p_win"
            out = T.window_unit_l1(tn, prefix, frames.to(DEV), hp.to(DEV), scale)
            ref = O.window_unit_l1(sd_cpu, prefix, frames, hp, scale)
        elif block == "osadapt":
            x = torch.randn(b, 64, h, w) * 0.3
            prefix = "adapt.2"
            out = T.osadapt(tn, prefix, x.to(DEV), scale)
            ref = O.osadapt(sd_cpu, prefix, x, scale)
        else:
            raise ValueError(block)
        g = torch.randn_like(ref)
        (out * g.to(DEV)).sum().backward()
        (ref * g).sum().backward()
    finally:
        O.BN_TRAIN = False
    keys = [k for k in sd if k.startswith(prefix + ".") and sd[k].is_floating_point() and "running_" not in k]
    rep = _grad_report(dict(net.named_parameters()), sd_cpu, keys)
    fwd = float((out.detach().cpu() - ref.detach()).abs().max() / (ref.detach().abs().max() + 1e-12))
    worst = max(rep.items(), key=lambda kv: kv[1])
    info = dict(forward_rel=fwd, n_tensors=len(rep), worst=worst, median=float(np.median(list(rep.values()))))
    assert fwd < 0.02, info
    assert all(np.isfinite(v) and v < tol for v in rep.values()), {k: v for k, v in rep.items() if not v < tol}
    return info


def check_train_step(b=2, h=16, w=20, scale=(2, 2), seed=0, steps=3, native_module=True):
    """Whole-net training step (forward + backward through all 707 parameter tensors + Adam + EMA): loss and gradient norm against the
    fp32 CPU oracle at step 0, every parameter receives a gradient, and the loss goes down over a few steps on a fixed batch."""
    import savsr_b200
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_input, make_state_dict
    from savsr_b200 import train as T
    sd = make_state_dict(seed)
    net = savsr_b200.SAVSR().to(DEV)
    net.load_state_dict(sd, strict=True)
    net.native_training = native_module          # True: the module's train-mode forward is the native launch list behind one autograd node
    x = make_input(b, h, w, 1234 + seed)
    H, W = O.get_hw(h, w, scale)
    gt = torch.rand(b, 3, H, W, generator=torch.Generator().manual_seed(5))
    sd_cpu = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    O.BN_TRAIN = True
    try:
        ref_loss = T.charbonnier(O.forward(sd_cpu, x, scale), gt)
        ref_loss.backward()
    finally:
        O.BN_TRAIN = False
    tr = T.Trainer(net, lr=2e-4)
    net.set_scale(scale); net.train()
    out = net(x.to(DEV))
    loss0 = T.charbonnier(out, gt.to(DEV))
    loss0.backward()
    params = dict(net.named_parameters())
    missing = [k for k, p in params.items() if p.grad is None]
    assert not missing, f"{len(missing)} parameters without gradient, e.g. {missing[:3]}"
    gn = float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in params.values())))
    rn = float(torch.sqrt(sum((v.grad ** 2).sum() for k, v in sd_cpu.items() if v.is_floating_point() and v.grad is not None)))
    info = dict(loss=float(loss0), ref_loss=float(ref_loss.detach()), grad_norm=gn, ref_grad_norm=rn, native=bool(net.__dict__.get("_train_state")))
    assert info["native"] == bool(native_module), info
    if native_module:                            # per-tensor parity as for the plan itself (the loss and the optimizer are the caller's here)
        keys = [k for k in params if sd_cpu[k].grad is not None]
        rep = _grad_report(params, sd_cpu, keys)
        info["worst"] = sorted(rep.items(), key=lambda kv: -kv[1])[:3]
        assert all(np.isfinite(v) and v < 0.10 for v in rep.values()), info
    assert abs(float(loss0) - float(ref_loss.detach())) < 2e-3 * max(1.0, abs(float(ref_loss.detach()))), info
    assert abs(gn - rn) < 0.05 * rn, info
    losses = [float(tr.step(x.to(DEV), gt.to(DEV), scale)) for _ in range(steps)]
    info["losses"] = losses
    assert losses[-1] < losses[0], info
    assert int(net.adapt[0].mask[1].num_batches_tracked) - int(sd["adapt.0.mask.1.num_batches_tracked"]) == steps + 1   # BatchNorm ran in train mode
    return info


# ------------------------------------------------------------------------------------------------ training kernels (train_ops.cu)
def check_pack_chunks(seed=4):
    """savsr_pack_conv_chunks (table-driven, one launch): forward orientation == savsr_pack_conv_weight bit for bit; transposed mode ==
    savsr_pack_conv_weight of the flipped, transposed filter of each 64-channel source (what autograd.py packs for the data gradient)."""
    import ctypes as C
    lib = K.load()
    g = torch.Generator().manual_seed(seed)
    info = {}
    for co, ci, ks in ((128, 192, 3), (64, 64, 3), (64, 192, 1), (16, 64, 3)):
        w = torch.randn(co, ci, ks, ks, generator=g).to(DEV)
        taps = ks * ks
        chunks, bufs = [], []
        if co % 64 == 0:
            fwd = torch.zeros(co * ci * taps * 2, dtype=torch.uint8, device=DEV)
            bufs.append(fwd)
            for ng in range(co // 64):
                for s in range(ci // 64):
                    ch = K.PackChunk()
                    ch.w, ch.dst = w.data_ptr(), fwd.data_ptr() + (ng * (ci // 64) + s) * taps * 8192
                    ch.co_total, ch.ci_total, ch.o_base, ch.i_base, ch.ksize, ch.transposed = co, ci, ng * 64, s * 64, ks, 0
                    chunks.append(ch)
        nh = (co + 63) // 64
        tr = torch.zeros((ci // 64) * nh * taps * 8192, dtype=torch.uint8, device=DEV)      # per source s: nh K-stacked chunks
        for s in range(ci // 64):
            for h in range(nh):
                ch = K.PackChunk()
                ch.w, ch.dst = w.data_ptr(), tr.data_ptr() + (s * nh + h) * taps * 8192
                ch.co_total, ch.ci_total, ch.o_base, ch.i_base, ch.ksize, ch.transposed = co, ci, h * 64, s * 64, ks, 1
                chunks.append(ch)
        arr = (K.PackChunk * len(chunks))(*chunks)
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(DEV)
        K.check(lib.savsr_pack_conv_chunks(ctx().handle, table.data_ptr(), 0, len(chunks), _stream()))
        torch.cuda.synchronize()
        if co % 64 == 0:
            assert torch.equal(fwd, pack_weight(w)), (co, ci, ks, "forward")
        wpad = torch.zeros(nh * 64, ci, ks, ks, device=DEV)
        wpad[:co] = w
        for s in range(ci // 64):
            # data-gradient filter of source s: [64 in-channels as outputs][nh * 64 out-channels as inputs], taps flipped
            wt = wpad[:, 64 * s:64 * s + 64].flip(-1, -2).permute(1, 0, 2, 3).contiguous()
            ref = pack_weight(wt)
            got = tr[s * nh * taps * 8192:(s + 1) * nh * taps * 8192]
            assert torch.equal(got, ref), (co, ci, ks, "transposed", s)
        info[(co, ci, ks)] = len(chunks)
    return info


def check_train_elementwise(B=2, H=6, W=20, seed=5):
    """savsr_slot_axpby, savsr_grad_prep (activation derivative, channel scale, pooled-mean gradient, bias gradient, NHWC + NCHW outputs)
    and savsr_slot_to_nchw3 against their definitions; W = 20 exercises the padded pitch (24), W = 72 a second 64-pixel chunk."""
    lib = K.load()
    g = torch.Generator().manual_seed(seed)
    ab = ArenaBox(8, B, H, W)
    dt = _h16()
    pitch = (W + 7) // 8 * 8
    ntslots = 6
    tar = torch.zeros(ntslots * B * 64 * H * pitch, dtype=dt, device=DEV)
    dv, out = torch.randn(B, 64, H, W, generator=g), torch.randn(B, 64, H, W, generator=g)
    ab.put(0, dv); ab.put(1, out)
    dvr, outr = bf16_round(dv).to(DEV), bf16_round(out).to(DEV)
    # axpby
    arr = (K.Axpby * 2)()
    arr[0].dst_slot, arr[0].x_slot, arr[0].y_slot, arr[0].alpha, arr[0].beta = 2, 0, 1, 0.5, 2.0
    arr[1].dst_slot, arr[1].x_slot, arr[1].y_slot, arr[1].alpha, arr[1].beta = 3, 1, -1, 1.0, 0.0
    K.check(lib.savsr_slot_axpby(ctx().handle, ab.a.handle, arr, 2, _stream()))
    assert torch.equal(ab.get(2), bf16_round(0.5 * dvr + 2.0 * outr)) and torch.equal(ab.get(3), outr)
    # grad_prep
    cs, ca = torch.rand(B, 64, generator=g).to(DEV) + 0.5, torch.randn(B, 96, generator=g).to(DEV)
    dbias = torch.zeros(64, device=DEV)
    info = {}
    for act, slope in ((K.ACT_LRELU, 0.2), (K.ACT_RELU, 0.0), (K.ACT_NONE, 0.0)):
        e = (K.GradPrep * 1)()
        e[0].dv_slot, e[0].out_slot, e[0].g_slot, e[0].gt_tslot, e[0].act, e[0].slope = 0, 1, 4, 1, act, slope
        e[0].cscale, e[0].cscale_stride = cs.data_ptr(), 64
        e[0].cadd, e[0].cadd_stride, e[0].cadd_mul = ca.data_ptr() + 32 * 4, 96, 0.25
        e[0].dbias = dbias.data_ptr()
        dbias.zero_()
        K.check(lib.savsr_grad_prep(ctx().handle, ab.a.handle, tar.data_ptr(), ntslots, pitch, e, 1, _stream()))
        ref = dvr * cs.view(B, 64, 1, 1) + 0.25 * ca[:, 32:].view(B, 64, 1, 1)
        if act != K.ACT_NONE:
            ref = ref * torch.where(outr > 0, torch.ones_like(outr), torch.full_like(outr, slope))
        ref = bf16_round(ref)
        got = ab.get(4)
        # one 16-bit rounding of an fp32 value whose last bit may differ (the kernel contracts dv * cs + ca into an FMA)
        assert bool(((got - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-30).all()), (act, float((got - ref).abs().max()))
        assert float((got != ref).float().mean()) < 1e-3, act
        gt = tar.view(ntslots, B, 64, H, pitch)[1].float()
        assert torch.equal(gt[..., :W], got) and float(gt[..., W:].abs().max() if pitch > W else 0.0) == 0.0, act      # both layouts hold the same values
        info[act] = float((dbias - got.sum(dim=(0, 2, 3))).abs().max() / (got.sum(dim=(0, 2, 3)).abs().max() + 1e-9))
        assert info[act] < 1e-5, info
    # three x-shifted NCHW copies
    n3 = (K.Nchw3 * 1)()
    n3[0].x_slot, n3[0].t_slot = 1, 3
    K.check(lib.savsr_slot_to_nchw3(ctx().handle, ab.a.handle, tar.data_ptr(), ntslots, pitch, n3, 1, _stream()))
    torch.cuda.synchronize()
    t3 = tar.view(ntslots, B, 64, H, pitch)[3:6].float()
    ref3 = torch.zeros(3, B, 64, H, pitch, device=DEV)
    ref3[1, ..., :W] = outr
    ref3[0, ..., 1:W] = outr[..., :W - 1]
    if pitch > W:
        ref3[0, ..., W] = outr[..., W - 1]           # the copy shifted right spills one pixel into the padding: harmless (dY is zero there)
    ref3[2, ..., :W - 1] = outr[..., 1:]
    assert torch.equal(t3[1], ref3[1]) and torch.equal(t3[2], ref3[2]) and torch.equal(t3[0][..., :W], ref3[0][..., :W]), "nchw3"
    return info


def check_wgrad_batched(B=2, H=10, W=20, seed=6):
    """savsr_conv_wgrad_batched: two items of one table (a shared 128 -> 64 filter's second source in OIHW layout, a per-sample 64 -> 64
    filter in the [tap][i][o] layout) against F.conv2d's weight gradient on the same rounded operands."""
    import torch.nn.functional as F
    lib = K.load()
    g = torch.Generator().manual_seed(seed)
    dt = _h16()
    pitch = (W + 7) // 8 * 8
    ab = ArenaBox(4, B, H, W)
    ntslots = 8
    tar = torch.zeros(ntslots * B * 64 * H * pitch, dtype=dt, device=DEV)
    x, dy = torch.randn(B, 64, H, W, generator=g), torch.randn(B, 64, H, W, generator=g)
    ab.put(0, x); ab.put(1, dy)
    xr, dyr = bf16_round(x).to(DEV), bf16_round(dy).to(DEV)
    n3 = (K.Nchw3 * 1)()
    n3[0].x_slot, n3[0].t_slot = 0, 0
    K.check(lib.savsr_slot_to_nchw3(ctx().handle, ab.a.handle, tar.data_ptr(), ntslots, pitch, n3, 1, _stream()))
    e = (K.GradPrep * 1)()
    e[0].dv_slot, e[0].out_slot, e[0].g_slot, e[0].gt_tslot, e[0].act = 1, -1, -1, 3, K.ACT_NONE
    K.check(lib.savsr_grad_prep(ctx().handle, ab.a.handle, tar.data_ptr(), ntslots, pitch, e, 1, _stream()))
    dw_shared = torch.zeros(64, 128, 3, 3, device=DEV)
    dw_tio = torch.zeros(B, 9, 64, 64, device=DEV)
    items = (K.WgradItem * 2)()
    items[0].x_tslot, items[0].g_tslot, items[0].dw = 0, 3, dw_shared.data_ptr()
    items[0].ci_total, items[0].ci_off, items[0].o_off, items[0].ksize, items[0].per_sample, items[0].layout = 128, 64, 0, 3, 0, K.WGRAD_OIHW
    items[1].x_tslot, items[1].g_tslot, items[1].dw = 0, 3, dw_tio.data_ptr()
    items[1].ci_total, items[1].ci_off, items[1].o_off, items[1].ksize, items[1].per_sample, items[1].layout = 64, 0, 0, 3, 1, K.WGRAD_TIO
    items[1].sample_stride = 9 * 64 * 64
    table = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(DEV)
    K.check(lib.savsr_conv_wgrad_batched(ctx().handle, tar.data_ptr(), ntslots, B, H, W, pitch, table.data_ptr(), 0, 2, _stream()))
    torch.cuda.synchronize()
    w0 = torch.zeros(64, 64, 3, 3, device=DEV, requires_grad=True)
    ref = torch.autograd.grad(F.conv2d(xr, w0, padding=1), w0, dyr)[0]                     # [o, i, ky, kx]
    scale = float(ref.abs().max())
    e0 = float((dw_shared[:, 64:] - ref).abs().max()) / scale
    assert float(dw_shared[:, :64].abs().max()) == 0.0 and e0 < 1e-4, e0
    per = torch.stack([torch.autograd.grad(F.conv2d(xr[n:n + 1], w0, padding=1), w0, dyr[n:n + 1])[0] for n in range(B)])      # [B, o, i, 3, 3]
    e1 = float((dw_tio - per.permute(0, 3, 4, 2, 1).reshape(B, 9, 64, 64)).abs().max()) / scale
    assert e1 < 1e-4, e1
    return dict(shared_rel=e0, per_sample_tio_rel=e1)


def check_adam_ema(n=100003, steps=5, seed=7):
    """savsr_adam_ema against torch.optim.Adam (lr 2e-4, betas 0.9 / 0.99, eps 1e-8, no weight decay) + the EMA of base_model.py:75-82."""
    lib = K.load()
    g = torch.Generator().manual_seed(seed)
    p0 = torch.randn(n, generator=g).to(DEV)
    p, m, v, ema = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), p0.clone()
    step_t = torch.zeros(1, device=DEV)
    rp = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([rp], lr=2e-4, betas=(0.9, 0.99), eps=1e-8)
    rema = p0.clone()
    for _ in range(steps):
        grad = torch.randn(n, generator=g).to(DEV) * 1e-3
        step_t.add_(1.0)
        K.check(lib.savsr_adam_ema(ctx().handle, p.data_ptr(), (grad * 2).data_ptr(), m.data_ptr(), v.data_ptr(), ema.data_ptr(), n, 2e-4, 0.9, 0.99, 1e-8,
                                   step_t.data_ptr(), 0.999, 0.5, _stream()))          # gradient pre-scaled by 2, grad_scale 0.5
        torch.cuda.synchronize()
        rp.grad = grad.clone()
        opt.step()
        rema.mul_(0.999).add_(rp.detach(), alpha=0.001)
    e_p = float((p - rp.detach()).abs().max())
    e_ema = float((ema - rema).abs().max())
    assert e_p < 2e-6 and e_ema < 2e-6, (e_p, e_ema)
    return dict(param_max_abs=e_p, ema_max_abs=e_ema, moved=float((p - p0).abs().max()))


def check_sta_lrelu(B=2, C=5, h=7, w=9, seed=3):
    """The fused sta_conv + LeakyReLU op of the training step (savsr_b200.autograd.sta_lrelu) against its ATen formulation
    (replicate pad -> unfold -> product with the activated kernels -> sum over the 25 taps), forward and both gradients, fp32."""
    import torch.nn.functional as F
    from savsr_b200 import autograd as A
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, h, w, generator=g).to(DEV).requires_grad_(True)
    k = torch.randn(B, C * 25, h, w, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(B, C, h, w, generator=g).to(DEV)
    out = A.sta_lrelu(x, k, 0.1)
    gx, gk = torch.autograd.grad(out, [x, k], go)
    kern = F.leaky_relu(k, 0.1).view(B, C, 25, h * w)
    ref = (F.unfold(F.pad(x, (2, 2, 2, 2), mode="replicate"), 5).view(B, C, 25, h * w) * kern).sum(2).view(B, C, h, w)
    rx, rk = torch.autograd.grad(ref, [x, k], go)
    info = {}
    for name, a, b in (("out", out, ref), ("dx", gx, rx), ("dk", gk, rk)):
        err = float((a - b).abs().max())
        info[name] = err
        assert err <= 2e-5 * max(1.0, float(b.abs().max())), (name, err)
    return info


def check_trainplan(b=2, h=16, w=20, scale=(2, 2), seed=0, tol_worst=0.10, tol_cos=0.999, steps=0, graph=False, native_attn=True, native_mask=True,
                    oracle_device="cpu"):
    """Row f1 stage B: the NATIVE training step (savsr_b200.trainplan: arena-resident forward / dgrad / batched wgrad, table-driven
    weight packing) against fp32 CPU autograd through the oracle: loss, per-parameter gradient error
    |g - r| / max(|r|, 1 % of the largest tensor gradient), and the cosine of the whole flat gradient.
    oracle_device="cuda": the same oracle evaluated in fp32 on the GPU (TF32 off) -- what makes the BASELINE cfg-5 batch (4 x 7 x 3 x 64 x 64)
    checkable in seconds; the oracle stays the checker, the product path is unchanged."""
    import savsr_b200
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_input, make_state_dict
    from savsr_b200 import train as T
    from savsr_b200 import trainplan as TP
    sd = make_state_dict(seed)
    net = savsr_b200.SAVSR().to(DEV)
    net.load_state_dict(sd, strict=True)
    net.set_scale(scale); net.train()
    x = make_input(b, h, w, 1234 + seed)
    H, W = O.get_hw(h, w, scale)
    gt = torch.rand(b, 3, H, W, generator=torch.Generator().manual_seed(5))
    odev = torch.device(oracle_device)
    sd_cpu = {k: (v.clone().to(odev).requires_grad_(True) if v.is_floating_point() else v.clone().to(odev)) for k, v in sd.items()}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    O.BN_TRAIN = True
    try:
        ref_loss = T.charbonnier(O.forward(sd_cpu, x.to(odev), scale), gt.to(odev))
        ref_loss.backward()
    finally:
        O.BN_TRAIN = False
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    if odev.type != "cpu":
        moved = {}                                     # the report below compares on the host
        for k, v in sd_cpu.items():
            c = v.detach().cpu()
            if v.grad is not None:
                c.requires_grad_(True)
                c.grad = v.grad.cpu()
            moved[k] = c
        sd_cpu = moved
        ref_loss = ref_loss.detach().cpu()
    tr = TP.NativeTrainer(net, use_graph=graph, native_attn=native_attn, native_mask=native_mask)
    plan = tr.plan_for(x.to(DEV), scale)
    plan.x_in.copy_(x.to(DEV)); plan.gt.copy_(gt.to(DEV))
    loss = tr._fwd_bwd(plan)
    torch.cuda.synchronize()
    params = dict(net.named_parameters())
    keys = [k for k in params if sd_cpu[k].grad is not None]
    rep = _grad_report(params, sd_cpu, keys)
    gflat = torch.cat([params[k].grad.flatten().cpu() for k in keys]).double()
    rflat = torch.cat([sd_cpu[k].grad.flatten() for k in keys]).double()
    cos = float(torch.dot(gflat, rflat) / (gflat.norm() * rflat.norm()))
    worst = sorted(rep.items(), key=lambda kv: -kv[1])[:5]
    nbt = int(net.f2p_win.blocks[1].osconv.attention.bn.num_batches_tracked) - int(sd["f2p_win.blocks.1.osconv.attention.bn.num_batches_tracked"])
    assert nbt == 5, nbt                      # the shared l1 BatchNorm ran once per propagation iteration, in train mode
    info = dict(loss=float(loss), ref_loss=float(ref_loss.detach()), cos=cos, grad_norm=float(gflat.norm()), ref_grad_norm=float(rflat.norm()),
                worst=worst, median=float(np.median(list(rep.values()))), launches=dict(plan.launches), slots=plan.n_slots, tslots=plan.n_tslots)
    assert abs(float(loss) - float(ref_loss.detach())) < 2e-3 * max(1.0, abs(float(ref_loss.detach()))), info
    assert np.isfinite(cos) and cos > tol_cos, info
    assert all(np.isfinite(v) and v < tol_worst for v in rep.values()), info
    if steps:
        # a second scale interleaved: the plans of all scales share one pair of arenas (and, with graphs, replay into the same addresses)
        scale_b = (1.5, 4) if tuple(scale) != (1.5, 4) else (2, 2)
        gt_b = torch.rand(b, 3, *O.get_hw(h, w, scale_b), generator=torch.Generator().manual_seed(6)).to(DEV)
        losses, losses_b = [], []
        for _ in range(steps):
            losses.append(float(tr.step(x.to(DEV), gt.to(DEV), scale)))
            losses_b.append(float(tr.step(x.to(DEV), gt_b, scale_b)))
        info["losses"], info["losses_second_scale"] = losses, losses_b
        assert len(tr.plans) == 2 and len({p.arena_t.data_ptr() for p in tr.plans.values()}) == 1, "plans of different scales must share the arenas"
        assert all(np.isfinite(l) for l in losses + losses_b) and losses[-1] < losses[0] and losses_b[-1] < losses_b[0], info
    return info


def check_train_golden(path):
    """Row f1 against the UNMODIFIED reference directly: loss, per-parameter gradient norms, the small gradient tensors and the projection of
    the whole gradient on a random direction, as recorded from the reference module in train() mode with its own CharbonnierLoss
    (tests/golden/train_*.npz, scripts/make_golden.py --train-only), vs SAVSR.forward in train mode + backward on the GPU (the native launch
    list for even sizes, the stage-A tape for odd ones).  Bounds as for the oracle comparison: 16-bit operands and 16-bit stored gradients."""
    import savsr_b200
    from oracle.state_dict_fixture import make_input, make_state_dict
    from savsr_b200 import train as T
    g = np.load(path)
    b, h, w, sd_seed, in_seed = (int(v) for v in g["dims"])
    scale = tuple(float(v) if float(v) != int(v) else int(v) for v in g["scale"])
    net = savsr_b200.SAVSR().to(DEV)
    net.load_state_dict(make_state_dict(sd_seed), strict=True)
    net.set_scale(scale); net.train()
    x = make_input(b, h, w, in_seed).to(DEV)
    gt = torch.from_numpy(g["gt"]).to(DEV)
    out = net(x)
    loss = T.charbonnier(out, gt)
    loss.backward()
    torch.cuda.synchronize()
    native = bool(net.__dict__.get("_train_state"))
    loss = loss.detach()
    assert abs(float(loss) - float(g["loss"])) < 2e-3, (float(loss), float(g["loss"]))
    names = [str(n) for n in g["names"]]
    params = dict(net.named_parameters())
    assert names == list(params.keys())
    norms = g["grad_norm"]
    big = float(norms.max())
    gen = torch.Generator().manual_seed(4242)
    proj, worst, worst_full = 0.0, (0.0, None), (0.0, None)
    for k, n_ref in zip(names, norms):
        gr = params[k].grad
        assert gr is not None, k
        gr = gr.detach().cpu()
        r = torch.randn(gr.shape, generator=gen, dtype=torch.float64)
        proj += float((gr.double() * r).sum())
        e = abs(float(gr.double().norm()) - float(n_ref)) / max(float(n_ref), 1e-2 * big)
        worst = max(worst, (e, k))
        if "grad." + k in g.files:
            ref = torch.from_numpy(g["grad." + k])
            worst_full = max(worst_full, (float((gr - ref).norm()) / max(float(ref.norm()), 1e-2 * big), k))
    gnorm = float(np.sqrt((norms ** 2).sum()))
    # BatchNorm buffers after the step (momentum update of the running statistics with the UNBIASED batch variance, call counters)
    bufs = dict(net.named_buffers())
    worst_buf = (0.0, None)
    nbuf = 0
    for key in g.files:
        if not key.startswith("buf."):
            continue
        k = key[4:]
        ref = torch.from_numpy(g[key])
        got = bufs[k].detach().cpu()
        nbuf += 1
        if k.endswith("num_batches_tracked"):
            assert int(got) == int(ref), (k, int(got), int(ref))
        else:
            worst_buf = max(worst_buf, (float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-3), k))
    assert nbuf > 0
    info = dict(native=native, loss=float(loss), ref_loss=float(g["loss"]), worst_norm=worst, worst_small_tensor=worst_full,
                proj_err_over_gnorm=abs(proj - float(g["grad_proj"])) / gnorm, worst_bn_buffer=worst_buf, bn_buffers=nbuf)
    assert worst_buf[0] < 2e-2, info
    # measured: 0.8 % / 1.6 % (b = 3, stage-A tape) and 9.4 % / 7.3 % (b = 2, native; the worst tensors are OSAdapt's scale_routing, upstream of a
    # train-mode BatchNorm over a batch of TWO 1x1 maps whose output is +-1 whatever its input -- the ill-conditioned case named in _grad_report)
    assert worst[0] < 0.15 and worst_full[0] < 0.15, info
    assert info["proj_err_over_gnorm"] < 0.05, info
    return info


def check_fp16_range_guard(seed=0):
    """precision='fp16': the first forward of a plan measures the largest trunk activation; an input that drives it beyond a quarter of the
    fp16 range raises instead of returning saturated values (bf16 handles the same input)."""
    import savsr_b200
    from oracle.state_dict_fixture import make_input, make_state_dict
    net = savsr_b200.SAVSR().to(DEV).eval()
    net.load_state_dict(make_state_dict(seed), strict=True)
    net.set_scale((2, 2))
    net.precision = "fp16"
    x = make_input(1, 16, 20, 1234 + seed).to(DEV)
    with torch.no_grad():
        y = net(x)
        peak = net.fp16_peak_activation
        assert torch.isfinite(y).all() and 0 < peak < 65504 / 4, ("peak", peak)
        big = x * (4 * 65504 / peak)                   # the trunk is roughly homogeneous in its input: this overshoots the guard
        raised = False
        try:
            net(big[:, :, :, :14, :18].contiguous())   # a new plan (another size), so the check runs again
        except FloatingPointError:
            raised = True
        assert raised, ("not raised", net.fp16_peak_activation)
        net.precision = "bf16"
        net(big[:, :, :, :14, :18].contiguous())       # the guard belongs to the fp16 path only
    return dict(peak_unit_input=peak)


def check_c_plan(b=2, h=13, w=15, scale=(1.5, 4), seed=2):
    """savsr_forward / savsr_plan_run (the recorded launch list replayed by one C call) against the same launches issued one by one from
    Python: bit-identical output, and the plan holds one record per op."""
    import savsr_b200
    from oracle.state_dict_fixture import make_input, make_state_dict
    net = savsr_b200.SAVSR().to(DEV).eval()
    net.load_state_dict(make_state_dict(seed), strict=True)
    net.set_scale(scale)
    x = make_input(b, h, w, 1234 + seed).to(DEV)
    with torch.no_grad():
        plan = net.plan_for(x)
        assert plan.cplan is not None and len(plan.cplan) == len(plan.ops), (len(plan.ops), plan.cplan)
        plan.x_in.copy_(x)
        st = torch.cuda.current_stream().cuda_stream
        for op in plan.ops:                         # the Python loop
            rc = op(st)
            assert not rc, rc
        ref = plan.out.clone()
        plan.out.zero_()
        out = torch.empty_like(ref)
        plan.forward_c(x.contiguous(), out)         # one C call
        torch.cuda.synchronize()
    assert torch.equal(out, ref), float((out - ref).abs().max())
    return dict(ops=len(plan.ops), shape=tuple(out.shape))


def check_img_metrics(n=3, H=37, W=53, seed=31):
    """tensor2img (bit-exact uint8 BGR) and PSNR-Y on the device vs the oracle's restatement of the reference metric chain."""
    from oracle import savsr_oracle as O
    from savsr_b200 import postproc
    torch.manual_seed(seed)
    gt = torch.rand(n, 3, H, W)
    sr = (gt + 0.05 * torch.randn(n, 3, H, W)).clone()
    sr[0, :, :4] = 1.7; sr[0, :, 4:8] = -0.3                     # exercise the clamp
    k = torch.arange(256, dtype=torch.float32)
    sr[1, 0, 0, :52] = ((k[:52] * 2 + 0.5) / 255.0)              # x.5 ties: round half to even
    sr[2] = gt[2]                                                # identical frame -> inf
    img, psnr = postproc.tensor2img_psnr(sr.to(DEV), gt.to(DEV))
    for i in range(n):
        assert np.array_equal(img[i].cpu().numpy(), O.tensor2img(sr[i])), f"uint8 image {i} not bit-exact"
    ref = [O.psnr_y(sr[i], gt[i]) for i in range(n)]
    got = psnr.cpu().tolist()
    assert got[2] == float("inf") and ref[2] == float("inf")
    assert all(abs(a - b) < 1e-6 for a, b in zip(got[:2], ref[:2])), (got, ref)
    # SSIM-Y: float64 on both sides, only the summation order differs
    ssim = postproc.ssim_y(sr.to(DEV), gt.to(DEV)).cpu().tolist()
    sref = [O.ssim_y(sr[i], gt[i]) for i in range(n)]
    assert all(abs(a - b) < 1e-9 for a, b in zip(ssim, sref)), (ssim, sref)
    assert abs(ssim[2] - 1.0) < 1e-12
    # the reference's own numbers (cv2.filter2D path) for the committed known-answer frames
    kat = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_kat.npz"))
    ksr, kgt = torch.from_numpy(kat["sr"]).to(DEV), torch.from_numpy(kat["gt"]).to(DEV)
    kss = postproc.ssim_y(ksr, kgt).cpu().numpy()
    assert np.abs(kss - kat["ssim_y"]).max() < 1e-9, (kss, kat["ssim_y"])
    _, kps = postproc.tensor2img_psnr(ksr, kgt, want_image=False)
    assert np.abs(kps.cpu().numpy() - kat["psnr_y"]).max() < 1e-6
    return dict(psnr=got, ref=ref, ssim=ssim, ssim_ref=sref)



# ------------------------------------------------------------------------------------------------ data path (row f2)
def check_lr_synthesis():
    """Device LR synthesis (as_mod_crop + antialiased bicubic downsample from uint8 BGR frames): bit-exact against the
    reference's own outputs (tests/golden/lr_kat.npz) and against the oracle at larger, odd shapes."""
    from oracle import lr_synthesis as L
    from savsr_b200 import datapath
    kat = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lr_kat.npz"))
    out = {}
    for name in ("x4", "x2p7", "x1p5x4", "x3p9", "x1p2x1p7", "x7p3x5p1"):
        frames, scale = kat[f"{name}.frames"], tuple(float(v) for v in kat[f"{name}.scale"])
        lr, gt = datapath.synthesize_lr(torch.from_numpy(frames).to(DEV), scale)
        assert tuple(lr.shape) == kat[f"{name}.lr"].shape, (name, lr.shape)
        assert np.array_equal(lr.cpu().numpy(), kat[f"{name}.lr"]), f"{name}: LR differs from the reference"
        hc, wc = (int(v) for v in kat[f"{name}.crop"])
        assert np.array_equal(gt.cpu().numpy(), L.frames_to_rgb(frames, (hc, wc))), f"{name}: GT tensor differs"
        out[name] = "bit-exact"
    rng = np.random.default_rng(11)
    for (t, h, w, scale) in ((2, 151, 203, (2.7, 2.7)), (1, 144, 180, (4, 4)), (1, 97, 131, (1.5, 4)), (1, 80, 64, (1, 2))):
        frames = rng.integers(0, 256, size=(t, h, w, 3), dtype=np.uint8)
        lr_o, gt_o = L.synthesize_lr(frames, scale)
        lr, gt = datapath.synthesize_lr(torch.from_numpy(frames).to(DEV), scale)
        assert np.array_equal(lr.cpu().numpy(), lr_o) and np.array_equal(gt.cpu().numpy(), gt_o), (h, w, scale)
        out[f"{h}x{w}@{scale}"] = "bit-exact"
    # row f4: float frames, up- and down-sizing by a few pixels (sr_model.py:291-304); CPU-ATen order via the oracle, and
    # torch's own CUDA kernel (what the reference runs there; a one-pass 2-D formulation) within float32 rounding
    import torch.nn.functional as F
    for (n, h, w, size) in ((2, 57, 70, (60, 66)), (1, 173, 389, (176, 384)), (1, 40, 40, (40, 31))):
        x = torch.randn(n, 3, h, w) * 0.3 + 0.5
        y = datapath.resize_aa_bicubic(x.to(DEV), size)
        yo = L.resize_aa_bicubic(x.numpy(), size)
        assert np.array_equal(y.cpu().numpy(), yo), ("vs oracle", h, w, size, float(np.abs(y.cpu().numpy() - yo).max()))
        ref = F.interpolate(x.to(DEV), size=size, mode="bicubic", antialias=True, align_corners=False)
        e_cuda = float((y - ref).abs().max())
        # ATen's CUDA kernel (one-pass 2-D, its own weight arithmetic) is what the reference runs for this step on a GPU; it
        # differs from ATen's CPU kernel -- and therefore from ours -- by ~3e-5, 1% of one uint8 level
        assert e_cuda < 1e-4, ("vs torch CUDA", h, w, size, e_cuda)
        out[f"resize {h}x{w}->{size}"] = f"bit-exact (CPU ATen order), {e_cuda:.1e} vs torch CUDA"
    # size-independent property at a full Vid4 frame: a constant image stays constant (weights sum to 1 within rounding)
    flat = torch.full((1, 576, 720, 3), 137, dtype=torch.uint8, device=DEV)
    lr, _ = datapath.synthesize_lr(flat, (4, 4), want_gt=False)
    assert tuple(lr.shape) == (1, 3, 144, 180) and float((lr - 137.0 / 255.0).abs().max()) < 5e-7
    return out


def check_evaluate_clip(seed=5):
    """The device test loop (LR synthesis -> windows -> net -> tensor2img / PSNR-Y / SSIM-Y) against the oracle chain on a
    tiny clip: LR bit-exact, SR within the bf16 tolerance, metrics consistent with the oracle's metrics of the same SR."""
    import savsr_b200
    from oracle import lr_synthesis as L
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_state_dict
    from savsr_b200 import datapath, sharding
    rng = np.random.default_rng(seed)
    T, H, W, scale = 4, 50, 66, (4, 4)
    base = rng.integers(0, 256, size=(1, H // 4 + 1, W // 4 + 1, 3), dtype=np.uint8)
    frames = np.repeat(np.repeat(base, 4, 1), 4, 2)[:, :H, :W]                 # smooth-ish content
    frames = np.clip(frames.astype(np.int32) + rng.integers(-20, 21, size=(T, H, W, 3)), 0, 255).astype(np.uint8)
    sd = make_state_dict(0)
    net = savsr_b200.SAVSR().to(DEV).eval()
    net.load_state_dict(sd)
    net.conv_impl = "halo"
    net.set_scale(scale)
    with torch.no_grad():
        res = datapath.evaluate_clip(net, torch.from_numpy(frames).to(DEV), scale, batch=3)
    lr_o, gt_o = L.synthesize_lr(frames, scale)
    lr_t = torch.from_numpy(lr_o)
    sr_o = torch.cat([O.forward(sd, sharding.gather_windows(lr_t, [i]), scale) for i in range(T)])
    sr = res["sr"].cpu()
    assert tuple(sr.shape) == tuple(sr_o.shape) == (T, 3, 48, 64)
    err = float((sr - sr_o).abs().max())
    assert err < 4e-3, err
    gt_t = torch.from_numpy(gt_o)
    for i in range(T):
        assert np.array_equal(res["images"][i].cpu().numpy(), O.tensor2img(sr[i]))
        assert abs(float(res["psnr_y"][i]) - O.psnr_y(sr[i], gt_t[i])) < 1e-6
        assert abs(float(res["ssim_y"][i]) - O.ssim_y(sr[i], gt_t[i])) < 1e-9
    dpsnr = abs(O.psnr_y(sr_o, gt_t) - float(res["psnr_y"].mean()))
    assert dpsnr < 0.05, dpsnr                                                  # north-star PSNR gate, bf16 path
    # a rank's share of the clip (sharding.shard_frames), in any order: the same frames, scored against the same ground-truth rows
    with torch.no_grad():
        sub = datapath.evaluate_clip(net, torch.from_numpy(frames).to(DEV), scale, frames=[3, 1], batch=3)
    assert tuple(sub["sr"].shape) == (2, 3, 48, 64)
    assert float((sub["sr"].cpu() - sr[[3, 1]]).abs().max()) < 1e-4              # (another batch size regroups the pooled sums: not bit-identical)
    assert float((sub["psnr_y"].cpu() - res["psnr_y"].cpu()[[3, 1]]).abs().max()) < 1e-2
    assert float((sub["ssim_y"].cpu() - res["ssim_y"].cpu()[[3, 1]]).abs().max()) < 1e-4
    return dict(sr_max_abs=err, psnr_delta_db=dpsnr, psnr=[round(float(v), 3) for v in res["psnr_y"]])


def check_clip_prefetch(seed=8):
    """ClipPrefetcher + evaluate_clips: same per-clip metrics as evaluating each clip alone; pinned and pageable hosts."""
    import savsr_b200
    from oracle.state_dict_fixture import make_state_dict
    from savsr_b200 import datapath
    rng = np.random.default_rng(seed)
    clips = [(f"clip{i}", rng.integers(0, 256, size=(4 + i, 32, 40, 3), dtype=np.uint8)) for i in range(3)]
    hosts = [(n, torch.from_numpy(f).pin_memory() if i % 2 == 0 else f) for i, (n, f) in enumerate(clips)]
    net = savsr_b200.SAVSR().to(DEV).eval()
    net.load_state_dict(make_state_dict(0))
    net.set_scale((2, 2))
    with torch.no_grad():
        got = datapath.evaluate_clips(net, hosts, (2, 2), DEV, batch=2)
        for name, frames in clips:
            ref = datapath.evaluate_clip(net, torch.from_numpy(frames).to(DEV), (2, 2), batch=3)
            assert abs(got[name]["psnr_y"] - float(ref["psnr_y"].mean())) < 1e-9, name
            assert abs(got[name]["ssim_y"] - float(ref["ssim_y"].mean())) < 1e-12, name
    assert set(got) == {"clip0", "clip1", "clip2", "average"}
    assert abs(got["average"]["psnr_y"] - sum(got[f"clip{i}"]["psnr_y"] for i in range(3)) / 3) < 1e-12
    return got


# ------------------------------------------------------------------------------------------------ whole forward
TAPS = ("f2p_last", "p2f_last", "align", "rg0", "rg3", "trunk", "satu_sta")


def run_forward(sd, x, scale, impl="tap", taps=(), graph=False, precision="bf16"):
    import savsr_b200
    net = savsr_b200.SAVSR().to(DEV)
    net.load_state_dict(sd, strict=True)
    net.eval()
    net.conv_impl, net.use_graph, net.debug_taps, net.precision = impl, graph, tuple(taps), precision
    net.set_scale(scale)
    with torch.no_grad():
        y = net(x.to(DEV))
    plan = net.plan_for(x.to(DEV))
    torch.cuda.synchronize()
    return y.cpu(), {k: plan.read_tap(k).cpu() for k in taps}, plan


def stage_report(got: dict, ref: dict, h, w) -> dict:
    """Relative error (max-abs / ref abs-max) per stage on the unpadded region."""
    rep = {}
    for k, g in got.items():
        if k not in ref:
            continue
        r = ref[k]
        g = g[..., :r.shape[-2], :r.shape[-1]]
        rep[k] = float((g - r).abs().max() / (r.abs().max() + 1e-12))
    return rep


def check_forward(b=1, h=16, w=20, scale=(2, 2), sd_seed=0, in_seed=1234, impl="tap", graph=False, tol=5e-3, stage_tol=0.05,
                  precision="bf16"):
    """End-to-end SAVSR forward vs the CPU oracle (bf16 operand path: tolerance on max-abs and PSNR)."""
    from oracle import savsr_oracle as O
    from oracle.state_dict_fixture import make_input, make_state_dict
    sd = make_state_dict(sd_seed)
    x = make_input(b, h, w, in_seed)
    probes = {}
    y_ref = O.forward(sd, x, scale, probes)
    y, taps, plan = run_forward(sd, x, scale, impl=impl, taps=TAPS, graph=graph, precision=precision)
    ref_stage = dict(f2p_last=probes["f2p_last"], p2f_last=probes["p2f_last"], align=probes["align"], rg0=probes["rg0"],
                     rg3=probes["rg3"], trunk=probes["trunk"], satu_sta=probes["satu_sta"])
    rep = stage_report(taps, ref_stage, h, w)
    err = float((y - y_ref).abs().max())
    tail_rel = float(((y - probes["skip"]) - probes["tail"]).abs().max() / (probes["tail"].abs().max() + 1e-12))
    info = dict(shape=tuple(y.shape), max_abs=err, tail_rel=tail_rel, psnr_vs_oracle=O.psnr_y(y, y_ref), stages=rep,
                launches=plan.n_launches)
    assert y.shape == y_ref.shape, info
    assert err < tol, info
    assert all(v < stage_tol for v in rep.values()), info
    return info
