"""CPU-only: the host side of the native training plan (savsr_b200/trainplan.py) against the recording fake of the C library:
the backward launch list is generated from the forward list, so the dataflow invariants the kernels rely on are checked here
(every gradient slot is written before it is read or accumulated, no launch accumulates twice into one slot, every weight
gets a weight-gradient item whose operands exist).  No kernel runs; the ATen islands execute on CPU tensors."""
import contextlib

import pytest
import torch

from fake_lib import mocked_engine
from savsr_b200 import _capi as K


@pytest.fixture(scope="module")
def recorded():
    import savsr_b200
    from savsr_b200 import trainplan as TP
    torch.manual_seed(0)
    net = savsr_b200.SAVSR()
    with mocked_engine() as lib:
        import savsr_b200.autograd as A
        import torch.nn.functional as F
        saved = (TP.context, TP._require_cuda, A.sta_lrelu)
        TP.context = lambda idx: __import__("savsr_b200.engine", fromlist=["context"]).context(idx)
        TP._require_cuda = lambda dev: None

        def sta_lrelu_torch(x, kpre, slope=0.1):           # the ATen formulation of the fused op (no kernel runs in this test)
            b, c, h, w = x.shape
            kern = F.leaky_relu(kpre, slope).view(b, c, 25, h * w)
            return (F.unfold(F.pad(x, (2, 2, 2, 2), mode="replicate"), 5).view(b, c, 25, h * w) * kern).sum(2).view(b, c, h, w)
        A.sta_lrelu = sta_lrelu_torch
        try:
            tr = TP.NativeTrainer(net, use_graph=False)
            lq = torch.rand(2, 7, 3, 16, 24)
            gt = torch.rand(2, 3, 32, 48)
            plan = tr.plan_for(lq, (2, 2))
            lib.calls.clear()
            loss = tr.step(lq, gt, (2, 2))
            calls = list(lib.calls)
        finally:
            TP.context, TP._require_cuda, A.sta_lrelu = saved
    return net, tr, plan, calls, loss


def _entries(args, idx_arr, idx_n):
    arr, n = args[idx_arr], args[idx_n]
    return [arr[i] for i in range(n)]


def test_step_runs_and_census(recorded):
    net, tr, plan, calls, loss = recorded
    names = [c[0] for c in calls]
    assert loss.numel() == 1          # (its value is meaningless here: the fake library never fills the exported activations)
    assert names.count("savsr_pack_frames") == 1
    assert names.count("savsr_adam_ema") == 1
    assert names.count("savsr_ca_scale_residual") == 32
    # ONE table-driven pack of the shared filters; the per-sample folded kernels of OSA-Conv are written packed by its prologue
    assert names.count("savsr_pack_conv_chunks") == 1
    # train-mode OSA prologue / fold backward: l1 5 iterations x 3 blocks (both directions per launch), l2 2, adapt 4
    assert names.count("savsr_osa_prologue_train") == 15 + 2 + 4 and names.count("savsr_osa_fold_backward") == 15 + 2 + 4
    assert names.count("savsr_slot_channel_dot") == 32 and names.count("savsr_ca_backward") == 32
    assert names.count("savsr_mask_forward_train") == 4 and names.count("savsr_mask_backward_train") == 4
    # weight gradients: one inline launch per OSA-Conv launch (l1: both directions share one) + the final batched launch
    assert names.count("savsr_conv_wgrad_batched") == 15 + 2 + 4 + 1
    # forward convolutions as in the inference plan, minus the N = 16 mask convs (mask net = ATen island here)
    fwd_convs = 5 * 14 + 9 + 4 * 18 + 1
    assert names.count("savsr_conv") > 2 * fwd_convs


def test_gradient_slots_are_written_before_use(recorded):
    net, tr, plan, calls, _ = recorded
    written = set()          # slots holding data
    for name, args in calls:
        if name == "savsr_pack_frames":
            written.add(args[6])
        elif name == "savsr_arena_import":
            written.add(args[1])
        elif name == "savsr_arena_export":
            assert args[1] in written, f"export of slot {args[1]} before it was written"
        elif name == "savsr_conv":
            dsts = []
            for g in _entries(args, 2, 3):
                for i in range(g.nsrc):
                    assert g.src_slot[i] in written or g.src_slot[i] == 0, f"conv reads unwritten slot {g.src_slot[i]}"
                if g.res1_slot >= 0:
                    assert g.res1_slot in written
                dsts.append(g.dst_slot)
            assert len(set(dsts)) == len(dsts), "two groups of one launch write the same slot"
            written.update(dsts)
        elif name == "savsr_slot_axpby":
            dsts = []
            for e in _entries(args, 2, 3):
                assert e.x_slot in written or e.x_slot == 0
                assert e.y_slot < 0 or e.y_slot in written or e.y_slot == 0
                dsts.append(e.dst_slot)
            assert len(set(dsts)) == len(dsts)
            written.update(dsts)
        elif name == "savsr_ca_scale_residual":
            assert args[2] in written and args[3] in written
            written.add(args[4])
        elif name == "savsr_mask_forward_train":
            assert all(a in written for a in args[3:6])
            written.add(args[6])
        elif name == "savsr_mask_backward_train":
            assert all(a in written for a in args[3:6])
            written.update(args[6:9])
        elif name == "savsr_grad_prep":
            for e in _entries(args, 5, 6):
                assert e.dv_slot in written
                assert e.act == K.ACT_NONE or e.out_slot in written
                if e.g_slot >= 0:
                    written.add(e.g_slot)


def test_every_trunk_weight_has_a_weight_gradient_item(recorded):
    net, tr, plan, calls, _ = recorded
    G = tr.flat.G
    targets = {}
    for it in plan.witems:
        targets.setdefault(it.dw, []).append(it)
    want = [n for n, p in net.named_parameters() if n.endswith(".weight") and p.dim() == 4 and p.shape[0] % 64 == 0 and p.shape[1] % 64 == 0
            and not n.startswith("upsample.")]
    for n in want:
        its = targets.get(G[n].data_ptr())
        assert its, f"no weight-gradient item for {n}"
        ci, co = G[n].shape[1], G[n].shape[0]
        uses = len(its) // ((ci // 64) * (co // 64))
        assert len(its) == uses * (ci // 64) * (co // 64), n
        assert {it.ci_off for it in its} == set(range(0, ci, 64)) and {it.o_off for it in its} == set(range(0, co, 64)), n
    # shared l1 weights are used once per propagation iteration
    assert len(targets[G["f2p_win.blocks.1.conv2.0.weight"].data_ptr()]) == 5 * 2
    # T-slots: every item's operands lie inside the T-arena
    for it in plan.witems:
        assert 0 <= it.x_tslot and it.x_tslot + 3 <= plan.n_tslots and 0 <= it.g_tslot < plan.n_tslots


def test_every_parameter_has_a_gradient_path(recorded):
    """Static completeness of the backward program: each of the trainable parameter tensors is the target of a weight-gradient item, a
    bias-gradient pointer of savsr_grad_prep, a field of the OSA / mask / channel-attention backward calls, a scatter of a staged gradient,
    or belongs to the SATU / tail island (whose gradients autograd delivers).  (The GPU parity test checks the values; this one checks that
    nothing is forgotten when the plan's structure changes.)"""
    import ctypes as C
    net, tr, plan, calls, _ = recorded
    G = tr.flat.G
    by_ptr = {}
    for name, g in G.items():
        by_ptr[g.data_ptr()] = name
    covered = set()

    def hit(ptr, nbytes=1):
        if not ptr:
            return
        for base, name in by_ptr.items():
            if base <= ptr < base + max(G[name].numel() * 4, 1):
                covered.add(name)

    for it in plan.witems:
        hit(it.dw)
    for dst in tr.weights._scatter_dst:
        hit(dst.data_ptr())
    for name, args in calls:
        if name == "savsr_grad_prep":
            for e in _entries(args, 5, 6):
                hit(e.dbias)
        elif name == "savsr_osa_fold_backward":
            for g in _entries(args, 3, 4):
                for f, _t in K.OsaGrads._fields_:
                    if f.startswith("d_"):
                        hit(getattr(g, f))
        elif name == "savsr_mask_backward_train":
            m = C.cast(args[2], C.POINTER(K.MaskTrain)).contents if not isinstance(args[2], K.MaskTrain) else args[2]
            for f in ("d_w4", "d_b4", "d_w7", "d_b7", "d_w11", "d_b11", "d_gamma"):
                hit(getattr(m, f))
            for l in range(4):
                hit(m.d_bn_w[l]); hit(m.d_bn_b[l])
        elif name == "savsr_ca_backward":
            for a in args[11:15]:
                hit(a)
    island = {n for n in G if n.startswith("upsample.") or n.startswith("tail.")}
    missing = sorted(set(G) - covered - island)
    assert not missing, f"{len(missing)} parameters without a gradient path, e.g. {missing[:6]}"
    assert len(G) == 707
