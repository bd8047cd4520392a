"""Row f2 on the CPU: the oracle's restatement of the reference's LR synthesis (as_mod_crop + antialiased bicubic Resize)
against known answers produced by the reference itself (scripts/make_golden.py --lr-only), and the host-side integer logic
of savsr_b200.datapath against both."""
import os

import numpy as np
import pytest

from oracle import lr_synthesis as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["x4", "x2p7", "x1p5x4", "x3p9", "x1p2x1p7", "x7p3x5p1"]


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLDEN, "lr_kat.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_lr_bit_exact_with_reference(kat, name):
    frames, scale = kat[f"{name}.frames"], tuple(float(s) for s in kat[f"{name}.scale"])
    lr, gt = L.synthesize_lr(frames, scale)
    assert tuple(gt.shape[-2:]) == tuple(int(v) for v in kat[f"{name}.crop"])
    assert lr.shape == kat[f"{name}.lr"].shape
    assert np.array_equal(lr, kat[f"{name}.lr"])          # bit-exact: same taps, same accumulation order


def test_as_mod_crop_sizes(kat):
    from savsr_b200 import datapath
    for h, w, sh, sw, hc, wc in kat["crop_sweep"]:
        scale = (float(sh), float(sw))
        assert L.as_mod_crop_size(int(h), int(w), scale) == (int(hc), int(wc))
        assert datapath.as_mod_crop_size(int(h), int(w), scale) == (int(hc), int(wc))
        assert datapath.lr_size(int(hc), int(wc), scale) == L.lr_size(int(hc), int(wc), scale)


def test_weights_partition_of_unity_and_support():
    for in_size, out_size in ((720, 180), (576, 213), (63, 23), (40, 40), (51, 30)):
        for xmin, ws in L.aa_bicubic_weights(in_size, out_size):
            assert 0 <= xmin and xmin + len(ws) <= in_size
            assert abs(float(ws.astype(np.float64).sum()) - 1.0) < 1e-6


def test_unsupported_scale_raises():
    with pytest.raises(ValueError):
        L.cal_step(3.14159)


def test_datapath_refuses_cpu_tensors():
    import torch
    from savsr_b200 import datapath
    with pytest.raises(RuntimeError, match="CUDA only"):
        datapath.synthesize_lr(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), (2, 2))
