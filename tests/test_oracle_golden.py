"""CPU: the oracle (oracle/savsr_oracle.py) against (1) golden vectors produced by the UNMODIFIED reference
(scripts/make_golden.py, run in the build container) and (2) the known-answer hashes of SURVEY.md A.3."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import savsr_oracle as O
from oracle.state_dict_fixture import make_input, make_state_dict, state_dict_spec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if "fingerprint" not in p and "metrics" not in p and "_kat" not in p and not os.path.basename(p).startswith("train_"))
TRAIN_CASES = sorted(glob.glob(os.path.join(GOLDEN, "train_*.npz")))


def sha12(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


def test_golden_files_present():
    assert len(CASES) == 5


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference_outputs(path):
    g = np.load(path)
    b, h, w = int(g["b"]), int(g["h"]), int(g["w"])
    scale = tuple(float(s) if float(s) != int(s) else int(s) for s in g["scale"])
    sd = make_state_dict(int(g["sd_seed"]))
    x = make_input(b, h, w, int(g["in_seed"]))
    probes = {}
    y = O.forward(sd, x, scale, probes)
    ref = torch.from_numpy(g["out"])
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) < 2e-5            # fp32 CPU vs fp32 CPU, different op grouping
    # per-stage probes recorded by hooks on the reference modules
    for key in [k[len("probe."):-len(".sample")] for k in g.files if k.endswith(".sample")]:
        t = probes[key].detach().float().contiguous().flatten()
        idx = torch.linspace(0, t.numel() - 1, steps=min(257, t.numel())).long()
        samp = torch.from_numpy(g[f"probe.{key}.sample"])
        assert tuple(g[f"probe.{key}.shape"]) == tuple(probes[key].shape), key
        assert float((t[idx] - samp).abs().max()) < 5e-4 * max(1.0, float(g[f"probe.{key}.absmax"])), key
    # index path: bit-exact
    H, W = O.get_hw(h, w, scale)
    assert sha12(O.satu_mlp_input(h, w, scale).numpy()) == str(g["satu_mlp_input_sha1"])[:12]
    grid0 = O.satu_grid(h, w, scale, probes["satu_offset"])[0].numpy()
    grid1 = O.satu_grid(h, w, scale, probes["satu_st_offset"])[0].numpy()
    assert np.abs(grid0 - g["grid0"]).max() < 1e-6 and np.abs(grid1 - g["grid1"]).max() < 1e-6
    zero = O.satu_grid(h, w, scale, torch.zeros(1, 2, H, W))[0].numpy()
    assert np.array_equal(zero[0, :, 0], O.satu_base_norm(W, w, scale[1]))


def _train_case(g):
    b, h, w, sd_seed, in_seed = (int(v) for v in g["dims"])
    scale = tuple(float(s) if float(s) != int(s) else int(s) for s in g["scale"])
    return b, h, w, sd_seed, in_seed, scale


@pytest.mark.parametrize("path", TRAIN_CASES, ids=[os.path.basename(p)[:-4] for p in TRAIN_CASES])
def test_oracle_train_mode_matches_reference_gradients(path):
    """Row f1: the oracle's train-mode restatement (BatchNorm on batch statistics) + Charbonnier + autograd against the loss and the
    gradients the UNMODIFIED reference produced in train() mode with its own CharbonnierLoss (scripts/make_golden.py --train-only):
    loss to 1e-6, every parameter's gradient norm, the projection of the whole gradient on a random direction, and the small gradient
    tensors element by element.  This pins the checker of the native training path (tests/gpu_checks.py:check_trainplan)."""
    g = np.load(path)
    b, h, w, sd_seed, in_seed, scale = _train_case(g)
    sd = make_state_dict(sd_seed)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    x = make_input(b, h, w, in_seed)
    gt = torch.from_numpy(g["gt"])
    O.BN_TRAIN = True
    try:
        out = O.forward(sd, x, scale)
        loss = torch.sqrt((out - gt) ** 2 + 1e-12).mean()           # CharbonnierLoss(loss_weight=1, reduction='mean'), basic_loss.py:22-24
        loss.backward()
    finally:
        O.BN_TRAIN = False
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    t = out.detach().flatten()
    idx = torch.linspace(0, t.numel() - 1, steps=min(257, t.numel())).long()
    assert tuple(g["out.shape"]) == tuple(out.shape) and float((t[idx] - torch.from_numpy(g["out.sample"])).abs().max()) < 1e-5
    names = [str(n) for n in g["names"]]
    assert names == [k for k, v in sd.items() if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    norms, present = g["grad_norm"], g["grad_present"]
    assert present.all() and len(names) == 707                      # every parameter of the net takes part in the step
    big = float(norms.max())
    gen = torch.Generator().manual_seed(4242)
    proj = 0.0
    for k, n_ref in zip(names, norms):
        gr = sd[k].grad
        assert gr is not None, k
        r = torch.randn(sd[k].shape, generator=gen, dtype=torch.float64)
        proj += float((gr.double() * r).sum())
        # (gradients upstream of a train-mode BatchNorm are ill-conditioned -- a conv bias there has a true gradient of zero -- hence the floor)
        assert abs(float(gr.double().norm()) - float(n_ref)) < 2e-3 * max(float(n_ref), 1e-2 * big), (k, float(gr.norm()), float(n_ref))
        if "grad." + k in g.files:
            ref = torch.from_numpy(g["grad." + k])
            assert float((gr - ref).norm()) < 2e-3 * max(float(ref.norm()), 1e-2 * big), k
    gnorm = float(np.sqrt((norms ** 2).sum()))
    assert abs(proj - float(g["grad_proj"])) < 1e-3 * gnorm * 10      # |<g, r>| ~ |g|; fp32 summation-order noise only


def test_train_goldens_present():
    assert len(TRAIN_CASES) == 2


def test_default_init_fingerprint_recorded():
    f = np.load(os.path.join(GOLDEN, "cfg1_default_init_fingerprint.npz"))
    assert abs(float(f["param_sum"]) - 466.810371) < 1e-4
    assert abs(float(f["out_mean"]) - 0.45701084) < 1e-6
    assert float(f["oracle_maxabs"]) < 1e-6                     # oracle == reference on BASELINE cfg 1 (64x64, x2)


# ---- SURVEY.md appendix A.3 known-answer tests -------------------------------------------------------------
def test_get_hw_round_half_even():
    assert O.get_hw(33, 35, (1.5, 1.5)) == (50, 52)
    assert O.get_hw(65, 144, (2.7, 2.7)) == (176, 389)
    assert O.get_hw(63, 63, (2.7, 2.7)) == (170, 170)
    assert O.get_hw(144, 180, (1.5, 4)) == (216, 720)


@pytest.mark.parametrize("n,s,nout,h12", [(64, 2, 128, "2bd9380ce50a"), (64, 1.5, 96, "6d6647816902"), (64, 2.7, 173, "9f45df9829b0"),
                                           (144, 1.5, 216, "418bcb97ae37"), (180, 4, 720, "021ddcc40e36"), (144, 2.7, 389, "1950fc1a4e73")])
def test_cell_index_kat(n, s, nout, h12):
    assert round(n * s) == nout
    cell = O.satu_cell(nout, s)
    assert sha12(cell) == h12 and cell[-1] == n - 1


@pytest.mark.parametrize("nout,s,h12", [(128, 2, "09a4fd75bffb"), (96, 1.5, "2a7cd8ea30b5"), (256, 4, "deed35b81c0d"), (173, 2.7, "dc73b1ab9e16"),
                                         (216, 1.5, "ecf48854a67f"), (576, 4, "c2caef33947a"), (720, 4, "1a7ef9781faa"),
                                         (389, 2.7, "1c06f3a13d28"), (486, 2.7, "9edc80a3cab8")])
def test_relative_coordinate_kat(nout, s, h12):
    assert sha12(O.satu_rel_coord(nout, s)) == h12


@pytest.mark.parametrize("n,s,h12", [(64, 2, "75a53550045b"), (180, 4, "0bdc27a18b13"), (144, 1.5, "b63dbb3b7a5c"), (144, 4, "5ee07c3fc281"),
                                      (180, 2.7, "5af2de225a9a"), (144, 2.7, "8f3dd66bd386")])
def test_base_corner_kat(n, s, h12):
    c = O.satu_base_corner(round(n * s), n, s)
    assert sha12(c) == h12 and c.min() == -1 and c.max() <= n - 1


def test_expert_mix_is_not_the_single_sum_shortcut():
    sd = make_state_dict(3)
    f = torch.randn(1, 64, 5, 6)
    r = torch.rand(1, 4, 5, 6)
    good = O.satu_expert_mix(sd, "upsample", f, r)
    wc, we = sd["upsample.weight_compress"].flatten(2), sd["upsample.weight_expand"].flatten(2)
    dense_c = torch.einsum("ehw,ekc->hwkc", r[0], wc)
    dense_e = torch.einsum("ehw,eck->hwck", r[0], we)
    ref = torch.einsum("hwck,hwkd,bdhw->bchw", dense_e, dense_c, f) + f       # the reference's materialised form (353-370)
    assert float((good - ref).abs().max()) < 1e-5


def test_state_dict_fixture_is_deterministic():
    a, b = make_state_dict(5), make_state_dict(5)
    assert list(a) == [k for k, _, _ in state_dict_spec()] and len(a) == 791
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["tail.weight"], make_state_dict(6)["tail.weight"])


def test_metric_chain_matches_reference_kat():
    """tensor2img + PSNR-Y of the oracle vs values produced by the reference's own img_util / psnr_ssim (make_golden.py)."""
    k = np.load(os.path.join(GOLDEN, "metrics_kat.npz"))
    sr, gt = torch.from_numpy(k["sr"]), torch.from_numpy(k["gt"])
    for i in range(sr.shape[0]):
        assert hashlib.sha1(np.ascontiguousarray(O.tensor2img(sr[i])).tobytes()).hexdigest() == str(k["img_sha1"][i])
        assert abs(O.psnr_y(sr[i], gt[i]) - float(k["psnr_y"][i])) < 1e-9
        assert abs(O.ssim_y(sr[i], gt[i]) - float(k["ssim_y"][i])) < 1e-12      # reference: cv2.filter2D in float64
    assert O.psnr_y(gt[0], gt[0]) == float("inf")
