"""CPU: the drop-in boundary -- state_dict layout, registry, C-ABI symbol table, loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

import savsr_b200
from oracle.state_dict_fixture import make_state_dict, state_dict_spec
from savsr_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
YAML_KWARGS = dict(num_in_ch=3, num_feat=64, num_frame=7, slid_win=3, fusion_win=5, interval=0, w1_num_block=4, w2_num_block=2,
                   n_resgroups=4, n_resblocks=8, center_frame_idx=None)     # options/test/SAVSR/test_SAVSR_Vid4_asBI.yml:830-842


@pytest.fixture(scope="module")
def net():
    return savsr_b200.build_network(dict(type="SAVSR", **YAML_KWARGS))


def test_registry_and_constructor(net):
    assert "SAVSR" in savsr_b200.ARCH_REGISTRY
    assert type(net).__name__ == "SAVSR" and net.scale == (4, 4)
    with pytest.raises(KeyError):
        savsr_b200.ARCH_REGISTRY.get("NoSuchArch")
    with pytest.raises(NotImplementedError):
        savsr_b200.SAVSR(interval=1)


def test_state_dict_layout_matches_reference(net):
    spec = state_dict_spec()                 # pinned to the reference by scripts/make_golden.py (strict load + key order)
    sd = net.state_dict()
    assert list(sd) == [k for k, _, _ in spec] and len(sd) == 791
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(p.numel() for p in net.parameters()) == 18890044
    net.load_state_dict(make_state_dict(1), strict=True)
    assert "SAVSR(" in str(net)


def test_set_scale_and_hw(net):
    net.set_scale((1.5, 4))
    assert net.scale == (1.5, 4)
    assert savsr_b200.get_HW(33, 35, (1.5, 1.5)) == (50, 52)      # python round: half to even
    assert savsr_b200.get_HW(144, 180, 4) == (576, 720)


def test_no_cpu_fallback(net):
    net.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 7, 3, 8, 8))
    with pytest.raises(ValueError):
        net.plan_for(torch.zeros(1, 5, 3, 8, 8))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "savsr_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(savsr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = _capi.load()                        # raises if the .so is missing or lacks a bound symbol
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/savsr_b200.h but not exported"
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    assert lib.savsr_abi_version() == _capi.ABI_VERSION == 4


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_capi.ConvGroup) == 104 and _capi.ConvGroup.weight.offset == 48 and _capi.ConvGroup.src_channels.offset == 96
    assert _capi.OsaParams.pool.offset == 16 + 16 * 8 and ctypes.sizeof(_capi.SatuWeights) == 96
    # training entry points (train_ops.cu / train_attn.cu carry the matching static_asserts on the C side)
    assert ctypes.sizeof(_capi.Axpby) == 20 and ctypes.sizeof(_capi.Nchw3) == 8
    assert ctypes.sizeof(_capi.GradPrep) == 72 and _capi.GradPrep.cscale.offset == 24 and _capi.GradPrep.dbias.offset == 64
    assert ctypes.sizeof(_capi.PackChunk) == 40 and ctypes.sizeof(_capi.WgradItem) == 48 and _capi.WgradItem.sample_stride.offset == 40
    assert ctypes.sizeof(_capi.MaskTrain) == 456 and _capi.MaskTrain.gamma.offset == 176 and _capi.MaskTrain.dmask.offset == 400
    assert ctypes.sizeof(_capi.OsaTrain) == 56 and _capi.OsaTrain.state.offset == 40 and ctypes.sizeof(_capi.OsaGrads) == 160


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_context_creation_fails_loudly_without_gpu():
    with pytest.raises(_capi.SavsrError):
        _capi.Context(0)
    lib = _capi.load()
    assert lib.savsr_arena_bytes(2, 3, 10, 12) == 2 * 3 * 10 * 12 * 64 * 2
    assert lib.savsr_packed_weight_bytes(64, 192, 3) == 64 * 192 * 9 * 2
    assert lib.savsr_ctx_sm_count(None) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/lbasicsr"), reason="reference tree only exists in the build container")
def test_overlay_serves_the_new_arch_to_the_reference_registry():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from savsr_b200 import overlay; overlay.install('/root/reference')\n"
            "import lbasicsr\n"
            "from lbasicsr.archs import build_network\n"
            "net = build_network(dict(type='SAVSR', num_in_ch=3, num_feat=64, num_frame=7, slid_win=3, fusion_win=5, interval=0,"
            " w1_num_block=4, w2_num_block=2, n_resgroups=4, n_resblocks=8, center_frame_idx=None))\n"
            "import savsr_b200.engine\n"
            "assert type(net).__module__ == 'lbasicsr.archs.savsr_arch' and hasattr(net, 'plan_for'), type(net)\n"
            "print('OVERLAY_OK', len(net.state_dict()))\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "OVERLAY_OK 791" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1_x2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "hr_mpix_per_s" and line["unit"] == "HR Mpix/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "HR Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == "cfg1_x2"
    import bench
    assert line["config"] == bench.workload_config("cfg1_x2", 1)      # both arms print the same config object


def test_bench_our_arm_needs_a_gpu():
    """No CPU fallback: without CUDA our arm exits with an error instead of measuring something else."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
