"""A recording stand-in for libsavsr_sm100 so the host-side plan logic can be tested without a GPU.
Test infrastructure only: every compute entry point just records its arguments and returns 0."""
import contextlib
from types import SimpleNamespace

import torch

from savsr_b200 import _capi as K
from savsr_b200 import engine


class FakeLib:
    def __init__(self):
        self.calls = []

    def savsr_packed_weight_bytes(self, co, ci, ks):
        return co * ci * ks * ks * 2

    def __getattr__(self, name):
        if not name.startswith("savsr_"):
            raise AttributeError(name)

        def fn(*args):
            self.calls.append((name, args))
            return 0
        return fn


class FakeArena:
    def __init__(self, ctx, base_ptr, nslots, batch, height, width):
        self.handle = SimpleNamespace(nslots=nslots, batch=batch, height=height, width=width, base=base_ptr)
        self.nslots, self.batch, self.height, self.width = nslots, batch, height, width
        self.tiles = ((width + 7) // 8) * ((height + 15) // 16)


@contextlib.contextmanager
def mocked_engine():
    lib = FakeLib()
    ctx = SimpleNamespace(lib=lib, handle="ctx", sm_count=148, set_format=lambda f: None, set_option=lambda o, v: None)
    saved = (engine.context, K.Arena, K.check, torch.cuda.current_stream, torch.cuda.device, torch.cuda.Stream, torch.cuda.stream)
    engine.context = lambda idx: ctx
    K.Arena = FakeArena
    K.check = lambda rc: None
    fake_stream = lambda *a, **k: SimpleNamespace(cuda_stream=0, synchronize=lambda: None, wait_stream=lambda s: None)   # noqa: E731
    torch.cuda.current_stream = fake_stream
    torch.cuda.Stream = fake_stream
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.device = lambda d: contextlib.nullcontext()
    try:
        yield lib
    finally:
        engine.context, K.Arena, K.check, torch.cuda.current_stream, torch.cuda.device, torch.cuda.Stream, torch.cuda.stream = saved


class CpuPlan(engine.Plan):
    """engine.Plan with the device check relaxed (host logic only; no kernel ever runs)."""

    def __init__(self, params, batch, h, w, scale, **kw):
        dev = torch.device("cpu")
        real_type = torch.device

        class _D:  # quacks like a cuda device for the constructor's guard
            type, index = "cuda", 0
        self._fake_dev = _D()
        orig_build = self._build
        engine.Plan.__init__.__globals__  # keep linters quiet
        # bypass guard by calling the pieces of __init__ manually
        self.device = dev
        self.ctx = engine.context(0)
        self.lib = self.ctx.lib
        self.impl = K.IMPL_NAMES[kw.get("conv_impl", "tap")]
        self.fmt = K.FMT_NAMES[kw.get("precision", "bf16")]
        self.B, self.h, self.w, self.t = batch, h, w, 7
        self.scale = engine.normalize_scale(scale)
        self.hp, self.wp = h + (h & 1), w + (w & 1)
        self.H, self.W = engine.get_hw(h, w, self.scale)
        self.P = params
        self._keep, self.ops, self.op_meta, self.n_launches = [], [], [], 0
        self.taps, self.tap_bufs, self.graph = set(kw.get("taps", ())), {}, None
        self._pool_cache, self._osa_cache = {}, {}
        self.store, self.nbytes, self._pack_src = kw.get("store") or engine.WeightStore(), 0, []
        self.cplan, self._recorded = None, 0            # (the C-side launch list is exercised on the GPU: tests/gpu_checks.check_c_plan)
        with torch.no_grad():
            orig_build()
