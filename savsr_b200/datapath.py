"""Device-side LR synthesis and clip evaluation.

The reference's test loop (lbasicsr/models/video_base_model.py:50-98 over lbasicsr/data/video_test_dataset.py:297-328) spends
its time on the CPU around the network: for every output frame it decodes 7 ground-truth PNGs, mod-crops them
(lbasicsr/data/transforms.py:47-69), converts to float RGB tensors and resizes them with torchvision's antialiased bicubic
Resize (lbasicsr/data/data_util.py:396-412), then converts the SR frame to uint8 and computes PSNR / SSIM with numpy / cv2.
Here the decoded uint8 frames of a clip are uploaded once and everything else -- crop, colour conversion, antialiased
bicubic downsample (bit-exact with the reference's), window gather, the network, tensor2img, PSNR-Y and SSIM-Y -- runs on the
GPU.  PNG decoding / encoding stays with the caller.  There is no CPU fallback.
"""
from __future__ import annotations

from math import floor
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

from . import _capi as K
from . import engine, postproc, sharding


# ------------------------------------------------------------------------------------------------ host integer logic
def cal_step(scale: float) -> int:
    """lbasicsr/data/transforms.py:31-44."""
    for step in (1, 2, 5, 10, 20, 50):
        if abs(scale * step - round(scale * step)) < 0.001:
            return step
    raise ValueError(f"unsupported scale {scale}: no step in (1, 2, 5, 10, 20, 50) makes scale * step an integer")


def as_mod_crop_size(h: int, w: int, scale: Sequence[float]) -> Tuple[int, int]:
    """Size kept by as_mod_crop (lbasicsr/data/transforms.py:47-69); the crop is the top-left region."""
    sh, sw = (scale, scale) if not isinstance(scale, (tuple, list)) else scale
    step_h, step_w = cal_step(sh), cal_step(sw)
    return round(floor(h / step_h / sh) * step_h * sh), round(floor(w / step_w / sw) * step_w * sw)


def lr_size(hc: int, wc: int, scale: Sequence[float]) -> Tuple[int, int]:
    """(round(h / scale_h), round(w / scale_w)), lbasicsr/data/data_util.py:398."""
    sh, sw = (scale, scale) if not isinstance(scale, (tuple, list)) else scale
    return round(hc / sh), round(wc / sw)


# ------------------------------------------------------------------------------------------------ resampling tables
_tables: Dict[Tuple[int, int, int], Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]] = {}


def aa_table(in_size: int, out_size: int, device: torch.device):
    """(xmin int32 [out], xsize int32 [out], weights float32 [out, taps], taps) of the antialiased bicubic resampler for one
    axis, built on the device once per (in, out) and cached."""
    key = (in_size, out_size, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _tables:
        ctx = engine.context(key[2])
        taps = ctx.lib.savsr_aa_max_taps(in_size, out_size)
        xmin = torch.empty(out_size, dtype=torch.int32, device=device)
        xsize = torch.empty(out_size, dtype=torch.int32, device=device)
        wts = torch.empty(out_size, taps, dtype=torch.float32, device=device)
        flag = torch.zeros(1, dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            K.check(ctx.lib.savsr_aa_table(ctx.handle, in_size, out_size, taps, xmin.data_ptr(), xsize.data_ptr(), wts.data_ptr(),
                                           flag.data_ptr(), torch.cuda.current_stream().cuda_stream))
        if int(flag.item()) != 0:
            raise RuntimeError(f"antialias table {in_size}->{out_size}: more than {taps} taps needed")
        _tables[key] = (xmin, xsize, wts, taps)
    return _tables[key]


def synthesize_lr(frames_bgr_u8: torch.Tensor, scale: Sequence[float], want_gt: bool = True
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """uint8 [T,H,W,3] BGR frames on a CUDA device (what cv2.imread returns, uploaded) ->
    (LR float32 [T,3,h,w] RGB, mod-cropped GT float32 [T,3,Hc,Wc] RGB or None), as read_img_seq(require_as_mod_crop=True) +
    arbitrary_scale_downsample(mode='torch') produce them (data_util.py:27-60, 371-420)."""
    if not frames_bgr_u8.is_cuda:
        raise RuntimeError("savsr_b200.datapath runs on CUDA only; there is no CPU fallback")
    if frames_bgr_u8.dtype != torch.uint8 or frames_bgr_u8.dim() != 4 or frames_bgr_u8.shape[-1] != 3:
        raise ValueError(f"expected uint8 [T,H,W,3], got {frames_bgr_u8.dtype} {tuple(frames_bgr_u8.shape)}")
    frames = frames_bgr_u8.contiguous()
    T, H, W, _ = frames.shape
    hc, wc = as_mod_crop_size(H, W, scale)
    if hc <= 0 or wc <= 0:
        raise ValueError(f"as_mod_crop of a {H}x{W} frame at scale {tuple(scale)} is empty")
    oh, ow = lr_size(hc, wc, scale)
    dev = frames.device
    ctx = engine.context(dev.index if dev.index is not None else torch.cuda.current_device())
    tw = aa_table(wc, ow, dev) if ow != wc else (None, None, None, 0)
    th = aa_table(hc, oh, dev) if oh != hc else (None, None, None, 0)
    tmp = torch.empty(T, 3, hc, ow, dtype=torch.float32, device=dev)
    lr = torch.empty(T, 3, oh, ow, dtype=torch.float32, device=dev)
    gt = torch.empty(T, 3, hc, wc, dtype=torch.float32, device=dev) if want_gt else None
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    with torch.cuda.device(dev):
        K.check(ctx.lib.savsr_lr_synthesize(ctx.handle, frames.data_ptr(), T, H, W, hc, wc, oh, ow,
                                            ptr(tw[0]), ptr(tw[1]), ptr(tw[2]), tw[3], ptr(th[0]), ptr(th[1]), ptr(th[2]), th[3],
                                            tmp.data_ptr(), lr.data_ptr(), ptr(gt), torch.cuda.current_stream().cuda_stream))
    return lr, gt


def resize_aa_bicubic(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """T.Resize(size, BICUBIC, antialias=True) of float32 frames [n,3,H,W] on the device: the reference's post-process for an
    SR result whose size differs from the ground truth (lbasicsr/models/sr_model.py:291-304, row f4)."""
    if not x.is_cuda:
        raise RuntimeError("savsr_b200.datapath runs on CUDA only; there is no CPU fallback")
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected [n,3,H,W], got {tuple(x.shape)}")
    x = x.float().contiguous()
    n, _, H, W = x.shape
    oh, ow = int(size[0]), int(size[1])
    if (oh, ow) == (H, W):
        return x
    dev = x.device
    ctx = engine.context(dev.index if dev.index is not None else torch.cuda.current_device())
    tw = aa_table(W, ow, dev) if ow != W else (None, None, None, 0)
    th = aa_table(H, oh, dev) if oh != H else (None, None, None, 0)
    tmp = torch.empty(n, 3, H, ow, dtype=torch.float32, device=dev)
    out = torch.empty(n, 3, oh, ow, dtype=torch.float32, device=dev)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    with torch.cuda.device(dev):
        K.check(ctx.lib.savsr_resize_aa(ctx.handle, x.data_ptr(), n, H, W, oh, ow, ptr(tw[0]), ptr(tw[1]), ptr(tw[2]), tw[3],
                                        ptr(th[0]), ptr(th[1]), ptr(th[2]), th[3], tmp.data_ptr(), out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
    return out


# ------------------------------------------------------------------------------------------------ one clip, end to end
def evaluate_clip(net: Callable[[torch.Tensor], torch.Tensor], frames_bgr_u8: torch.Tensor, scale: Sequence[float],
                  frames: Optional[Sequence[int]] = None, batch: int = 17, num_frames: int = 7,
                  want_images: bool = True) -> Dict[str, torch.Tensor]:
    """The reference's per-clip test loop (video_base_model.py:50-98 with the YAML metrics psnr_y / ssim_y) on the device:
    LR synthesis -> 7-frame windows (reflection padding) -> net -> uint8 BGR images + PSNR-Y + SSIM-Y against the
    mod-cropped ground truth.  `net` must already be set to `scale`.  `frames`: output frames owned by this rank
    (default all; see sharding.shard_frames).  Returns {"images": uint8 [n,H,W,3] or None, "psnr_y": float64 [n],
    "ssim_y": float64 [n], "sr": float32 [n,3,H,W]}."""
    lr, gt = synthesize_lr(frames_bgr_u8, scale, want_gt=True)
    idx = list(range(lr.shape[0])) if frames is None else list(frames)
    # the ground-truth rows of this rank's frames, selected BEFORE the forwards are enqueued: building a device index is a blocking copy,
    # and behind the forwards it would hold the host until they have finished (an idle GPU before the post-processing launches)
    if idx == list(range(gt.shape[0])):
        gt_sel = gt
    else:
        gt_sel = gt.index_select(0, torch.tensor(idx, dtype=torch.long).to(gt.device)) if idx else gt[:0]
    sr = sharding.infer_clip(net, lr, idx, batch=batch, num_frames=num_frames)
    if tuple(sr.shape) != tuple(gt_sel.shape):            # arbitrary-scale BI post-process, sr_model.py:291-294
        sr = resize_aa_bicubic(sr, (gt_sel.shape[-2], gt_sel.shape[-1]))
    images, psnr = postproc.tensor2img_psnr(sr, gt_sel, want_image=want_images)
    ssim = postproc.ssim_y(sr, gt_sel)
    return dict(images=images, psnr_y=psnr, ssim_y=ssim, sr=sr)


# ------------------------------------------------------------------------------------------------ clip prefetch
class ClipPrefetcher:
    """Uploads the NEXT clip's decoded uint8 frames on a side stream while the current clip is being processed -- the role
    of the reference's CUDAPrefetcher (lbasicsr/data/prefetch_dataloader.py:84-125) for whole clips.  `clips` yields
    (name, uint8 [T,H,W,3] BGR host tensor or ndarray); pinned host tensors make the copy truly asynchronous.

        for name, frames in ClipPrefetcher(clips, device):      # frames: uint8 CUDA tensor, ready on the current stream
            res = evaluate_clip(net, frames, scale)
    """

    def __init__(self, clips, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("savsr_b200.datapath runs on CUDA only; there is no CPU fallback")
        self._it = iter(clips)
        self._stream = torch.cuda.Stream(self.device)
        self._next = None
        self._preload()

    def _preload(self):
        try:
            name, frames = next(self._it)
        except StopIteration:
            self._next = None
            return
        host = frames if torch.is_tensor(frames) else torch.from_numpy(frames)
        if host.dtype != torch.uint8 or host.dim() != 4 or host.shape[-1] != 3:
            raise ValueError(f"clip {name!r}: expected uint8 [T,H,W,3], got {host.dtype} {tuple(host.shape)}")
        with torch.cuda.stream(self._stream):
            dev = host.to(self.device, non_blocking=True)
        self._next = (name, dev, host)           # keep the host tensor alive until the copy has been consumed

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
        name, dev, _host = self._next
        dev.record_stream(torch.cuda.current_stream(self.device))
        self._preload()
        return name, dev


def evaluate_clips(net: Callable[[torch.Tensor], torch.Tensor], clips, scale: Sequence[float], device, batch: int = 17,
                   num_frames: int = 7, want_images: bool = False) -> Dict[str, Dict[str, float]]:
    """Per-clip averages like the reference's validation summary (video_base_model.py:132-146): {clip: {"psnr_y", "ssim_y"}}
    plus the dataset average under the key "average".  Uploads of clip i+1 overlap the processing of clip i."""
    out: Dict[str, Dict[str, float]] = {}
    for name, frames in ClipPrefetcher(clips, device):
        res = evaluate_clip(net, frames, scale, batch=batch, num_frames=num_frames, want_images=want_images)
        out[name] = dict(psnr_y=float(res["psnr_y"].mean()), ssim_y=float(res["ssim_y"].mean()))
    if out:
        out["average"] = dict(psnr_y=sum(v["psnr_y"] for v in out.values()) / len(out),
                              ssim_y=sum(v["ssim_y"] for v in out.values()) / len(out))
    return out
