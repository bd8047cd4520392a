"""Clip-level driver: window gather, frame sharding across ranks, batched inference.

Every output frame of a clip is an independent forward over its own 7-frame LR window (the reference's
hot loop, lbasicsr/models/video_base_model.py:50-59, is ``for idx in range(rank, len(dataset),
world_size)``).  So multi-GPU inference shards *frames* across ranks with no collective on the data
path; only the optional gather of results uses ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch


def frame_window_indices(crt_idx: int, n_frames: int, num_frames: int = 7) -> List[int]:
    """Reflection-padded window around `crt_idx` (lbasicsr/data/data_util.py:63-112, padding='reflection')."""
    if num_frames % 2 != 1:
        raise ValueError("num_frames should be an odd number.")
    last = n_frames - 1
    pad = num_frames // 2
    if n_frames <= pad:     # the reflected index -i (or 2 * last - i) would leave the clip; the reference raises IndexError here
        raise ValueError(f"a clip of {n_frames} frames is too short for reflection padding of a {num_frames}-frame window")
    out = []
    for i in range(crt_idx - pad, crt_idx + pad + 1):
        if i < 0:
            i = -i
        elif i > last:
            i = 2 * last - i
        out.append(i)
    return out


def shard_frames(n_frames: int, rank: int, world_size: int, contiguous: bool = False) -> List[int]:
    """Frames of a clip owned by `rank`.  Default = the reference's rank-strided loop
    (video_base_model.py:50); `contiguous` gives equal contiguous chunks instead."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    if contiguous:
        per = (n_frames + world_size - 1) // world_size
        return list(range(min(rank * per, n_frames), min((rank + 1) * per, n_frames)))
    return list(range(rank, n_frames, world_size))


def gather_windows(clip: torch.Tensor, frames: Sequence[int], num_frames: int = 7) -> torch.Tensor:
    """clip [T, 3, h, w] -> windows [len(frames), num_frames, 3, h, w] (temporal halo = duplicated input)."""
    T = clip.shape[0]
    idx = torch.tensor([frame_window_indices(f, T, num_frames) for f in frames], dtype=torch.long, device=clip.device)
    return clip[idx.reshape(-1)].reshape(len(frames), num_frames, *clip.shape[1:])


def infer_clip(net: Callable[[torch.Tensor], torch.Tensor], clip: torch.Tensor, frames: Optional[Sequence[int]] = None,
               batch: int = 1, num_frames: int = 7, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run `net` on the windows of `frames` (default: all), `batch` windows per forward.
    Returns [len(frames), 3, H, W].  The last, ragged batch is run at its own size.
    If `out` (e.g. a pinned host tensor) is given, each batch is copied into it on a side stream while the next
    batch computes, and `out` is returned once all copies are enqueued behind the current stream."""
    frames = list(range(clip.shape[0])) if frames is None else list(frames)
    copy_stream = torch.cuda.Stream(clip.device) if (out is not None and clip.is_cuda) else None
    outs = []
    for i in range(0, len(frames), batch):
        chunk = frames[i:i + batch]
        y = net(gather_windows(clip, chunk, num_frames))
        if out is None:
            outs.append(y)
        elif copy_stream is None:
            out[i:i + len(chunk)].copy_(y)
        else:
            copy_stream.wait_stream(torch.cuda.current_stream(clip.device))
            with torch.cuda.stream(copy_stream):
                out[i:i + len(chunk)].copy_(y, non_blocking=True)
            y.record_stream(copy_stream)
    if out is not None:
        if copy_stream is not None:
            torch.cuda.current_stream(clip.device).wait_stream(copy_stream)
        return out
    if not outs:
        return torch.empty(0)
    return torch.cat(outs, 0)


def gather_outputs(local: torch.Tensor, frames: Sequence[int], n_frames: int, group=None) -> Optional[torch.Tensor]:
    """All-gather per-rank results [n_local, 3, H, W] into clip order [n_frames, 3, H, W] on every rank.
    Ranks may own different numbers of frames (ragged): results are padded to the longest shard."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = [None] * world
    dist.all_gather_object(counts, list(frames), group=group)
    longest = max(len(c) for c in counts)
    shape = list(local.shape[1:])
    shapes = [None] * world
    dist.all_gather_object(shapes, shape if local.numel() else None, group=group)
    shape = next(s for s in shapes if s is not None)
    pad = torch.zeros([longest] + shape, dtype=torch.float32, device=local.device)
    if local.numel():
        pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty([n_frames] + shape, dtype=torch.float32, device=local.device)
    for r in range(world):
        for j, f in enumerate(counts[r]):
            out[f] = bufs[r][j]
    return out
