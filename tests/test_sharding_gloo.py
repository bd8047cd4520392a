"""CPU: frame sharding across ranks (world_size 2, gloo) and the window gather, against single-process results."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from savsr_b200 import sharding


def fake_net(x):
    """Stand-in for the SR network: deterministic function of the 7-frame window -> [b, 3, 2h, 2w]."""
    b, t, c, h, w = x.shape
    wts = torch.arange(1, t + 1, dtype=x.dtype).view(1, t, 1, 1, 1)
    y = (x * wts).sum(1)
    return torch.nn.functional.interpolate(y, scale_factor=2, mode="nearest")


def test_reflection_window_indices():
    assert sharding.frame_window_indices(0, 34) == [3, 2, 1, 0, 1, 2, 3]          # data_util.py:63-112 'reflection'
    assert sharding.frame_window_indices(33, 34) == [30, 31, 32, 33, 32, 31, 30]
    assert sharding.frame_window_indices(10, 34) == [7, 8, 9, 10, 11, 12, 13]
    assert sharding.frame_window_indices(0, 5, 5) == [2, 1, 0, 1, 2]
    assert sharding.frame_window_indices(1, 4) == [2, 1, 0, 1, 2, 3, 2]               # shortest clip a 7-frame window supports
    with pytest.raises(ValueError, match="too short"):
        sharding.frame_window_indices(0, 3)                                         # reflection would index frame 3 of 3


def test_window_indices_match_reference_kat():
    """Known answers produced by the reference's own generate_frame_indices (scripts/make_golden.py --window-only): every frame of clips
    of 4..41 frames (7-frame windows) and 3..41 frames (5-frame windows), reflection padding."""
    import numpy as np
    kat = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "window_kat.npz"))
    assert len(kat.files) == 38 + 39
    for key in kat.files:
        nf, T = int(key.split(".")[0][2:]), int(key.split(".")[1][1:])
        want = kat[key]
        got = np.array([sharding.frame_window_indices(i, T, nf) for i in range(T)])
        assert np.array_equal(got, want), key


def test_shard_frames_partition():
    for n, world in ((34, 1), (34, 2), (34, 8), (5, 2), (3, 8), (0, 2)):
        for contiguous in (False, True):
            parts = [sharding.shard_frames(n, r, world, contiguous) for r in range(world)]
            assert sorted(f for p in parts for f in p) == list(range(n))
    assert sharding.shard_frames(34, 1, 8) == [1, 9, 17, 25, 33]                 # video_base_model.py:50 rank-strided loop
    with pytest.raises(ValueError):
        sharding.shard_frames(10, 2, 2)


def test_infer_clip_ragged_batches():
    clip = torch.rand(11, 3, 6, 8)
    full = sharding.infer_clip(fake_net, clip, batch=11)
    for b in (1, 4, 5):
        assert torch.equal(sharding.infer_clip(fake_net, clip, batch=b), full)
    assert sharding.infer_clip(fake_net, clip, frames=[], batch=4).numel() == 0
    out = torch.empty(11, 3, 12, 16)
    assert sharding.infer_clip(fake_net, clip, batch=4, out=out) is out and torch.equal(out, full)


def test_gather_windows_index_cache():
    """The gather index is built once per (clip length, frame list, window, device) and kept on the device (building it per call is a
    blocking host-to-device copy); different clips, frame lists and lengths must not alias."""
    sharding._WINDOW_INDEX_CACHE.clear()
    a, b = torch.rand(9, 3, 4, 5), torch.rand(9, 3, 4, 5)
    for clip in (a, b, a):
        for frames in ([0, 8, 4], [4], list(range(9))):
            got = sharding.gather_windows(clip, frames)
            want = torch.stack([clip[sharding.frame_window_indices(f, 9)] for f in frames])
            assert torch.equal(got, want)
    assert len(sharding._WINDOW_INDEX_CACHE) == 3                       # one entry per frame list, shared by both clips
    c = torch.rand(12, 3, 4, 5)                                         # same frame list, longer clip: another reflection pattern
    assert torch.equal(sharding.gather_windows(c, [0, 8, 4])[1], c[sharding.frame_window_indices(8, 12)])
    assert torch.equal(sharding.gather_windows(a, [8], num_frames=5)[0], a[sharding.frame_window_indices(8, 9, 5)])
    assert len(sharding._WINDOW_INDEX_CACHE) == 5
    for i in range(300):                                                # bounded
        sharding.gather_windows(c, [i % 12, (i // 12) % 12, 3])
    assert len(sharding._WINDOW_INDEX_CACHE) <= 256


def _worker(rank, world, port, n_frames, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        clip = torch.rand(n_frames, 3, 6, 8, generator=torch.Generator().manual_seed(7))
        ref = sharding.infer_clip(fake_net, clip, batch=3)
        ok = True
        for contiguous in (False, True):
            mine = sharding.shard_frames(n_frames, rank, world, contiguous)
            local = sharding.infer_clip(fake_net, clip, frames=mine, batch=2) if mine else torch.empty(0, 3, 12, 16)
            # gather to rank 0 (one pre-sized dist.gather, ragged shards padded): rank 0 gets the clip, the others nothing
            full = sharding.gather_outputs(local, n_frames, rank, world, dst=0, contiguous=contiguous)
            ok &= (full is None) if rank != 0 else bool(torch.equal(full, ref))
            # every rank gets the clip (all_gather_into_tensor), through the asynchronous interface
            work, finish = sharding.gather_outputs(local, n_frames, rank, world, dst=None, contiguous=contiguous, async_op=True)
            work.wait()
            ok &= bool(torch.equal(finish(), ref))
            # the reference's metric reduction (video_base_model.py:106-113): rows of the owned frames, summed on rank 0
            rows = torch.stack([local.flatten(1).mean(1), local.flatten(1).amax(1)], 1) if mine else torch.empty(0, 2)
            red = sharding.reduce_metrics(rows, mine, n_frames)
            if rank == 0:
                want = torch.stack([ref.flatten(1).mean(1), ref.flatten(1).amax(1)], 1)
                ok &= bool(torch.allclose(red, want, atol=1e-6))
        with pytest.raises(ValueError):
            sharding.gather_outputs(torch.empty(n_frames + 1, 1), n_frames, rank, world)
        ret[rank] = ok and len(sharding.shard_frames(n_frames, rank, world)) == len(range(rank, n_frames, world))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5, 8])
def test_two_rank_frame_sharding_and_gather(n_frames):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + n_frames
    mp.spawn(_worker, args=(world, port, n_frames, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
