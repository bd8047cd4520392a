#!/usr/bin/env python
"""torchrun check of data-parallel training through the MODULE path: SAVSR wrapped in DistributedDataParallel exactly as
lbasicsr/models/base_model.py:98-99 does, its train-mode forward = the native launch list behind one autograd node, the caller's loss and
torch.optim.Adam.  Different data per rank; after every step the parameters must be identical on all ranks (DDP's gradient all-reduce saw
the gradients the native backward returned) and must have moved.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_train_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import savsr_b200  # noqa: E402
from savsr_b200 import train as T  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = savsr_b200.SAVSR().to(dev)
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, betas=(0.9, 0.99))
    gen = torch.Generator().manual_seed(100 + rank)
    p0 = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    losses = []
    for it, scale in enumerate([(2, 2), (1.5, 4), (2, 2), (1.5, 4)]):
        lq = torch.rand(2, 7, 3, 32, 32, generator=gen).to(dev)
        gt = torch.rand(2, 3, round(32 * scale[0]), round(32 * scale[1]), generator=gen).to(dev)
        net.set_scale(scale)
        model.train()
        opt.zero_grad()
        loss = T.charbonnier(model(lq), gt)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        flat = torch.cat([p.detach().flatten() for p in net.parameters()])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        diff = max(float((b - both[0]).abs().max()) for b in both)
        assert diff == 0.0, f"step {it}: parameters differ across ranks by {diff}"
    moved = float((flat - p0).abs().max())
    assert moved > 0 and all(l == l for l in losses), (moved, losses)
    assert net.__dict__.get("_train_state") is not None, "the native train-mode path was not taken"
    if rank == 0:
        print(f"DDP_OK world={world} losses={losses} max parameter change {moved:.3e}; parameters identical on all ranks after every step")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
