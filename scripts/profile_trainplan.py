"""ncu target: a few eager native training steps at the cfg5 shape (4 x 7 x 3 x 64 x 64, x4)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import savsr_b200  # noqa: E402
from savsr_b200 import trainplan as TP  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = savsr_b200.SAVSR().to(dev)
tr = TP.NativeTrainer(net, use_graph=False)
lq = torch.rand(4, 7, 3, 64, 64, device=dev)
gt = torch.rand(4, 3, 256, 256, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    tr.step(lq, gt, (4, 4))
torch.cuda.synchronize()
