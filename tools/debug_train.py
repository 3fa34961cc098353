"""Per-tensor distance to the fp64 oracle gradient: the native training step vs torch's fp32 CPU evaluation."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import i3d_state_dict, quiet, rel_err
from interpreting_video_features_b200.train import I3DTrainer
from oracle import synthetic, train_oracle
dev = torch.device("cuda:0")
sd, _ = quiet(i3d_state_dict, 174)
x = synthetic.clips(2, t=16, h=96, w=96)
target = torch.tensor([5, 77])
tr = I3DTrainer(sd, 2, (16, 96, 96), avg_pool=(2, 3, 3), device=dev, optimizer="sgd")
_, _, g64, _ = train_oracle.loss_and_grads(sd, x, target, avg_pool=(2, 3, 3), dtype=torch.float64)
_, _, g32, _ = train_oracle.loss_and_grads(sd, x, target, avg_pool=(2, 3, 3))
tr.forward_backward(x, target)
for k in g64:
    print("%-40s mine %.2e torch32 %.2e  |g| %.3e" % (k, rel_err(tr.grads[k].cpu(), g64[k]), rel_err(g32[k], g64[k]), float(g64[k].norm())))
