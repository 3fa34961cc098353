"""Per-tensor distance to the fp64 oracle gradient: the native training step vs torch's fp32 CPU evaluation.
usage: python tools/debug_train.py [fp32|bf16]"""
import os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from common import i3d_state_dict, quiet, rel_err
from interpreting_video_features_b200.train import I3DTrainer
from oracle import synthetic, train_oracle
mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
sd, _ = quiet(i3d_state_dict, 174)
x = synthetic.clips(2, t=16, h=96, w=96)
target = torch.tensor([5, 77])
tr = I3DTrainer(sd, 2, (16, 96, 96), avg_pool=(2, 3, 3), device=dev, optimizer="sgd", mode=mode)
l64, lg64, g64, _ = train_oracle.loss_and_grads(sd, x, target, avg_pool=(2, 3, 3), dtype=torch.float64, quant=(mode == "bf16"))
_, _, g32, _ = train_oracle.loss_and_grads(sd, x, target, avg_pool=(2, 3, 3))
loss = tr.forward_backward(x, target)
print("mode %s loss %.6f fp64 %.6f logits rel %.2e" % (mode, float(loss), l64, rel_err(tr.logits.cpu(), lg64)))
mine, cosv = [], []
for k in g64:
    a, b = tr.grads[k].cpu().double().flatten(), g64[k].flatten()
    cos = float((a @ b) / (a.norm() * b.norm() + 1e-300))
    mine.append(rel_err(a, b)); cosv.append(cos)
    print("%-40s mine %.2e cos %.5f torch32 %.2e  |g| %.3e" % (k, mine[-1], cos, rel_err(g32[k], g64[k]), float(g64[k].norm())))
print("median rel %.3e worst %.3e | median cos %.5f worst cos %.5f" % (statistics.median(mine), max(mine), statistics.median(cosv), min(cosv)))
