"""One mixed-precision training step (8 clips of 16x224x224) under the profiler:
ncu --profile-from-start off ... python tools/profile_train.py [bf16|fp32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from interpreting_video_features_b200.train import I3DTrainer  # noqa: E402
from oracle import synthetic  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
n = 8
sd = {k: v.detach().clone() for k, v in bench.state_dict().state_dict().items()}
x = torch.stack([synthetic.uniform_clip(5000 + i) for i in range(n)]).to(dev)
target = torch.arange(n) % 174
tr = I3DTrainer(sd, n, (16, 224, 224), device=dev, optimizer="sgd", lr=1e-3, momentum=0.9, dropout_p=0.5, mode=mode)
for _ in range(2):
    tr.step(x, target)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(x, target)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one %s training step for %d clips" % (mode, n))
