"""In-situ CUDA-event timing of one training step (I3DTrainer, 8 clips of 16x224x224) by launch family."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from interpreting_video_features_b200 import ops
from interpreting_video_features_b200.train import I3DTrainer
from oracle import synthetic
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
sd = {k: v.detach().clone() for k, v in bench.state_dict().state_dict().items()}
x = torch.stack([synthetic.uniform_clip(5000 + i) for i in range(n)]).to(dev)
tr = I3DTrainer(sd, n, (16, 224, 224), device=dev, optimizer="sgd", lr=1e-3, dropout_p=0.5, mode=mode)
tr.step(x, torch.arange(n) % 174)
torch.cuda.synchronize()
events = []
names = ["conv3d", "bn_train_fwd", "bn_train_bwd", "conv3d_wgrad", "conv3d_wgrad_s2d", "optim_step_multi", "maxpool3d_fwd", "maxpool3d_bwd", "head_train_fwd",
         "head_train_bwd", "optim_step", "perturb_fwd", "dropout_mask"]
orig = {}
for nm in names:
    orig[nm] = getattr(ops, nm)
    def wrap(*a, _f=orig[nm], _n=nm, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = _f(*a, **k); e1.record()
        tag = _n
        if _n == "conv3d":
            tag = "conv3d dgrad" if (k.get("transposed") or a[2].buf.dtype == torch.float32 and mode == "bf16") else "conv3d fwd"
        events.append((tag, e0, e1))
        return r
    setattr(ops, nm, wrap)
import interpreting_video_features_b200.engine as eng
orig_pack = eng.pack
tr.step(x, torch.arange(n) % 174)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for tag, e0, e1 in events:
    a = agg.setdefault(tag, [0, 0.0]); a[0] += 1; a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print("one training step (%s), %d clips: %.1f ms in %d timed launches (weight re-packing not timed)" % (mode, n, tot, len(events)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("  %-16s %4d launches %8.2f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
