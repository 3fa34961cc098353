"""Timing experiment on the ConvLSTM's recurrent convolutions (layer 1 of config C3: 5x5, 32 -> 128 gate channels on
8 x 60 x 80, fp32 pre-activations accumulated in place; its data gradient 128 -> 32) with parts of the slab kernel
switched off (IVF_SLAB_DIAG: 1|2 = no operand loads, 4 = no epilogue loads / stores; results are garbage).
GPU only:  python tools/clstm_conv_diag.py"""
import os
import subprocess
import sys

CODE = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from interpreting_video_features_b200 import _lib, engine, ops
from interpreting_video_features_b200.ops import Act
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
N, H, W = 8, 60, 80
cin, cout = int(os.environ["CIN"]), int(os.environ["COUT"])
x = Act(torch.randn((N, 1, H, W, cin), generator=g).to(dev).bfloat16(), N, 1, H, W, cin, 0, cin)
acc = Act(torch.zeros((N, 1, H, W, cout), dtype=torch.float32, device=dev), N, 1, H, W, cout, 0, cout)
w = engine.pack_fwd(torch.randn((cout, cin, 1, 5, 5), generator=g).to(dev) * 0.05, "bf16")
f = lambda: ops.conv3d(x, w, acc, (1, 5, 5), (1, 1, 1), (0, 2, 2), acc_in=acc)
for _ in range(3): f()
torch.cuda.synchronize()
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
best = 1e9
for _ in range(5):
    junk.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): f()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 250)
print("%.1f" % best)
'''
for name, cin, cout in (("h-conv fwd 32->128", 32, 128), ("h-conv dgrad 128->32", 128, 32)):
    row = []
    for diag in (0, 3, 4, 7):
        env = dict(os.environ, CIN=str(cin), COUT=str(cout), IVF_SLAB_DIAG=str(diag), IVF_SLAB_VERBOSE="0")
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        row.append(r.stdout.strip() or r.stderr.strip()[-200:])
    print("%-22s us per launch: all on %s | no operand loads %s | no epilogue traffic %s | neither %s" % (name, *row))
