"""Timing experiment on one slab-kernel layer with the TMA producers switched off (IVF_SLAB_DIAG, results are
garbage): what part of a layer's time is operand loading.  GPU only:  python tools/layer_diag.py"""
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
N = 8
dhw = tuple(int(v) for v in os.environ["DHW"].split(","))
cin, cout = int(os.environ["CIN"]), int(os.environ["COUT"])
k = (3, 3, 3)
x = Act(torch.randn((N,) + dhw + (cin,), generator=g).to(dev).bfloat16(), N, *dhw, cin, 0, cin)
out = Act(torch.zeros((N,) + dhw + (cout,), dtype=torch.bfloat16, device=dev), N, *dhw, cout, 0, cout)
w = engine.pack_fwd(torch.randn((cout, cin) + k, generator=g).to(dev) * 0.05, "bf16")
sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
f = lambda: ops.conv3d(x, w, out, k, (1, 1, 1), (1, 1, 1), flags=_lib.EP_RELU, scale=sc, shift=sh)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print("%.1f" % (e0.elapsed_time(e1) * 50))
'''
for name, dhw, cin, cout in (("3b.b1b", "8,28,28", 96, 128), ("3c.b1b", "8,28,28", 128, 192), ("4b.b1b", "4,14,14", 96, 208),
                             ("4f.b1b", "4,14,14", 160, 320), ("5c.b1b", "2,7,7", 192, 384),
                             ("3b.b2b", "8,28,28", 16, 32), ("4c.b2b", "4,14,14", 24, 64), ("5b.b2b", "2,7,7", 32, 128)):
    row = []
    for diag in (0, 3, 4, 7):
        env = dict(os.environ, DHW=dhw, CIN=str(cin), COUT=str(cout), IVF_SLAB_DIAG=str(diag))
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        row.append(r.stdout.strip() or r.stderr.strip()[-120:])
    print("%-8s %s %d->%d  us: all on %s | no operand loads %s | no epilogue traffic %s | neither %s" %
          (name, dhw, cin, cout, row[0], row[1], row[2], row[3]))
