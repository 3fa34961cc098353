"""Timing experiment on the stem convolution (results garbage under IVF_SLAB_DIAG): which producer bounds which
tile plan.  GPU only."""
import os
import subprocess
import sys

CODE = r'''
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("tools/x"))))
from interpreting_video_features_b200 import _lib, engine, ops
from interpreting_video_features_b200.ops import Act
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
N, dhw, cin, cout, k = 8, (8, 112, 112), 24, 64, (4, 4, 4)
x = Act(torch.randn((N,) + dhw + (32,), generator=g).to(dev).bfloat16(), N, *dhw, 32, 0, cin)
out = Act(torch.zeros((N,) + dhw + (64,), dtype=torch.bfloat16, device=dev), N, *dhw, 64, 0, cout)
w = engine.pack_fwd(torch.randn((cout, cin) + k, generator=g).to(dev) * 0.05, "bf16")
sc = torch.ones(cout, device=dev); sh = torch.zeros(cout, device=dev)
plan = tuple(int(v) for v in os.environ["PLAN"].split(","))
f = lambda: ops.conv3d(x, w, out, k, (1, 1, 1), (1, 1, 1), flags=_lib.EP_RELU, scale=sc, shift=sh, plan=plan)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print("%.1f" % (e0.elapsed_time(e1) * 100))
'''
KCH64 = os.environ.get("KCH64", "0")
for plan in ("1,4,2,2,1", "1,3,2,2,1", "1,2,2,2,1", "1,2,2,1,1", "2,2,2,2,1", "2,2,2,1,1"):  # kwm, mt, acc, ncta, ntiles
    row = []
    for diag in (0, 3):
        env = dict(os.environ, PLAN=plan, IVF_SLAB_DIAG=str(diag), IVF_SLAB_KCH64=KCH64, IVF_SLAB_VERBOSE="0")
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
        row.append(r.stdout.strip() or r.stderr.strip()[-80:])
    print("KCH64=%s plan (kwm,mt,acc,ncta,nt) %-12s us: loads on %s | no loads %s" % ((KCH64, plan) + tuple(row)))
