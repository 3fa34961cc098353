"""Where do the ~20 us of a thin slab-kernel launch go?  One process, isolated back-to-back launches of small
3x3x3 layers with pairs / PDL / tile-plan knobs toggled through the environment (read per launch by the library).
GPU only: python tools/thin_layer_diag.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpreting_video_features_b200 import _lib, engine, ops  # noqa: E402
from interpreting_video_features_b200.ops import Act  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
N = 8


def layer(dhw, cin, cout):
    x = Act(torch.randn((N,) + dhw + (cin,), generator=g).to(dev).bfloat16(), N, *dhw, cin, 0, cin)
    out = Act(torch.zeros((N,) + dhw + (cout,), dtype=torch.bfloat16, device=dev), N, *dhw, cout, 0, cout)
    w = engine.pack_fwd(torch.randn((cout, cin, 3, 3, 3), generator=g).to(dev) * 0.05, "bf16")
    sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    return lambda: ops.conv3d(x, w, out, (3, 3, 3), (1, 1, 1), (1, 1, 1), flags=_lib.EP_RELU, scale=sc, shift=sh)


def timed(f, reps=50):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


LAYERS = [("3b.b2b 16->32 @8x28x28", (8, 28, 28), 16, 32), ("4c.b2b 24->64 @4x14x14", (4, 14, 14), 24, 64),
          ("5b.b2b 32->128 @2x7x7", (2, 7, 7), 32, 128), ("4b.b1b 96->208 @4x14x14", (4, 14, 14), 96, 208)]
CONFIGS = [{}, {"IVF_SLAB_2CTA": "0"}, {"IVF_SLAB_ACC": "1"}, {"IVF_SLAB_MT": "1"}, {"IVF_SLAB_MT": "2"},
           {"IVF_SLAB_MT": "4"}, {"IVF_SLAB": "0"}, {"IVF_SLAB_VERBOSE": "1"}]
fs = [(n, layer(d, ci, co)) for n, d, ci, co in LAYERS]
for cfg in CONFIGS:
    for k in ("IVF_SLAB_2CTA", "IVF_SLAB_ACC", "IVF_SLAB_MT", "IVF_SLAB", "IVF_SLAB_VERBOSE"):
        os.environ.pop(k, None)
    os.environ.update(cfg)
    if cfg.get("IVF_SLAB_VERBOSE"):
        for n, f in fs:
            f()
        torch.cuda.synchronize()
        continue
    print("%-22s" % (cfg or "default"), " | ".join("%s %.1f us" % (n.split()[0], timed(f)) for n, f in fs), flush=True)
