"""Times the four stage pools of I3D (forward and pure-routing backward) at the C2 geometry, L2 flushed between
launches, under the environment the caller sets (IVF_POOL_ROWS, IVF_POOL_ROWS_WAVES, IVF_POOL_S2ROUTE, ...).
GPU only:  python tools/pool_bench.py [clips]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpreting_video_features_b200 import ops  # noqa: E402
from interpreting_video_features_b200.ops import Act, same_pad  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
CASES = [("2a", (1, 3, 3), (1, 2, 2), (8, 112, 112), 64), ("3a", (1, 3, 3), (1, 2, 2), (8, 56, 56), 192),
         ("4a", (3, 3, 3), (2, 2, 2), (8, 28, 28), 480), ("5a", (2, 2, 2), (2, 2, 2), (4, 14, 14), 832)]
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(f, reps=5):
    best = 1e9
    for _ in range(reps):
        junk.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3)
    return best


tot_f = tot_b = 0.0
for name, k, s, dhw, c in CASES:
    geo = [same_pad(sz, kk, ss) for sz, kk, ss in zip(dhw, k, s)]
    pf, od = tuple(q[0] for q in geo), tuple(q[2] for q in geo)
    x = Act(torch.relu(torch.randn((n,) + dhw + (c,), device=dev)).bfloat16(), n, *dhw, c)
    out = Act.empty(n, *od, c, torch.bfloat16, dev)
    am = torch.empty((out.pixels, c), dtype=torch.uint8, device=dev)
    dy = Act(torch.randn((n,) + od + (c,), device=dev).bfloat16(), n, *od, c)
    dx = x.like()
    f = lambda: ops.maxpool3d_fwd(x, out, am, k, s, pf, nonneg=True)
    b = lambda: ops.maxpool3d_bwd(dy, am, dx, k, s, pf)
    f(); b()
    torch.cuda.synchronize()
    tf, tb = timed(f), timed(b)
    mb_f = (x.buf.numel() * 2 + out.buf.numel() * 3) / 1e6
    mb_b = (out.buf.numel() * 3 + x.buf.numel() * 2) / 1e6
    print("%s fwd %6.1f us (%5.0f MB, %4.2f TB/s) | bwd %6.1f us (%5.0f MB, %4.2f TB/s)" %
          (name, tf, mb_f, mb_f / tf, tb, mb_b, mb_b / tb))
    tot_f += tf
    tot_b += tb
print("total fwd %.1f us, bwd %.1f us" % (tot_f, tot_b))
