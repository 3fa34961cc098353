"""Where the end-to-end time of one 8-clip, 300-iteration search goes (host wall clock with synchronisation
between phases).  GPU only: python tools/e2e_breakdown.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CLIPS, NCLS, N_ITER, state_dict  # noqa: E402
from interpreting_video_features_b200 import ops, search  # noqa: E402
from oracle import synthetic  # noqa: E402

dev = torch.device("cuda:0")
model = state_dict().to(dev).eval().set_mode("bf16")
clips = torch.stack([synthetic.uniform_clip(i) for i in range(CLIPS)]).pin_memory()
targets = torch.arange(CLIPS) % NCLS


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print("  %-34s %8.1f ms" % (label, (t1 - t0) * 1e3))
    return t1


for rep in range(2):
    print("search %d" % rep)
    torch.cuda.synchronize()
    t = t00 = time.perf_counter()
    x = clips.to(dev, non_blocking=True).contiguous()
    t = tick("H2D", t)
    engs = search.make_engines(model, x, CLIPS, 1)
    t = tick("engines (cached after the first)", t)
    ms = search.MaskSearch(engs, 0.01, 0.02, 0.2, N_ITER, "freeze", 0.9, True)
    t = tick("MaskSearch()", t)
    ms.set_input(x)
    ms.set_targets(targets.to(dev))
    t = tick("set_input/targets", t)
    raw, probs = ms.init_masks(targets.to(dev).long(), "central")
    t = tick("init_masks", t)
    ms.m.copy_(raw)
    ops.sigmoid(ms.m, ms.sig)
    ms._capture()
    t = tick("graph capture + instantiate", t)
    for _ in range(N_ITER):
        ms.graph.replay()
    t = tick("300 replays", t)
    ms.forward(ms.sig.clone(), "reverse")
    t = tick("reverse score", t)
    print("  total %.1f ms" % ((t - t00) * 1e3))
    t0 = time.perf_counter()
    res = search.find_masks_batched(model, clips, targets, n_iter=N_ITER, micro_batch=CLIPS, device=dev)
    res["time_mask"].cpu()
    torch.cuda.synchronize()
    print("  find_masks_batched: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
