"""In-situ per-launch timing of one eager mask-search iteration (warm L2: every kernel runs right after its producer,
as in the captured graph): CUDA events around every libivf launch, all queued behind a device-side sleep so that the
pairs bracket kernel time rather than host launch latency.  Complements the ncu launch list (cold cache).
GPU only:  python tools/step_events.py [clips] [reps]"""
import collections
import os
import sys

import torch

os.environ.setdefault("IVF_STREAMS", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CLIPS, NCLS, state_dict  # noqa: E402
from interpreting_video_features_b200 import ops, search  # noqa: E402
from oracle import synthetic  # noqa: E402

clips_n = int(sys.argv[1]) if len(sys.argv) > 1 else CLIPS
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
model = state_dict().to(dev).eval().set_mode("bf16")
clips = torch.stack([synthetic.uniform_clip(i) for i in range(clips_n)])
ms = search.MaskSearch(search.make_engines(model, clips, clips_n, 1), use_graph=False)
ms.set_input(clips.to(dev))
ms.set_targets((torch.arange(clips_n) % NCLS).to(dev))
ms.m.copy_(torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4, device=dev).repeat(clips_n, 1))
ops.sigmoid(ms.m, ms.sig)
for _ in range(2):
    ms._iteration()
torch.cuda.synchronize()

NAMES = ["conv3d", "conv3d_pair", "conv1x1_split", "maxpool3d_fwd", "maxpool3d_bwd", "perturb_fwd", "perturb_bwd", "head_fwd",
         "head_bwd", "mask_loss_adam"]
events = []
orig = {}


def wrap(name):
    f = getattr(ops, name)
    orig[name] = f

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f(*a, **k)
        e1.record()
        events.append((name, a, k, e0, e1))
        return r
    setattr(ops, name, timed)


for n in NAMES:
    if hasattr(ops, n):
        wrap(n)


def describe(name, a, k):
    try:
        if name in ("conv3d", "conv1x1_split"):
            x, out = a[0], a[2]
            kern = a[3] if name == "conv3d" else (1, 1, 1)
            fl = k.get("flags", 0)
            tag = "fwd" if fl else ("dgrad" + ("+acc" if k.get("acc_in") is not None else "") + ("+mask" if k.get("mask") is not None else ""))
            return "%s k%s %d->%d @%dx%dx%dx%d %s" % (name, "x".join(map(str, kern)), x.c, out.c, x.n, x.d, x.h, x.w, tag)
        if name == "conv3d_pair":
            return "conv3d_pair " + " + ".join("k%s %d->%d @%dx%dx%dx%d %s" % (
                "x".join(map(str, q["kernel"])), q["x"].c, q["out"].c, q["x"].n, q["x"].d, q["x"].h, q["x"].w,
                "fwd" if q.get("flags", 0) else "dgrad") for q in a[:2])
        if name == "maxpool3d_fwd":
            x = a[0]
            return "pool_fwd k%s s%s c%d @%dx%dx%dx%d" % ("x".join(map(str, a[3])), "x".join(map(str, a[4])), x.c, x.n, x.d, x.h, x.w)
        if name == "maxpool3d_bwd":
            dx = a[2]
            return "pool_bwd k%s s%s c%d @%dx%dx%dx%d" % ("x".join(map(str, a[3])), "x".join(map(str, a[4])), dx.c, dx.n, dx.d, dx.h, dx.w)
    except Exception:
        pass
    return name


acc = collections.OrderedDict()
for rep in range(reps):
    del events[:]
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.08 * 1.9e9))
    ms._iteration()
    torch.cuda.synchronize()
    for i, (name, a, k, e0, e1) in enumerate(events):
        acc.setdefault(i, [describe(name, a, k), name, []])[2].append(e0.elapsed_time(e1) * 1e3)
fam = collections.defaultdict(float)
tot = 0.0
print("in-situ per-launch times (us, median of %d eager iterations, %d clips)" % (reps, clips_n))
for i, (desc, name, ts) in acc.items():
    t = sorted(ts)[len(ts) // 2]
    fam[name] += t
    tot += t
    print("%3d %7.1f  %s" % (i, t, desc))
print("total %.1f us over %d launches" % (tot, len(acc)))
for n, t in sorted(fam.items(), key=lambda kv: -kv[1]):
    print("  %-16s %8.1f us  %4.1f%%" % (n, t, 100 * t / tot))
