"""Sweep the slab kernel's tile plans (kw-merge, accumulators per tile, TMEM buffering, CTA pairs, N tiles) for
the convolution shapes of config C2 and print measured times next to the cost model's choice.
GPU only: python tools/tune_slab.py [reps]"""
import ctypes as C
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpreting_video_features_b200 import _lib, engine, ops  # noqa: E402
from interpreting_video_features_b200.ops import Act  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
N = 8
ENV = ("IVF_SLAB_2CTA", "IVF_SLAB_KWM", "IVF_SLAB_MT", "IVF_SLAB_ACC", "IVF_SLAB_NT")


def desc_of(x, out, k, pf, flags=0):
    d = _lib.ConvDesc()
    d.n, d.id, d.ih, d.iw = x.n, x.d, x.h, x.w
    d.od, d.oh, d.ow = out.d, out.h, out.w
    d.cin, d.cout = x.c, out.c
    d.kd, d.kh, d.kw = k
    d.sd = d.sh = d.sw = 1
    d.pd, d.ph, d.pw = pf
    d.in_ld, d.in_coff, d.out_ld, d.out_coff = x.ld, x.coff, out.ld, out.coff
    d.dtype, d.flags = _lib.IVF_BF16, flags
    return d


def plan(d):
    out = (C.c_int * 12)()
    ok = lib.ivf_conv_slab_plan(C.byref(d), 148, out)
    return tuple(out) if ok else None


def bench(name, dhw, cin, cout, k, pf, dgrad):
    g = torch.Generator().manual_seed(0)
    cin_buf = 32 if cin == 24 else cin
    cout_buf = 32 if cout == 24 else cout
    x = Act(torch.randn((N,) + dhw + (cin_buf,), generator=g).to(dev).bfloat16(), N, *dhw, cin_buf, 0, cin)
    out = Act(torch.zeros((N,) + dhw + (cout_buf,), dtype=torch.bfloat16, device=dev), N, *dhw, cout_buf, 0, cout)
    w = torch.randn((cout, cin) + k, generator=g).to(dev) * 0.05
    wp = engine.pack_fwd(w, "bf16")
    scale = torch.ones(cout, device=dev)
    shift = torch.zeros(cout, device=dev)
    mask = Act(torch.randn((N,) + dhw + (cout_buf,), generator=g).to(dev).bfloat16(), N, *dhw, cout_buf, 0, cout)
    kw = dict(mask=mask, mask_scale=scale) if dgrad else dict(flags=_lib.EP_RELU, scale=scale, shift=shift)
    flags = (_lib.EP_MASK if dgrad else (_lib.EP_RELU | _lib.EP_AFFINE))
    d = desc_of(x, out, k, pf, flags)

    def run_once():
        ops.conv3d(x, wp, out, k, (1, 1, 1), pf, **kw)

    def timeit():
        run_once()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run_once()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    for v in ENV:
        os.environ.pop(v, None)
    base_plan = plan(d)
    base = timeit()
    results = {}
    for pair, kwm, mt, acc, nt in itertools.product((0, 1), (1, 2, 3, 4), (1, 2, 3, 4), (1, 2), (0, 1, 2, 3)):
        if k[2] % kwm:
            continue
        os.environ.update({"IVF_SLAB_2CTA": str(pair), "IVF_SLAB_KWM": str(kwm), "IVF_SLAB_MT": str(mt),
                           "IVF_SLAB_ACC": str(acc), "IVF_SLAB_NT": str(nt)})
        p = plan(d)
        if p is None:
            continue
        key = (p[11], p[10], p[3], p[5], p[2], p[1])  # ncta kwm mt acc ntiles bn
        if key in results or p[10] != kwm or p[3] != mt or p[5] != acc or p[11] != pair + 1:
            continue
        results[key] = timeit()
    for v in ENV:
        os.environ.pop(v, None)
    best = sorted(results.items(), key=lambda kv: kv[1])[:4]
    bp = (base_plan[11], base_plan[10], base_plan[3], base_plan[5], base_plan[2], base_plan[1]) if base_plan else None
    print("%-18s model %s %.1f us | best " % (name, bp, base) +
          " ; ".join("%s %.1f" % (kk, vv) for kk, vv in best), flush=True)


SHAPES = [
    ("stem", (8, 112, 112), 24, 64, (4, 4, 4), (1, 1, 1)),
    ("2c", (8, 56, 56), 64, 192, (3, 3, 3), (1, 1, 1)),
    ("3b.b1b", (8, 28, 28), 96, 128, (3, 3, 3), (1, 1, 1)),
    ("3b.b2b", (8, 28, 28), 16, 32, (3, 3, 3), (1, 1, 1)),
    ("3c.b1b", (8, 28, 28), 128, 192, (3, 3, 3), (1, 1, 1)),
    ("3c.b2b", (8, 28, 28), 32, 96, (3, 3, 3), (1, 1, 1)),
    ("4b.b1b", (4, 14, 14), 96, 208, (3, 3, 3), (1, 1, 1)),
    ("4d.b1b", (4, 14, 14), 128, 256, (3, 3, 3), (1, 1, 1)),
    ("4f.b1b", (4, 14, 14), 160, 320, (3, 3, 3), (1, 1, 1)),
    ("4f.b2b", (4, 14, 14), 32, 128, (3, 3, 3), (1, 1, 1)),
    ("5c.b1b", (2, 7, 7), 192, 384, (3, 3, 3), (1, 1, 1)),
]
for name, dhw, cin, cout, k, pf in SHAPES:
    bench(name + " fwd", dhw, cin, cout, k, pf, False)
    pfd = tuple(kk - 1 - p for kk, p in zip(k, pf))
    bench(name + " dgrad", dhw, cout, cin, k, pfd, True)
