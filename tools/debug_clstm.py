"""Diagnostic (GPU): where does the bf16 ConvLSTM mask gradient leave the fp32 one? Compares the bf16 and
fp32 engines' intermediate gradients on the golden clip (hid 32)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_clstm import build, engine
from common import GOLD, rel_err
from oracle import synthetic

dev = torch.device("cuda")
hid = 32
g = np.load(os.path.join(GOLD, "clstm_hid%d.npz" % hid))
_, sd = build(hid)
x1 = synthetic.clips(1, t=32, h=120, w=160) / 255.0
x = torch.cat([x1, synthetic.clips(2, t=32, h=120, w=160)[1:] / 255.0])
masks = torch.stack([torch.from_numpy(g["mask"]), torch.rand(32, generator=torch.Generator().manual_seed(8))])
res = {}
for mode in ("fp32", "bf16"):
    eng = engine(sd, hid, 2, mode, dev)
    eng.set_input(x.to(dev)); eng.set_targets(torch.tensor([2, 4]))
    lg = eng.forward(masks.to(dev), "reverse").clone().cpu()
    dm = eng.backward().clone().cpu()
    he = eng.he
    r = dict(logits=lg, dm=dm)
    for l, rec in enumerate(eng.layers):
        r["dH%d" % l] = rec["dH"].float().cpu()[..., :hid]
        r["dpre%d" % l] = rec["dpre"].buf.float().cpu().view(-1, 4, he)[..., :hid]
        r["h%d" % l] = rec["h"].buf.float().cpu()[..., :hid]
        gp = rec["g_pooled"].buf.float().cpu()
        if rec["s2d_out"]:  # [N,1,h/4,w/4,(dy,dx,c)] -> [N,1,h/2,w/2,c]
            n_, _, a, b, _ = gp.shape
            gp = gp.view(n_, 1, a, b, 2, 2, he).permute(0, 1, 2, 4, 3, 5, 6).reshape(n_, 1, 2 * a, 2 * b, he)
        r["gpool%d" % l] = gp[..., :hid]
        r["argmax%d" % l] = rec["argmax"].cpu()[..., :hid].float()
    if mode == "fp32":
        r["gx"] = eng.g_xin.buf.float().cpu()  # [T*B,1,H,W,3]
    else:
        gs = eng.g_xin.buf.float().cpu()[..., :12]  # [T*B,1,H/2,W/2,(dh,dw,c)]
        n = gs.shape[0]
        r["gx"] = gs.view(n, 1, 60, 80, 2, 2, 3).permute(0, 1, 2, 4, 3, 5, 6).reshape(n, 1, 120, 160, 3)
    res[mode] = r
for k in res["fp32"]:
    a, b = res["bf16"][k], res["fp32"][k]
    print("%-8s rel_err(bf16 vs fp32) = %.4f   |fp32| = %.3e" % (k, rel_err(a, b), float(b.norm())))
print("dm fp32 clip0", res["fp32"]["dm"][0].numpy())
print("dm bf16 clip0", res["bf16"]["dm"][0].numpy())
print("golden        ", g["dmask"])
for l in (1, 0):
    for k in ("gpool%d" % l, "dH%d" % l, "dpre%d" % l):
        a, b = res["bf16"][k], res["fp32"][k]
        a, b = a.reshape(32, 2, -1), b.reshape(32, 2, -1)
        print(k, "per-step:", [round(rel_err(a[t], b[t]), 3) if float(b[t].norm()) > 0 else None for t in range(32)])
    am_b, am_f = res["bf16"]["argmax%d" % l], res["fp32"]["argmax%d" % l]
    print("argmax mismatch fraction layer", l, float((am_b != am_f).float().mean()))
# per-step error of dH0 (layer 0) to see growth along BPTT
d0b, d0f = res["bf16"]["dH0"].view(32, 2, -1), res["fp32"]["dH0"].view(32, 2, -1)
print("dH0 per-step rel err:", [round(rel_err(d0b[t], d0f[t]), 3) for t in range(32)])
gb, gf = res["bf16"]["gx"].view(32, 2, -1), res["fp32"]["gx"].view(32, 2, -1)
print("gx per-step rel err:", [round(rel_err(gb[t], gf[t]), 3) for t in range(32)])

# structured clips: is the bf16 mask gradient usable where the clip has real temporal content?
xs = synthetic.clips(2, kind="moving_square", t=32, h=120, w=160) / 255.0
ms = torch.stack([torch.from_numpy(g["mask"]), torch.rand(32, generator=torch.Generator().manual_seed(8))])
out = {}
for mode in ("fp32", "bf16"):
    eng = engine(sd, hid, 2, mode, dev)
    eng.set_input(xs.to(dev)); eng.set_targets(torch.tensor([2, 4]))
    eng.forward(ms.to(dev), "reverse")
    out[mode] = eng.backward().clone().cpu()
for i in range(2):
    a, b = out["bf16"][i], out["fp32"][i]
    print("moving-square clip %d: dm rel_err %.4f cos %.4f |dm| %.3e" % (
        i, rel_err(a, b), float(torch.nn.functional.cosine_similarity(a, b, dim=0)), float(b.norm())))
for i in range(2):
    a, b = res["bf16"]["dm"][i], res["fp32"]["dm"][i]
    print("noise clip %d: dm rel_err %.4f cos %.4f" % (i, rel_err(a, b), float(torch.nn.functional.cosine_similarity(a, b, dim=0))))
