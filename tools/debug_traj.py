"""Debug: bf16 search trajectory of one clip against the oracle loop (fp32 and matched-rounding)."""
import sys, os
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from common import i3d_state_dict, quiet, rel_err
from oracle import i3d_oracle, mask_oracle, synthetic
from interpreting_video_features_b200.engine import I3DEngine
from interpreting_video_features_b200.search import MaskSearch
import warnings; warnings.filterwarnings("ignore")
dev = torch.device("cuda")
sd, _ = quiet(i3d_state_dict, 174)
AP = (2, 2, 2)
x = synthetic.clips(3, kind="square", t=16, h=64, w=64)
inits = torch.stack([torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4), torch.tensor([5.] * 13 + [-5.] * 3),
                     torch.tensor([2.5, -2.5, -2.5, 2.5, 2.5, -2.5, 2.5, -2.5, -2.5, -2.5, 2.5, 2.5, -2.5, 2.5, -2.5, -2.5])])
xp = torch.cat([mask_oracle.perturb_sequence(x[i:i + 1], torch.sigmoid(inits[i]), "freeze") for i in range(3)])
sd_head = i3d_oracle.sharpen_head_only(sd, torch.cat([x, xp]), AP)
with torch.no_grad():
    targets = i3d_oracle.forward(sd_head, xp, AP).argmax(dim=1)
i = 1
for mode in ("bf16", "fp32"):
    eng = I3DEngine(sd_head, 3, (16, 64, 64), mode=mode, softmax=True, avg_pool=AP, device=dev)
    rec = {}
    res = MaskSearch(eng, lam1=0.01, lam2=0.02, n_iter=50, perturb="freeze", use_graph=False).run(
        x.to(dev), targets, raw_masks=inits.to(dev), record=rec)
    print(mode, "ours class traj", ["%.3f" % float(c[i]) for c in rec["class"]][:50:3])
    print(mode, "ours final", (res["time_mask"][i].cpu() > 0.5).int().tolist())
    print(mode, "ours raw mask it0,1,2,5:", [rec["mask"][k][i].cpu().numpy().round(2).tolist() for k in (0, 1, 2, 5)])
    print(mode, "ours dm_class it0:", rec["dm_class"][0][i].cpu().numpy().round(4).tolist())
for quant in (False, True):
    model = i3d_oracle.Model(sd_head, AP, True, quant=quant)
    tm = inits[i].clone().requires_grad_(); r = {}
    final, cls = mask_oracle.mask_search(x[i:i + 1], model, 0, [int(targets[i])], tm, 0.01, 0.02, 50, record=r)
    print("oracle quant", quant, "class traj", ["%.3f" % c for c in r["class"]][:50:3])
    print("oracle raw mask it0,1,2,5:", [r["mask"][k].numpy().round(2).tolist() for k in (0, 1, 2, 5)])
    s0 = torch.sigmoid(inits[i]); 
    print("oracle grad it0 (raw, incl reg):", r["grad"][0].numpy().round(5).tolist())
    mi = s0.clone().requires_grad_()
    out = model(mask_oracle.perturb_sequence(x[i:i + 1], mi, "freeze"))[0, int(targets[i])]
    (g,) = torch.autograd.grad(out, mi)
    print("oracle dm_class it0:", g.numpy().round(4).tolist())
