"""Debug: bf16 class gradient at a SATURATED mask (sigmoid(+-5)) against the imposed-decision oracle, link by link."""
import sys, os
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from common import i3d_state_dict, quiet, rel_err
from test_gpu_parity import engine_decisions, oracle_grad, cosine, SMALL
from oracle import i3d_oracle, mask_oracle, synthetic
from interpreting_video_features_b200.engine import I3DEngine
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
sig = torch.sigmoid(inits)
for softmax in (False, True):
    eng = I3DEngine(sd_head, 3, (16, 64, 64), mode="bf16", softmax=softmax, avg_pool=AP, device=dev)
    eng.set_input(x.to(dev)); eng.set_targets(targets)
    out = eng.forward(sig.to(dev), "freeze").clone().cpu()
    dm = eng.backward().clone().cpu()
    for i in range(3):
        force = engine_decisions(eng, clip=i)
        # oracle with imposed decisions, keeping the input gradient
        mi = sig[i].clone().requires_grad_()
        P = mask_oracle.perturb_sequence(x[i:i+1], mi, "freeze"); P.retain_grad()
        o = i3d_oracle.forward(sd_head, P, AP, softmax, quant=True, force=force)
        o[0, int(targets[i])].backward()
        g_f, gP = mi.grad.clone(), P.grad.clone()
        gs = eng.g_xin.buf[..., :24].float().cpu().view(3, 8, 32, 32, 2, 2, 2, 3)
        got_field = gs.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(3, 3, 16, 64, 64)[i:i+1]
        # last link only: oracle perturbation backward from OUR input gradient
        mi2 = sig[i].clone().requires_grad_()
        P2 = mask_oracle.perturb_sequence(x[i:i+1], mi2, "freeze")
        (g_link,) = torch.autograd.grad(P2, mi2, got_field)
        print("softmax", softmax, "clip", i, "out ours %.4f forced %.4f" % (float(out[i, targets[i]]), float(o[0, int(targets[i])])),
              "| field rel %.3e | dm vs forced rel %.3e cos %.5f | dm vs (oracle perturb' of OUR field) rel %.3e" % (
                  rel_err(got_field, gP), rel_err(dm[i], g_f), cosine(dm[i], g_f), rel_err(dm[i], g_link)))
        if i == 1:
            print("   ours  ", dm[i].numpy().round(3).tolist())
            print("   forced", g_f.numpy().round(3).tolist())
            print("   link  ", g_link.numpy().round(3).tolist())
