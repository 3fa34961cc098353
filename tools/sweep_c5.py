#!/usr/bin/env python
"""Config C5 (BASELINE.json configs[4]): the combined I3D-vs-ConvLSTM comparison sweep - per clip, Grad-CAM and the
temporal-mask search on both architectures, each in bf16 and fp32 mode, with the maximum relative error of every
quantity against the oracle (the CPU restatement of the reference).  Test infrastructure: imports oracle/.

    python tools/sweep_c5.py [--clips 2] [--iters 10] [--small] > profiles/r02_c5_sweep.json
North-star tolerances: logits / probabilities / CAMs 1e-2 relative in bf16, 1e-4 in fp32; final-mask IoU >= 0.95."""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def rel(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def iou(a, b):
    a, b = torch.as_tensor(a) > 0.5, torch.as_tensor(b) > 0.5
    u = float((a | b).sum())
    return 1.0 if u == 0 else float((a & b).sum()) / u


def cam_err(got, want):
    ok = ~np.isnan(want)
    nan_ok = bool(np.array_equal(np.isnan(got), np.isnan(want)))
    return (float(np.abs(got[ok] - want[ok]).max()) if ok.any() else 0.0), nan_ok


def sweep(n_clips=2, n_iter=10, small=False, dev=None):
    from interpreting_video_features_b200 import search
    from interpreting_video_features_b200.pt.grad_cam_videos import GradCamVideo
    from interpreting_video_features_b200.pt.models import CLSTM_4, I3D_doubled, I3D_doubled_kth
    from oracle import clstm_oracle, gradcam_oracle, i3d_oracle, mask_oracle, synthetic
    dev = dev or torch.device("cuda")
    geo = dict(smth=(16, 64, 64, (2, 2, 2)) if small else (16, 224, 224, (2, 7, 7)),
               kth=(32, 120, 160, (4, 4, 5)))
    report = {"clips": n_clips, "iterations": n_iter, "small_smth_geometry": bool(small), "cases": []}
    # ---- I3D: C2 (smth, freeze search + Grad-CAM) and C1 (KTH Grad-CAM)
    for name, ctor, ncls, kw in (("i3d_smth", I3D_doubled, 174, {}), ("i3d_kth", I3D_doubled_kth, 6, dict(finalTimeLength=4))):
        t, h, w, ap = geo["smth" if name == "i3d_smth" else "kth"]
        torch.manual_seed(0)
        ref_model = quiet(ctor.Model, ncls, last_stride=1, stride_mod_layers="", softMax=1, **kw)
        sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
        x = synthetic.clips(n_clips, kind="square", t=t, h=h, w=w).floor()
        omodel = i3d_oracle.Model(sd, ap, True)
        with torch.no_grad():
            want_p = omodel(x)
        targets = want_p.argmax(dim=1)
        oracle_masks, oracle_cams = [], []
        for i in range(n_clips):
            if name == "i3d_smth":
                raw = mask_oracle.init_mask(x[i:i + 1], omodel, 0, [int(targets[i])])
                final, _ = mask_oracle.mask_search(x[i:i + 1], omodel, 0, [int(targets[i])], raw, 0.01, 0.02, n_iter)
                oracle_masks.append(final)
            oracle_cams.append(gradcam_oracle.gradcam_i3d(sd, x[i:i + 1], int(targets[i]), (w, h), True, avg_pool=ap)[0])
        for mode in ("fp32", "bf16"):
            model = quiet(ctor.Model, ncls, last_stride=1, stride_mod_layers="", softMax=1, **kw)
            model.load_state_dict(sd)
            model.avg_pool.kernel_size = list(ap)
            model = model.to(dev).eval().set_mode(mode)
            with torch.no_grad():
                got_p = model(x.to(dev)).cpu()
            case = {"model": name, "mode": mode, "clip": [3, t, h, w], "probs_rel_err": rel(got_p, want_p)}
            gc = GradCamVideo(model=model, target_layer_names=["Mixed_5c"], class_dict=None, use_cuda=True,
                              input_spatial_size=(w, h), normalizePerFrame=True, archType="I3D")
            cams, _ = gc.batched(x.to(dev), targets.tolist())
            errs = [cam_err(cams[i], oracle_cams[i]) for i in range(n_clips)]
            case["cam_max_abs_err"], case["cam_nan_pattern_equal"] = max(e[0] for e in errs), all(e[1] for e in errs)
            if name == "i3d_smth":
                res = search.find_masks_batched(model, x, targets, n_iter=n_iter, micro_batch=n_clips, device=dev)
                tm = res["time_mask"].cpu()
                case["final_mask_iou_min"] = min(iou(tm[i], oracle_masks[i]) for i in range(n_clips))
                case["final_mask_max_abs_diff"] = max(float((tm[i] - oracle_masks[i]).abs().max()) for i in range(n_clips))
            report["cases"].append(case)
            del model, gc
            torch.cuda.empty_cache()
    # ---- ConvLSTM: C3 (KTH, reverse search + Grad-CAM), hidden 32 as in the paper
    t, h, w, _ = geo["kth"]
    kwc = dict(num_layers=2, kernel=5, conv_stride=2, effective_step=(7, 15, 23, 31))
    torch.manual_seed(0)
    cm = quiet(CLSTM_4.Model, num_classes=6, nb_lstm_units=32, channels=3, conv_kernel_size=(5, 5), lstm_layers=2, step=32,
               conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31], batch_normalization=True,
               dropout=0.5, add_softmax=True).eval()
    sdc = {k: v.detach().clone() for k, v in cm.state_dict().items()}
    x = synthetic.clips(n_clips, kind="square", t=t, h=h, w=w).floor()
    om = clstm_oracle.Model(sdc, hidden=32, softmax=True, **kwc)
    with torch.no_grad():
        want_p = om(x)
    targets = want_p.argmax(dim=1)
    raw0 = torch.tensor([-5.] * 8 + [5.] * 16 + [-5.] * 8)
    oracle_masks, oracle_cams = [], []
    for i in range(n_clips):
        final, _ = mask_oracle.mask_search(x[i:i + 1], om, 0, [int(targets[i])], raw0.clone().requires_grad_(), 0.02, 0.04,
                                           n_iter, mask_type="reverse")
        oracle_masks.append(final)
        oracle_cams.append(gradcam_oracle.gradcam_clstm(sdc, x[i:i + 1], int(targets[i]), (w, h), True, hidden=32, **kwc)[0])
    for mode in ("fp32", "bf16"):
        model = quiet(CLSTM_4.Model, num_classes=6, nb_lstm_units=32, channels=3, conv_kernel_size=(5, 5), lstm_layers=2,
                      step=32, conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31],
                      batch_normalization=True, dropout=0.5, add_softmax=True)
        model.load_state_dict(sdc)
        model = model.to(dev).eval().set_mode(mode)
        with torch.no_grad():
            got_p = model(x.to(dev)).cpu()
        case = {"model": "clstm_kth_hid32", "mode": mode, "clip": [3, t, h, w], "probs_rel_err": rel(got_p, want_p)}
        gc = GradCamVideo(model=model, target_layer_names=["clstm"], class_dict=None, use_cuda=True,
                          input_spatial_size=(w, h), normalizePerFrame=True, archType="CLSTM")
        cams, _ = gc.batched(x.to(dev), targets.tolist())
        errs = [cam_err(cams[i], oracle_cams[i]) for i in range(n_clips)]
        case["cam_max_abs_err"], case["cam_nan_pattern_equal"] = max(e[0] for e in errs), all(e[1] for e in errs)
        eng = model._engine(x, batch=n_clips)
        ms = search.MaskSearch(eng, 0.02, 0.04, 0.2, n_iter, "reverse", 0.9)
        res = ms.run(x.to(dev), targets, raw_masks=raw0.repeat(n_clips, 1).to(dev))
        tm = res["time_mask"].cpu()
        case["final_mask_iou_min"] = min(iou(tm[i], oracle_masks[i]) for i in range(n_clips))
        case["final_mask_max_abs_diff"] = max(float((tm[i] - oracle_masks[i]).abs().max()) for i in range(n_clips))
        report["cases"].append(case)
    return report


def check(report):
    """The north star's tolerances; returns the list of violations."""
    bad = []
    for c in report["cases"]:
        tol = 1e-4 if c["mode"] == "fp32" else 1e-2
        if c["probs_rel_err"] > tol:
            bad.append((c["model"], c["mode"], "probs", c["probs_rel_err"]))
        if not c["cam_nan_pattern_equal"]:
            bad.append((c["model"], c["mode"], "cam NaN pattern"))
        if c["cam_max_abs_err"] > (2e-3 if c["mode"] == "fp32" else 1.5e-1):
            bad.append((c["model"], c["mode"], "cam", c["cam_max_abs_err"]))
        if "final_mask_iou_min" in c and c["final_mask_iou_min"] < 0.95:
            bad.append((c["model"], c["mode"], "IoU", c["final_mask_iou_min"]))
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--small", action="store_true", help="16x64x64 instead of 16x224x224 for the smth model")
    a = ap.parse_args()
    rep = sweep(a.clips, a.iters, a.small)
    rep["violations"] = check(rep)
    print(json.dumps(rep, indent=1))
    sys.exit(1 if rep["violations"] else 0)
