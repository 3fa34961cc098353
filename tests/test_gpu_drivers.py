"""GPU (-m gpu): the repaired drivers on the native path (two-clip runs against the oracle loop and the result
schema), the visualisation kernel bit for bit against the host pipeline, and uint8 clip ingest."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from common import GOLD, REPO, i3d_state_dict, quiet, rel_err

pytestmark = pytest.mark.gpu

PT = os.path.join(REPO, "interpreting_video_features_b200", "pt")
if PT not in sys.path:
    sys.path.insert(0, PT)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from interpreting_video_features_b200 import _lib
    _lib.handle()
    return torch.device("cuda")


@pytest.mark.parametrize("kind", ["freeze", "reverse"])
@pytest.mark.parametrize("as_u8", [False, True])
def test_viz_triptych_bit_exact(dev, kind, as_u8):
    """ivf_viz_triptych (+ dots) == pt/visualisation.py:96-122,35-93 restated on the host (oracle/viz_oracle.py,
    pinned against cv2/numpy), including an all-NaN Grad-CAM slice."""
    import visualisation
    from oracle import mask_oracle, viz_oracle
    rs = np.random.RandomState(3)
    t, h, w = 8, 30, 44
    clip = torch.from_numpy(np.floor(rs.rand(3, t, h, w) * 256).astype(np.float32))
    cam = rs.rand(t, h, w).astype(np.float32)
    cam[2] = np.nan
    cam[5, 0, 0], cam[5, 1, 1] = 0.0, 1.0
    tm = torch.tensor([0.2, 0.7, 0.9, 0.4, 0.6, 0.55, 0.1, 0.8])
    pert = mask_oracle.perturb_sequence(clip[None], tm.clone(), kind, snap_values=True)[0].numpy()
    want = viz_oracle.triptych(clip.numpy(), cam, pert)
    x = clip.to(torch.uint8) if as_u8 else clip
    got = visualisation.triptych(x.to(dev), cam, tm, kind, draw_dots=False)
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), int(np.abs(got.astype(int) - want.astype(int)).max())
    dotted = visualisation.triptych(x, cam, tm, kind, draw_dots=True)
    assert np.array_equal(dotted, viz_oracle.draw_dots(want, tm.numpy(), w, h))
    assert float(tm[0]) == pytest.approx(0.2)  # the caller's mask is not snapped by the visualisation


def test_uint8_ingest_equals_float(dev):
    from interpreting_video_features_b200.engine import I3DEngine
    sd, _ = quiet(i3d_state_dict, 174)
    x8 = (torch.rand((2, 3, 16, 64, 64), generator=torch.Generator().manual_seed(0)) * 256).to(torch.uint8)
    eng = I3DEngine(sd, 2, (16, 64, 64), mode="bf16", softmax=True, avg_pool=(2, 2, 2), device=dev)
    eng.set_input(x8.float())
    a = eng.forward(None).clone()
    eng.set_input(x8.pin_memory())
    b = eng.forward(None).clone()
    assert torch.equal(eng.x.cpu(), x8.float()) and torch.equal(a, b)


def test_smth_driver_native_two_clips(dev, tmp_path):
    """find_masks of the repaired smth driver on the native fp32 path, two clips of interest out of three, six
    iterations, against the oracle's reference loop: masks, freeze / reverse scores, Grad-CAM heat maps, the pickles."""
    import FindMasksComparison_I3D_smth as drv
    from interpreting_video_features_b200.pt.models import I3D_doubled
    from oracle import gradcam_oracle, i3d_oracle, mask_oracle, synthetic
    sd, _ = quiet(i3d_state_dict, 174)
    x = synthetic.clips(3, kind="square", t=16, h=64, w=64).floor()  # decoded-frame values
    sds = i3d_oracle.sharpen_head_only(sd, x, (2, 2, 2))
    model = quiet(I3D_doubled.Model, 174, last_stride=1, stride_mod_layers="", softMax=1)
    model.load_state_dict(sds)
    model.avg_pool.kernel_size = [2, 2, 2]
    model = torch.nn.DataParallel(model, [0]).to(dev).eval()
    model.module.set_mode("fp32")
    labels = torch.tensor([5, 9, 11])
    csv = tmp_path / "coi.csv"
    csv.write_text("5,11\n100,102\n")
    drv.RESIZE_SIZE_WIDTH, drv.RESIZE_SIZE_HEIGHT = 64, 64
    batches = [(x.to(torch.uint8), labels, ["100", "101", "102"])]
    masks = drv.find_masks(batches, model, {"gradCamType": "guessed", "batch_size": 3}, 0.01, 0.02, 6, "central", "freeze",
                           classOI=str(csv), doGradCam=True, runTempMask=True, sub_dir="t", out_root=str(tmp_path),
                           verbose=False)
    tm = pickle.load(open(tmp_path / "results" / "allTimeMaskResults_t_coi.csv_.p", "rb"))
    gc = pickle.load(open(tmp_path / "results" / "allGradCamResults_t_coi.csv_.p", "rb"))
    assert len(masks) == 2 and [r["video_id"] for r in tm] == ["100", "102"] and [r["video_id"] for r in gc] == [100, 102]
    omodel = i3d_oracle.Model(sds, (2, 2, 2), True)
    with torch.no_grad():
        out = omodel(x)
    for r, g, b in zip(tm, gc, (0, 2)):
        pred = int(out[b].argmax())
        assert r["pred_class"] == pred and r["true_class"] == int(labels[b])
        assert abs(r["original_score_guess"] - float(out[b].max())) < 1e-4
        assert abs(r["original_score_true"] - float(out[b, labels[b]])) < 1e-4 * max(float(out[b, labels[b]]), 1e-3) + 1e-7
        traw = mask_oracle.init_mask(x[b:b + 1], omodel, 0, [pred])
        final, cls = mask_oracle.mask_search(x[b:b + 1], omodel, 0, [pred], traw, 0.01, 0.02, 6)
        assert np.abs(r["time_mask"] - final.numpy()).max() < 2e-3
        assert abs(r["freeze_score"] - cls) < 2e-3 * abs(cls) + 1e-6
        with torch.no_grad():
            rev = omodel(mask_oracle.perturb_sequence(x[b:b + 1], final, "reverse"))[0, pred]
        assert abs(r["reverse_score"] - float(rev)) < 5e-3 * abs(float(rev)) + 1e-6
        want, _, _ = gradcam_oracle.gradcam_i3d(sds, x[b:b + 1], pred, (64, 64), True, avg_pool=(2, 2, 2))
        ok = ~np.isnan(want)
        assert g["GCHeatMap"].shape == (16, 64, 64) and np.array_equal(np.isnan(g["GCHeatMap"]), np.isnan(want))
        if ok.any():  # an all-zero class-activation map normalises to NaN everywhere, here and in the reference
            assert np.abs(g["GCHeatMap"][ok] - want[ok]).max() < 2e-3
        folder = tmp_path / "cam_saved_images" / "t" / str(r["true_class"]) / (
            r["video_id"] + "g_%d_gs%5.4f_cs%5.4f" % (pred, r["original_score_guess"], r["original_score_true"])) / "combined"
        assert float((folder / ("ClassScoreFreezecase%s.txt" % r["video_id"])).read_text()) == r["freeze_score"]
        assert float((folder / ("ClassScoreReversecase%s.txt" % r["video_id"])).read_text()) == r["reverse_score"]


def test_kth_driver_native_clstm_with_images(dev, tmp_path):
    """The KTH driver with the ConvLSTM (hidden 4, the shipped config): Grad-CAM at the recurrent layer with archType
    'CLSTM' (SURVEY 3.7 bug 8 repaired), reverse perturbation, image triptychs written through the GPU visualisation."""
    import FindMasksComparison_I3D_KTH as drv
    import visualisation
    from test_gpu_clstm import build
    from oracle import synthetic
    model, sd = build(4, softmax=False)
    model = torch.nn.DataParallel(model, [0]).to(dev).eval()
    x = synthetic.clips(2, kind="square", t=32, h=120, w=160).floor()
    cfg = {"gradCamType": "guessed", "splitType": "original", "conv_model": "models.CLSTM_4", "batch_size": 2}
    tags = ["person17_boxing_d1_1", "person03_boxing_d1_1"]
    masks = drv.find_masks([(x.to(torch.uint8), torch.tensor([0, 0]), tags)], model, cfg, 0.02, 0.04, 3, 1, "central",
                           "reverse", doGradCam=True, runTempMask=True, sub_dir="k", out_root=str(tmp_path), verbose=False,
                           viz=visualisation.driver_hook(160, 120))
    tm = pickle.load(open(tmp_path / "results" / "I3d_KTH_allTimeMaskResults_original_k.p", "rb"))
    gc = pickle.load(open(tmp_path / "results" / "I3d_KTH_allGradCamResults_original_k.p", "rb"))
    assert len(masks) == 1 and len(tm) == 1 and tm[0]["video_id"] == tags[0] and tm[0]["time_mask"].shape == (32,)
    assert gc[0]["GCHeatMap"].shape == (32, 120, 160)
    folder = [p for p in (tmp_path / "cam_saved_images" / "k" / "0").iterdir()][0] / "combined"
    names = sorted(os.listdir(folder))
    assert "img01.jpg" in names and "img32.jpg" in names and "casefreeze%s_0.png" % tags[0] in names
    assert "MASKVALScasereverse%s.txt" % tags[0] in names
    import cv2
    img = cv2.imread(str(folder / "img01.jpg"))
    assert img.shape == (120, 480, 3)
