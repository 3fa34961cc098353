"""CPU: the repaired drivers' RESULT SCHEMA, file layout, command line and config files against what the reference
drivers' source declares (tests/golden/result_schema.json, extracted with `ast` by oracle/pin_result_schema.py), the
clip loaders on generated JPEG frames, and the visualisation helpers that stay on the host.  The compute backend is a
stub here (no GPU); tests/test_gpu_drivers.py runs the same drivers on the native path."""
import json
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from common import GOLD, REPO

PT = os.path.join(REPO, "interpreting_video_features_b200", "pt")
if PT not in sys.path:
    sys.path.insert(0, PT)

SCHEMA = json.load(open(os.path.join(GOLD, "result_schema.json")))


class StubBackend:
    """Deterministic stand-in for the native compute: scores from the clip mean, a fixed mask, a ramp heat map."""

    def __init__(self, ncls, t, hw):
        self.ncls, self.t, self.hw = ncls, t, hw
        self.searched = []

    def forward(self, clips):
        base = clips.float().mean(dim=(1, 2, 3, 4))
        logits = torch.stack([torch.roll(torch.arange(self.ncls).float(), int(b) % self.ncls) for b in base]) * 0.3
        return torch.softmax(logits, dim=1)

    def search(self, clips, targets, lam1, lam2, n_iter, perturb):
        n = len(clips)
        self.searched.append((n, [int(t) for t in targets], lam1, lam2, n_iter, perturb))
        tm = np.tile(np.linspace(0.01, 0.99, self.t, dtype=np.float32), (n, 1))
        return dict(time_mask=tm, freeze_score=np.full(n, 0.25, np.float32), reverse_score=np.full(n, 0.5, np.float32))

    def gradcam(self, clips, targets):
        h, w = self.hw
        return np.tile(np.linspace(0, 1, self.t * h * w, dtype=np.float32).reshape(1, self.t, h, w), (len(clips), 1, 1, 1))


def test_smth_driver_schema_files_and_pickles(tmp_path):
    import FindMasksComparison_I3D_smth as drv
    ncls, t = 10, 4
    g = torch.Generator().manual_seed(0)
    batches = [(torch.rand((3, 3, t, 6, 8), generator=g) * 255, torch.tensor([1, 4, 7]), ["101", "102", "103"]),
               (torch.rand((3, 3, t, 6, 8), generator=g) * 255, torch.tensor([4, 4, 2]), ["104", "105", "106"])]
    csv = tmp_path / "subset.csv"
    csv.write_text("4,7\n102,103\n105,\n")  # class 4: clips 102, 105; class 7: clip 103
    backend = StubBackend(ncls, t, (6, 8))
    masks = drv.find_masks(batches, None, {"gradCamType": "guessed", "batch_size": 3}, 0.01, 0.02, 7, "central", "freeze",
                           classOI=str(csv), doGradCam=True, runTempMask=True, backend=backend, sub_dir="unit",
                           out_root=str(tmp_path), verbose=False)
    assert len(masks) == 3
    assert [s[0] for s in backend.searched] == [2, 1] and backend.searched[0][2:] == (0.01, 0.02, 7, "freeze")
    frag = SCHEMA["smth"]["name_fragments"]
    tm_path = tmp_path / "results" / ("allTimeMaskResults_unit_subset.csv_.p")
    gc_path = tmp_path / "results" / ("allGradCamResults_unit_subset.csv_.p")
    assert "allTimeMaskResults_" in frag and "allGradCamResults_" in frag
    tm = pickle.load(open(tm_path, "rb"))
    gc = pickle.load(open(gc_path, "rb"))
    assert [list(r.keys()) for r in tm] == [SCHEMA["smth"]["appended_dict_keys"]["clips_time_mask_results"]] * 3
    assert [list(r.keys()) for r in gc] == [SCHEMA["smth"]["appended_dict_keys"]["clips_grad_cam_results"]] * 3
    assert [r["video_id"] for r in tm] == ["102", "103", "105"] and [r["video_id"] for r in gc] == [102, 103, 105]
    for r in tm:
        assert isinstance(r["true_class"], int) and isinstance(r["pred_class"], int)
        assert isinstance(r["time_mask"], np.ndarray) and r["time_mask"].shape == (t,) and r["time_mask"].dtype == np.float32
        assert all(isinstance(r[k], float) for k in ("original_score_guess", "original_score_true", "freeze_score",
                                                     "reverse_score"))
        assert r["freeze_score"] == 0.25 and r["reverse_score"] == 0.5
    for r in gc:
        assert r["GCHeatMap"].shape == (t, 6, 8) and r["GCHeatMap"].dtype == np.float32
    # folder: cam_saved_images/<subDir>/<true>/<id>g_<pred>_gs%5.4f_cs%5.4f/combined + the two class-score files
    r = tm[0]
    folder = tmp_path / "cam_saved_images" / "unit" / str(r["true_class"]) / (
        "102g_%d_gs%5.4f_cs%5.4f" % (r["pred_class"], r["original_score_guess"], r["original_score_true"])) / "combined"
    assert folder.is_dir()
    assert float((folder / "ClassScoreFreezecase102.txt").read_text()) == 0.25
    assert float((folder / "ClassScoreReversecase102.txt").read_text()) == 0.5
    # 'guessed' searched the predicted class, and pred is the argmax of the model output
    out0 = backend.forward(batches[0][0])
    assert backend.searched[0][1] == [int(out0[1].argmax()), int(out0[2].argmax())]


def test_kth_driver_schema_selection_and_label_targets(tmp_path):
    import FindMasksComparison_I3D_KTH as drv
    assert len(drv.clips_of_interest("original")) == 24 and ["person17", "boxing", "d1", "_1"] in drv.clips_of_interest("original")
    assert ["person09", "running", "d2", "_1"] in drv.clips_of_interest("other")
    t = 4
    g = torch.Generator().manual_seed(1)
    tags = ["person17_boxing_d1_1\n", "person01_boxing_d1_1", "person25_walking_d4_1", "person24_running_d2_2"]
    batches = [(torch.rand((4, 3, t, 6, 8), generator=g) * 255, torch.tensor([0, 0, 5, 4]), tags)]
    backend = StubBackend(6, t, (6, 8))
    cfg = {"gradCamType": "true", "splitType": "original", "conv_model": "models.I3D_doubled_kth", "batch_size": 4}
    drv.find_masks(batches, None, cfg, 0.02, 0.04, 5, 1, "central", "reverse", doGradCam=True, runTempMask=True,
                   backend=backend, sub_dir="k", out_root=str(tmp_path), verbose=False)
    assert backend.searched == [(2, [0, 5], 0.02, 0.04, 5, "reverse")]  # the labels, not the guesses
    tm = pickle.load(open(tmp_path / "results" / "I3d_KTH_allTimeMaskResults_original_k.p", "rb"))
    gc = pickle.load(open(tmp_path / "results" / "I3d_KTH_allGradCamResults_original_k.p", "rb"))
    assert [list(r.keys()) for r in tm] == [SCHEMA["kth"]["appended_dict_keys"]["clipsTimeMaskResults"]] * 2
    assert [list(r.keys()) for r in gc] == [SCHEMA["kth"]["appended_dict_keys"]["clipsGradCamResults"]] * 2
    assert [r["video_id"] for r in tm] == ["person17_boxing_d1_1", "person25_walking_d4_1"]
    assert any("I3d_KTH_allTimeMaskResults_original_" in f for f in SCHEMA["kth"]["name_fragments"])


def test_cli_flags_and_configs_cover_the_reference():
    import utils
    parser = utils.build_parser()
    have = set()
    for a in parser._actions:
        have.update(a.option_strings)
    for flags in SCHEMA["cli_flags"]:
        for f in flags:
            if f.startswith("-"):
                assert f in have, "reference flag %s missing from the drop-in parser" % f
    args = utils.load_args(["-c", os.path.join(PT, "configs", "config_i3d_smth.py"), "--use_cuda", "-g", "0,1",
                            "--gradCamType", "true", "--optIter", "5"])
    assert args.mod_stride_layers == "" and args.optIter == 5 and args.subsetFile is None
    cfg = utils.merged_config(args)
    assert cfg["gradCamType"] == "true" and cfg["maskPerturbType"] == "freeze" and cfg["conv_model"] == "models.I3D_doubled"
    for name, keys in SCHEMA["config_keys"].items():
        ours = utils.load_module(os.path.join(PT, "configs", name)).config
        missing = [k for k in keys if k not in ours]
        assert not missing, (name, missing)
        for k in ("maskPerturbType", "gradCamType", "splitType"):  # bug 3
            assert k in ours
    dev, ids = utils.setup_cuda_devices(args)
    assert ids == [0, 1] and dev.type == "cuda"
    sd = utils.remove_module_from_checkpoint_state_dict({"module.a.w": 1, "b": 2})
    assert list(sd.items()) == [("a.w", 1), ("b", 2)]


def _write_frames(folder, t, h, w, seed):
    from PIL import Image
    os.makedirs(folder, exist_ok=True)
    rs = np.random.RandomState(seed)
    frames = []
    for i in range(t):
        arr = (rs.rand(h, w, 3) * 255).astype(np.uint8)
        Image.fromarray(arr).save(os.path.join(folder, "frame%02d.jpg" % (i + 1)), quality=95)
        with Image.open(os.path.join(folder, "frame%02d.jpg" % (i + 1))) as im:
            frames.append(np.frombuffer(im.tobytes(), dtype=np.uint8).reshape(h, w, 3))  # the reference's decode
    return np.array(frames)


def test_loaders_decode_like_the_reference(tmp_path):
    """pt/data_loader_jpg.py:23-41 / pt/data_loader_kth.py:20-43: [3,T,H,W] float 0..255 (and the same values as
    uint8 with as_uint8=True), labels and ids."""
    from data_loader_jpg import ImLoader
    from data_loader_kth import KTHImLoader
    t, h, w = 3, 10, 12
    want = _write_frames(str(tmp_path / "smth" / "7" / "4242"), t, h, w, 0)
    _write_frames(str(tmp_path / "smth" / "12" / "17"), t, h, w, 1)
    ds = ImLoader(str(tmp_path / "smth"), clip_size=t, get_item_id=True)
    assert len(ds) == 2 and sorted(ds.classes) == [7, 12]
    i7 = [i for i, it in enumerate(ds.path_data) if it.id == "4242"][0]
    data, label, vid = ds[i7]
    assert data.dtype == torch.float32 and tuple(data.shape) == (3, t, h, w) and label == 7 and vid == "4242"
    assert torch.equal(data, torch.from_numpy(want).float().permute(3, 0, 1, 2))
    d8, _, _ = ImLoader(str(tmp_path / "smth"), clip_size=t, get_item_id=True, as_uint8=True)[i7]
    assert d8.dtype == torch.uint8 and torch.equal(d8.float(), data)
    wk = _write_frames(str(tmp_path / "kth" / "0"), t, h, w, 2)
    (tmp_path / "kth" / "0" / "class.txt").write_text("3\n")
    (tmp_path / "kth" / "0" / "label.txt").write_text("person17_boxing_d1_1\n")
    dk = KTHImLoader(str(tmp_path / "kth"), clip_size=t, get_item_id=True)
    data, label, tag = dk[0]
    assert len(dk) == 1 and label == 3 and tag.strip() == "person17_boxing_d1_1"
    assert torch.equal(data, torch.from_numpy(wk).float().permute(3, 0, 1, 2))


def test_red_dots_geometry_matches_oracle():
    import visualisation
    from oracle import viz_oracle
    mask = torch.tensor([0.2, 0.7, 0.9, 0.4, 0.6, 0.1, 0.55, 0.3])
    dots = visualisation.find_temp_mask_red_dots(64, 40, mask.clone(), True)
    imgs = np.zeros((8, 40, 192, 3), np.uint8)
    drawn = viz_oracle.draw_dots(imgs, mask.numpy(), 64, 40)
    for i, d in enumerate(dots):
        x0 = 128 + d["xStart"]
        assert d["yStart"] == -2 and d["xEnd"] - d["xStart"] == 64 // 12
        assert d["channel"] == (2 if mask[i] > 0.5 else 1)
        assert drawn[i, -1, x0, d["channel"]] == 255 and drawn[(i + 1) % 8, -1, x0, d["channel"]] == 150
        assert drawn[i, -3, x0].sum() == 0
