"""The per-clip interpretation loop of the reference drivers (pt/FindMasksComparison_I3D_smth.py:125-315,
pt/FindMasksComparison_I3D_KTH.py:126-380) on the native fast path, with the reference's RESULT SCHEMA:

  * per clip of interest: class prediction, temporal-mask search (init_mask 'central', N Adam steps), freeze score
    (class score of the last iteration), reverse score, Grad-CAM at the mask's target class;
  * clips_time_mask_results: list of dicts with the keys of smth.py:243-251 / KTH.py:290-298;
    clips_grad_cam_results: list of dicts with the keys of smth.py:272-277 / KTH.py:330-334;
  * cam_saved_images/<subDir>/<true class>/<id>g_<pred>_gs%5.4f_cs%5.4f/combined/ClassScore{Freeze,Reverse}case<id>.txt
    (smth.py:222-239), optional image triptychs, and the two pickles (smth.py:306-313, KTH.py:372-378).

What differs from the reference's execution, not from its results: all selected clips of a loader batch are searched
together (one mask per clip), the model is evaluated once per batch, Grad-CAM runs batched.  Repairs
of the shipped drivers are listed in SURVEY §3.7 (numbers in comments below); one more: smth.py:218 stores
int(max probability) as 'original_score_guess' (always 0 for a probability) while KTH.py:294 and the folder name of
smth.py:289 use the float - the float is kept here.

`backend` makes the compute injectable: the CPU test suite checks schema, paths and pickles with a stub backend;
NativeBackend is the product path (libivf kernels, no CPU fallback).
"""
import os
import pickle

import numpy as np
import torch


class NativeBackend:
    """model: a drop-in models.I3D_doubled[_kth].Model / models.CLSTM_4.Model (possibly DataParallel-wrapped)."""

    def __init__(self, model, arch="I3D", cam_size=(224, 224), micro_batch=8, device=None):
        from .pt.grad_cam_videos import GradCamVideo
        self.model = model.module if hasattr(model, "module") else model
        self.arch = arch
        self.micro_batch = int(micro_batch)
        self.device = device if device is not None else next(self.model.parameters()).device
        layer = "Mixed_5c" if arch == "I3D" else "clstm"  # bug 8: the CLSTM child is `clstm`
        self.grad_cam = GradCamVideo(model=self.model, target_layer_names=[layer], class_dict=None, use_cuda=True,
                                     input_spatial_size=cam_size, normalizePerFrame=True, archType=arch)

    def forward(self, clips):
        """[B, classes] scores of the unperturbed clips (what `output = model(input_var)` returns)."""
        with torch.no_grad():
            x = clips.to(self.device, non_blocking=True)
            return self.model(x if x.dtype != torch.uint8 else x.float()).detach().float().cpu()

    def search(self, clips, targets, lam1, lam2, n_iter, perturb):
        from . import search
        res = search.find_masks_batched(self.model, clips, targets, lam1=lam1, lam2=lam2, n_iter=n_iter,
                                        perturb=perturb, init="central", micro_batch=min(self.micro_batch, max(len(clips), 1)),
                                        device=self.device)
        return {k: res[k].detach().cpu().numpy() for k in ("time_mask", "freeze_score", "reverse_score")}

    def gradcam(self, clips, targets):
        cams = []
        for s in range(0, len(clips), self.micro_batch):
            c, _ = self.grad_cam.batched(clips[s:s + self.micro_batch], [int(t) for t in targets[s:s + self.micro_batch]])
            cams.append(c)
        return np.concatenate(cams) if cams else np.zeros((0,), dtype=np.float32)

    def perturbed(self, clip, time_mask, perturb):
        """The clip under the snapped mask, for the image triptych (pt/visualisation.py:113-116)."""
        from .pt import mask as M
        m = torch.as_tensor(time_mask, dtype=torch.float32, device=self.device).clone()
        x = clip[None].to(self.device)
        return M.perturb_sequence(x if x.dtype != torch.uint8 else x.float(), m, perturbation_type=perturb,
                                  snap_values=True)[0].detach().cpu()


def score_folder(sub_dir, true_class, video_id, pred_class, score_guess, score_true, root="cam_saved_images"):
    """smth.py:222-225 / KTH.py:273-277."""
    return os.path.join(root, str(sub_dir), str(true_class),
                        str(video_id) + "g_" + str(pred_class) + "_gs%5.4f" % score_guess + "_cs%5.4f" % score_true,
                        "combined")


def time_mask_record(true_class, pred_class, video_id, time_mask, score_guess, score_true, freeze, reverse):
    """Keys and value types of smth.py:243-251."""
    return {'true_class': true_class, 'pred_class': pred_class, 'video_id': video_id,
            'time_mask': np.asarray(time_mask, dtype=np.float32), 'original_score_guess': float(score_guess),
            'original_score_true': float(score_true), 'freeze_score': float(freeze), 'reverse_score': float(reverse)}


def grad_cam_record(true_class, pred_class, video_id, heat_map):
    """Keys of smth.py:272-277."""
    return {'true_class': true_class, 'pred_class': pred_class, 'video_id': video_id,
            'GCHeatMap': np.asarray(heat_map, dtype=np.float32)}


def process_batch(backend, sequence, labels, video_ids, selected, grad_cam_type, lam1, lam2, n_iter, perturb, sub_dir,
                  run_temp_mask=True, do_grad_cam=True, out_root=".", video_id_cast=str, viz=None, verbose=True):
    """One loader batch: sequence [B,3,T,H,W] (fp32 0..255 or uint8), labels [B], video_ids list[B]; `selected` =
    batch indices of the clips of interest.  Returns (time-mask records, Grad-CAM records, masks)."""
    tm_records, gc_records, masks = [], [], []
    if not selected:
        return tm_records, gc_records, masks
    labels = [int(v) for v in labels]
    output = backend.forward(sequence)  # once per batch (the reference recomputes it per clip, smth.py:176)
    pred = output.argmax(dim=1).tolist()
    # "guessed": the mask and the CAM explain the predicted class, else the label (smth.py:179-184,266-267)
    targets = [pred[b] if grad_cam_type == "guessed" else labels[b] for b in selected]
    clips = sequence[selected]
    found = backend.search(clips, torch.tensor(targets), lam1, lam2, n_iter, perturb) if run_temp_mask else None
    cams = backend.gradcam(clips, targets) if do_grad_cam else None
    for j, b in enumerate(selected):
        vid = video_ids[b]
        true_class, pred_class = labels[b], int(pred[b])
        score_guess, score_true = float(output[b].max()), float(output[b, labels[b]])
        folder = os.path.join(out_root, score_folder(sub_dir, true_class, vid, pred_class, score_guess, score_true))
        os.makedirs(folder, exist_ok=True)
        if found is not None:
            time_mask = found["time_mask"][j]
            with open(os.path.join(folder, "ClassScoreFreezecase" + str(vid) + ".txt"), "w+") as f:
                f.write(str(float(found["freeze_score"][j])))
            with open(os.path.join(folder, "ClassScoreReversecase" + str(vid) + ".txt"), "w+") as f:
                f.write(str(float(found["reverse_score"][j])))
            tm_records.append(time_mask_record(true_class, pred_class, vid, time_mask, score_guess, score_true,
                                               found["freeze_score"][j], found["reverse_score"][j]))
            masks.append(torch.from_numpy(np.asarray(time_mask)))
            if verbose:
                print("resulting mask is: ", np.round(time_mask, 3))
        if cams is not None:
            gc_records.append(grad_cam_record(true_class, pred_class, video_id_cast(vid), cams[j]))
        if viz is not None and cams is not None and found is not None:
            for kind in ("freeze", "reverse"):  # smth.py:296-301
                viz(sequence[b], cams[j], found["time_mask"][j], kind, folder, str(vid), backend)
    return tm_records, gc_records, masks


def dump_results(tm_records, gc_records, tm_path, gc_path):
    """smth.py:306-313 / KTH.py:372-378: two pickles of lists of dicts."""
    os.makedirs(os.path.dirname(tm_path) or ".", exist_ok=True)
    with open(tm_path, "wb") as f:
        pickle.dump(tm_records, f)
    with open(gc_path, "wb") as f:
        pickle.dump(gc_records, f)
