"""Drop-in for the reference's visualisation module (pt/visualisation.py): create_image_arrays,
vizualize_results_on_gradcam, find_temp_mask_red_dots, vizualize_results - the step right after the hot path.

The per-frame host loop of the reference (cv2.applyColorMap, float blend, max-normalise, concatenate, the perturbed
clip recomputed per FRAME, :96-122) is one libivf kernel per clip (ivf_viz_triptych: colour map, blend, per-frame
maximum, the three panels and the temporal-mask dots), and the perturbed clip is computed once.  Only the encoding
(JPEG via cv2, PNG via PIL, the GIF via ImageMagick `convert` when installed) runs on the host.
Repairs (SURVEY 3.7 bug 7): `perturb_sequence` is imported, `args` is not needed."""
import os
import shutil

import numpy as np
import torch

try:
    from interpreting_video_features_b200 import _lib, ops
except ImportError:
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from interpreting_video_features_b200 import _lib, ops

from mask import perturb_sequence  # the drop-in next to this file


def find_temp_mask_red_dots(imageWidth, imageHeight, mask, roundUpMask):
    """pt/visualisation.py:67-93: one dot per frame; rounds the caller's mask in place when roundUpMask."""
    n = len(mask)
    dot_w = int(imageWidth // (n + 4))
    pad = int((imageWidth - dot_w * n) // n)
    dot_h = int(imageHeight // 20)
    dots = []
    for i in range(n):
        if roundUpMask:
            mask[i] = 1 if mask[i] > 0.5 else 0
        dots.append({'yStart': -dot_h, 'yEnd': imageHeight, 'xStart': i * (dot_w + pad),
                     'xEnd': i * (dot_w + pad) + dot_w, 'channel': 1 if mask[i] == 0 else 2})
    return dots


def triptych(clip, cam, time_mask, perturbation_type, draw_dots=True):
    """clip [3,T,H,W] (fp32 0..255 or uint8, host or device), cam [T,H,W] float32 (numpy or tensor), time_mask [T] ->
    uint8 numpy [T, H, 3W, 3] BGR: [frame | heat-map blend | perturbed frame (+ dots)], computed on the GPU."""
    dev = clip.device if clip.is_cuda else torch.device("cuda")
    x = clip.to(dev).contiguous()
    xf = x if x.dtype == torch.float32 else x.float()
    cam_d = torch.as_tensor(np.ascontiguousarray(cam) if isinstance(cam, np.ndarray) else cam,
                            dtype=torch.float32).to(dev).contiguous()
    m = torch.as_tensor(time_mask, dtype=torch.float32).detach().clone().to(dev)
    pert = perturb_sequence(xf[None], m.clone(), perturbation_type=perturbation_type, snap_values=True)[0].contiguous()
    t, hh, ww = cam_d.shape
    out = torch.empty((t, hh, 3 * ww, 3), dtype=torch.uint8, device=dev)
    ops.viz_triptych(x, cam_d, pert, m, out, draw_dots=draw_dots)
    host = torch.empty(out.shape, dtype=torch.uint8, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return host.numpy()


def create_image_arrays(input_sequence, gradcamMask, timeMask, intraBidx, temporalMaskType, output_folder, targTag,
                        RESIZE_FLAG, RESIZE_SIZE_WIDTH, RESIZE_SIZE_HEIGHT):
    """pt/visualisation.py:96-130: writes img%02d.jpg (the triptych per frame), mygif.gif, the dotted PNGs and the
    mask text file; returns the [3, T, H, 3W] array the reference returns (BGR planes, dots drawn)."""
    import cv2
    if RESIZE_FLAG:
        raise _lib.IvfError("RESIZE_FLAG=1 is not used by the drivers (both pass 0) and has no native kernel")
    os.makedirs(output_folder, exist_ok=True)
    clip = input_sequence[intraBidx]
    plain = triptych(clip, gradcamMask, timeMask, temporalMaskType, draw_dots=False)
    for i in range(plain.shape[0]):
        cv2.imwrite(os.path.join(output_folder, "img%02d.jpg" % (i + 1)), plain[i])
    if shutil.which("convert"):  # ImageMagick, as the reference's os.system call (:123-125)
        os.system("convert -delay 10 -loop 0 {}.jpg {}".format(os.path.join(output_folder, "*"),
                                                               os.path.join(output_folder, "mygif.gif")))
    dotted = triptych(clip, gradcamMask, timeMask, temporalMaskType, draw_dots=True)
    case = temporalMaskType + targTag
    from PIL import Image
    for i in range(dotted.shape[0]):  # :60-61, channel order reversed to RGB for PIL
        Image.fromarray(dotted[i][:, :, ::-1].copy(), mode="RGB").save(os.path.join(output_folder, "case%s_%d.png" % (case, i)))
    rounded = (torch.as_tensor(timeMask).detach().float().cpu() > 0.5).float()
    with open(os.path.join(output_folder, "MASKVALScase" + case + ".txt"), "w+") as f:
        f.write(str(rounded))
    return np.transpose(dotted, (3, 0, 1, 2))


def vizualize_results_on_gradcam(gradCamImage, mask, rootDir, case="0", roundUpMask=True, imageWidth=224,
                                 imageHeight=224):
    """pt/visualisation.py:35-64 on an already assembled [3,T,H,3W] array (host data: kept for callers that build
    their own arrays; create_image_arrays draws the dots on the GPU)."""
    from PIL import Image
    m = torch.as_tensor(mask).detach().float().cpu().clone()
    os.makedirs(rootDir, exist_ok=True)
    dots = find_temp_mask_red_dots(imageWidth, imageHeight, m, roundUpMask)
    off = imageWidth * 2
    for i in range(len(m)):
        for j, dot in enumerate(dots):
            gradCamImage[:, i, dot["yStart"]:, off + dot["xStart"]:off + dot["xEnd"]] = 0
            gradCamImage[dot["channel"], i, dot["yStart"]:, off + dot["xStart"]:off + dot["xEnd"]] = 255 if i == j else 150
        Image.fromarray(gradCamImage[::-1, i].transpose(1, 2, 0).astype(np.uint8), mode="RGB").save(
            rootDir + "/case" + case + "_" + str(i) + ".png")
    with open(rootDir + "/MASKVALScase" + case + ".txt", "w+") as f:
        f.write(str(m))


def vizualize_results(orig_seq, pert_seq, mask, rootDir=None, case="0", markImgs=True, iterTest=False):
    """pt/visualisation.py:8-32: the perturbed frames as PNGs with the mask value marked in the corner."""
    from PIL import Image
    root = (rootDir if rootDir is not None else "vizualisations/") + "/PerturbImgs/"
    os.makedirs(root, exist_ok=True)
    pert = pert_seq.detach().float().cpu().numpy().copy()
    mvals = torch.as_tensor(mask).detach().float().cpu()
    for i in range(pert.shape[1]):
        if markImgs:
            pert[1:, i, :10, :10] = 0
            pert[0, i, :10, :10] = float(mvals[i]) * 255
        Image.fromarray(pert[:, i].transpose(1, 2, 0).astype(np.uint8)).save(root + "case" + case + "pert" + str(i) + ".png")
    with open(root + "case" + case + ".txt", "w+") as f:
        f.write(str(mvals))


def driver_hook(width, height):
    """The callable drivers.process_batch invokes per clip and perturbation type."""
    def hook(clip, cam, time_mask, kind, folder, tag, backend):
        create_image_arrays(clip[None], cam, torch.as_tensor(time_mask), 0, kind, folder, tag, 0, width, height)
    return hook
