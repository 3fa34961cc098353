"""Oracle (test infrastructure): Grad-CAM tail restated from pt/grad_cam_videos.py:64-142 in numpy,
with cv2.resize(INTER_LINEAR) restated as an explicit half-pixel bilinear so the oracle has no
OpenCV dependency (pinned against cv2 in oracle/pin_against_reference.py).
"""
import numpy as np
import torch


def resize_bilinear(src, dsize):
    """cv2.resize(src, dsize=(width, height)) for float32, INTER_LINEAR: source coordinate
    (d + 0.5) * scale - 0.5, floor, replicated border (opencv modules/imgproc/src/resize.cpp)."""
    w_out, h_out = dsize
    h_in, w_in = src.shape

    def taps(n_out, n_in):
        scale = n_in / n_out
        f = ((np.arange(n_out) + 0.5) * scale - 0.5).astype(np.float32)
        i0 = np.floor(f).astype(np.int64)
        frac = (f - i0).astype(np.float32)
        lo = i0 < 0
        i0[lo], frac[lo] = 0, 0.0
        hi = i0 >= n_in - 1
        i0[hi], frac[hi] = n_in - 1, 0.0
        i1 = np.minimum(i0 + 1, n_in - 1)
        return i0, i1, frac

    y0, y1, fy = taps(h_out, h_in)
    x0, x1, fx = taps(w_out, w_in)
    src = src.astype(np.float32)
    rows0 = src[y0][:, x0] * (1 - fx)[None, :] + src[y0][:, x1] * fx[None, :]
    rows1 = src[y1][:, x0] * (1 - fx)[None, :] + src[y1][:, x1] * fx[None, :]
    return (rows0 * (1 - fy)[:, None] + rows1 * fy[:, None]).astype(np.float32)


def cam_from_features(target, grads_val, clip_size, input_spatial_size, normalize_per_frame=True):
    """pt/grad_cam_videos.py:96-140.  target [C,T',h,w] activations, grads_val [1,C,T',h,w]."""
    weights = np.mean(grads_val, axis=(2, 3, 4))[0, :]
    cam = np.zeros(target.shape[1:], dtype=np.float32)
    for i, w in enumerate(weights):
        cam += w * target[i]
    cam = np.maximum(cam, 0)
    step = clip_size // target.shape[1]
    cam_vid = []
    for i in range(cam.shape[0]):
        m = resize_bilinear(cam[i], (input_spatial_size[0], input_spatial_size[1]))
        cam_vid.append(np.repeat(np.expand_dims(m, 0), step, axis=0))
    cam_vid = np.array(cam_vid)
    with np.errstate(invalid="ignore", divide="ignore"):
        if normalize_per_frame:
            for i in range(cam_vid.shape[0]):
                cam_vid[i] = cam_vid[i] - np.min(cam_vid[i])
                cam_vid[i] = cam_vid[i] / np.max(cam_vid[i])
        else:
            cam_vid = cam_vid - np.min(cam_vid)
            cam_vid = cam_vid / np.max(cam_vid)
    if cam_vid.shape[0] > 1:
        cam_vid = np.concatenate(cam_vid, axis=0)
    if cam_vid.shape[0] == 1:
        cam_vid = np.squeeze(cam_vid, 0)
    return cam_vid, cam


def gradcam_i3d(sd, x, index=None, input_spatial_size=(224, 224), normalize_per_frame=True,
                avg_pool=(2, 7, 7), softmax=True, quant=False, layer="Mixed_5c"):
    """pt/grad_cam_videos.py:27-43,64-98 for archType 'I3D'; the target layer is any endpoint the reference's
    FeatureExtractor can hook (pt/pytorch-grad-cam/grad-cam.py:23-31: children of the model by name), the drivers
    use Mixed_5c.  x [1,3,T,H,W]; returns (cam [T,H,W] float32, output [1,classes], lowres cam)."""
    from . import i3d_oracle

    feat, _ = i3d_oracle.features(sd, x, upto=layer, quant=quant)
    feat = feat.detach().requires_grad_(True)
    top = feat if layer == "Mixed_5c" else i3d_oracle.features(sd, feat, quant=quant, start_after=layer)[0]
    output = i3d_oracle.head(sd, top, avg_pool, softmax)
    if index is None:
        index = int(np.argmax(output.detach().cpu().numpy()))
    score = output[0, int(index)]  # sum(one_hot * output)
    (grad,) = torch.autograd.grad(score, feat)
    cam, lowres = cam_from_features(feat.detach().cpu().numpy()[0], grad.cpu().numpy(), x.shape[2],
                                    input_spatial_size, normalize_per_frame)
    return cam, output.detach(), lowres


def gradcam_clstm(sd, x, index=None, input_spatial_size=(160, 120), normalize_per_frame=True, quant=False, **kw):
    """Grad-CAM of the ConvLSTM classifier as the reference intends it (pt/pytorch-grad-cam/grad-cam.py:33-49
    with the child names repaired, SURVEY bug 8; pt/grad_cam_videos.py:87-91): target = the stacked
    effective-step outputs [E,B,C,h,w], gradient of the SOFTMAX score (the walked children include `sm`)
    w.r.t. a detached copy of that stack — non-zero for the last step only, because the classifier reads
    output[-1]."""
    import torch.nn.functional as F
    from . import clstm_oracle

    _, outputs = clstm_oracle.forward(sd, x, return_outputs=True, quant=quant, **kw)
    agg = torch.stack(outputs).detach().requires_grad_(True)  # [E,B,C,h,w]
    out = F.softmax(F.linear(agg[-1].reshape(agg.shape[1], -1), sd["endFC.weight"], sd["endFC.bias"]), dim=1)
    if index is None:
        index = int(np.argmax(out.detach().cpu().numpy()))
    (grads,) = torch.autograd.grad(out[0, int(index)], agg)
    grads_val = grads.detach().cpu().numpy().transpose(1, 2, 0, 3, 4)  # [B,C,E,h,w]
    target = agg.detach().permute(1, 2, 0, 3, 4).cpu().numpy()[0]
    cam, lowres = cam_from_features(target, grads_val, x.shape[2], input_spatial_size, normalize_per_frame)
    return cam, out.detach(), lowres
