"""Oracle (test infrastructure): the image triptych of pt/visualisation.py:96-122 (create_image_arrays) and the
temporal-mask dots of :35-93, restated in numpy with the colour map taken from cv2 itself.  CPU only."""
import numpy as np


def jet_lut():
    import cv2
    return cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)


def triptych(clip, cam, perturbed, lut=None):
    """clip/perturbed float32 [3,T,H,W] RGB 0..255, cam float32 [T,H,W] -> uint8 [T,H,3W,3] BGR (:99-120)."""
    lut = jet_lut() if lut is None else lut
    frames = np.flip(np.transpose(clip, (1, 2, 3, 0)), 3)  # [T,H,W,3] BGR
    out = []
    for i in range(frames.shape[0]):
        img = frames[i]
        with np.errstate(invalid="ignore"):
            idx = np.uint8(np.nan_to_num(255 * cam[i], nan=0.0))
        heat = np.float32(lut[idx])
        blend = heat + np.float32(img)
        blend = blend / np.max(blend)
        pert = np.uint8(np.transpose(perturbed[:, i], (1, 2, 0)))[:, :, ::-1]
        out.append(np.concatenate((np.uint8(img), np.uint8(255 * blend), pert), axis=1))
    return np.array(out)


def draw_dots(images, mask, width, height, round_up=True):
    """images uint8 [T,H,3W,3] (modified copy returned); mask [T] (:35-93)."""
    images = images.copy()
    m = np.array(mask, dtype=np.float32)
    n = len(m)
    dot_w = int(width // (n + 4))
    pad = int((width - dot_w * n) // n)
    dot_h = int(height // 20)
    if round_up:
        m = (m > 0.5).astype(np.float32)
    off = 2 * width
    for i in range(n):
        for j in range(n):
            x0 = off + j * (dot_w + pad)
            ch = 1 if m[j] == 0 else 2
            images[i, -dot_h:, x0:x0 + dot_w, :] = 0
            images[i, -dot_h:, x0:x0 + dot_w, ch] = 255 if i == j else 150
    return images
