"""Frame decoding shared by the two loaders: <dir>/frame01.jpg ... -> one clip tensor [3, T, H, W].

The reference decodes with PIL into uint8 [T,H,W,3], calls .float() and permutes (pt/data_loader_jpg.py:26-37).
Here the clip can stay uint8 (as_uint8=True): it is a quarter of the bytes in the DataLoader's pinned buffers and
over PCIe, and the engines convert on the device (ivf_u8_to_f32) - the values are the same 0..255 integers."""
import os

import numpy as np
import torch
from PIL import Image


def read_clip(directory, clip_size, as_uint8=False, pattern="frame%02d.jpg"):
    frames = []
    for i in range(clip_size):
        with Image.open(os.path.join(directory, pattern % (i + 1))) as im:
            arr = np.asarray(im if im.mode == "RGB" else im.convert("RGB"), dtype=np.uint8)
        frames.append(arr)
    clip = torch.from_numpy(np.stack(frames))       # [T, H, W, 3] uint8
    clip = clip.permute(3, 0, 1, 2).contiguous()    # [3, T, H, W] (Batch, Channel, T, H, W after collation)
    return clip if as_uint8 else clip.float()
