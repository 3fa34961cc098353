"""Oracle/test infrastructure: deterministic synthetic clips (seeds model 0,
data 1 + clip index; uniform 0..255 like the loaders, pt/data_loader_jpg.py:27-37) plus a
structured 'moving square' clip with real temporal content."""
import torch


def uniform_clip(index, c=3, t=16, h=224, w=224):
    g = torch.Generator().manual_seed(1 + index)
    return torch.rand((c, t, h, w), generator=g) * 255.0


def moving_square_clip(index, c=3, t=16, h=224, w=224):
    """A bright square moving diagonally over a seeded low-amplitude background."""
    g = torch.Generator().manual_seed(1001 + index)
    x = torch.rand((c, t, h, w), generator=g) * 40.0
    side = max(h, w) // 5
    for u in range(t):
        y0 = int((h - side) * u / max(t - 1, 1))
        x0 = int((w - side) * ((u * (index + 2)) % t) / max(t - 1, 1))
        x[:, u, y0:y0 + side, x0:x0 + side] += 200.0
    return x.clamp_(0, 255)


def clips(n, kind="uniform", **kw):
    f = uniform_clip if kind == "uniform" else moving_square_clip
    return torch.stack([f(i, **kw) for i in range(n)])


def uniform_clip_u8(index, c=3, t=16, h=224, w=224):
    """The same clip as decoded frames: uint8 0..255 (pt/data_loader_jpg.py:27-30 reads uint8 frames and calls
    .float() on them); .float() of this is what the reference path sees."""
    return uniform_clip(index, c, t, h, w).to(torch.uint8)
