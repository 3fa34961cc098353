"""Drop-in for the reference's mask module (pt/mask.py): same four functions, same signatures and
semantics, computed by libivf kernels.  The per-call surface cannot remove the reference's
(B-1)/B dead work (one mask for a whole batch, one clip read); the batched fast path lives in
interpreting_video_features_b200.search.
"""
import torch

try:
    from interpreting_video_features_b200 import _lib, ops
except ImportError:  # used from a sys.path that only contains pt/
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from interpreting_video_features_b200 import _lib, ops


class _PerturbFn(torch.autograd.Function):
    """pt/mask.py:11-56 as one fused scan kernel each way."""

    @staticmethod
    def forward(ctx, seq, mask, mode):
        if not seq.is_cuda:
            raise _lib.IvfError("perturb_sequence: the native path runs on a B200 only (no CPU fallback)")
        x = seq.detach().float().contiguous()
        m = mask.detach().float().contiguous().to(x.device)
        out = torch.empty_like(x)
        ops.perturb_fwd(x, m, mode, _lib.PFMT_NCDHW_F32, out)
        ctx.x, ctx.m, ctx.mode = x, m, mode
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.needs_input_grad[0]:
            raise _lib.IvfError("perturb_sequence: gradient w.r.t. the clip is not on the hot path")
        x, m = ctx.x, ctx.m
        dm = torch.empty((x.shape[0], x.shape[2]), dtype=torch.float32, device=x.device)
        ops.perturb_bwd(x, m, ctx.mode, _lib.PFMT_NCDHW_F32, gout.float().contiguous(), dm)
        if m.dim() == 1:  # one mask for the whole batch (the reference's semantics): sum over clips
            dm = dm.sum(dim=0) if dm.shape[0] > 1 else dm[0]
        return None, dm, None


def perturb_sequence(seq, mask, perturbation_type='freeze', snap_values=False):
    """seq [B,C,T,H,W] float 0..255; mask [T].  pt/mask.py:4-56."""
    if snap_values:  # in place on the caller's mask, pt/mask.py:5-10
        with torch.no_grad():
            mask.copy_((mask > 0.5).to(mask.dtype))
    if perturbation_type not in ('freeze', 'reverse'):
        raise ValueError("perturbation_type must be 'freeze' or 'reverse'")
    return _PerturbFn.apply(seq, mask, perturbation_type)


def find_submasks_from_mask(mask, thresh=0.1):
    """Runs of consecutive mask > thresh, as lists of frame indices (pt/mask.py:60-85).  Host-side
    result by contract; one device->host copy (the reference makes one per frame)."""
    vals = torch.as_tensor(mask).detach().float().cpu().tolist()
    thresh = float(torch.tensor(thresh, dtype=torch.float32))  # fp32 comparison, as the reference's tensor > float
    submasks, cur = [], None
    for j, v in enumerate(vals):
        if v > thresh:
            if cur is None:
                cur = []
            cur.append(j)
        elif cur is not None:
            submasks.append(cur)
            cur = None
    if cur is not None:
        submasks.append(cur)
    return submasks


class _TvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask, p, q):
        if not mask.is_cuda:
            raise _lib.IvfError("calc_tv_norm: the native path runs on a B200 only (no CPU fallback)")
        m = mask.detach().float().contiguous()
        val = torch.empty(1, dtype=torch.float32, device=m.device)
        dm = torch.empty_like(m)
        ops.tv_norm(m, p, q, val, dm)
        ctx.dm = dm
        return val[0]

    @staticmethod
    def backward(ctx, g):
        return g * ctx.dm, None, None


def calc_tv_norm(mask, p=3, q=3):
    """Total-variation norm of pt/mask.py:88-100 (value and gradient in one kernel)."""
    return _TvFn.apply(mask, float(p), float(q))


def init_mask(seq, model, batch_index, target, threshold=0.9, mode='central', mask_type='freeze'):
    """pt/mask.py:103-169.  'central': the first centred window whose class-score drop ratio falls
    below `threshold` (or the last one tried), mapped 0 -> -5 / 1 -> +5; 'random': U > 0.7 -> +-2.5."""
    T = seq.shape[2]
    dev = seq.device
    if mode == "central":
        with torch.no_grad():
            tgt = int(target[batch_index])
            frozen = perturb_sequence(seq, torch.ones(T, device=dev), 'freeze')  # every frame = frame 0
            fully_frozen_score = model(frozen)[batch_index, tgt]
            orig_score = model(seq)[batch_index, tgt]
            new_mask = torch.ones(T, device=dev)
            for i in range(1, T // 2):
                new_mask = torch.ones(T, device=dev)
                new_mask[:i] = 0
                new_mask[-i:] = 0
                central_score = model(perturb_sequence(seq, new_mask, perturbation_type=mask_type))[batch_index, tgt]
                score_ratio = (orig_score - central_score) / (orig_score - fully_frozen_score)
                if score_ratio < threshold:
                    break
            mask = torch.where(new_mask == 0, torch.full_like(new_mask, -5.0), torch.full_like(new_mask, 5.0))
    elif mode == "random":
        mask = (torch.rand(T, device=dev) > 0.7).float()
        mask = (mask - 0.5) * 5
        if torch.abs(mask.sum()) == 2.5 * len(mask):
            mask[8] += 0.1
    else:
        raise ValueError("mode must be 'central' or 'random'")
    mask = mask.clone()
    mask.requires_grad_()
    print("initial mask is: ", mask)
    return mask
