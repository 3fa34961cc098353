"""Oracle (test infrastructure): temporal-mask operations, restated from pt/mask.py and the
optimisation loop of pt/FindMasksComparison_I3D_smth.py:188-216.  CPU, fp32, torch autograd.
"""
import torch


def find_submasks_from_mask(mask, thresh=0.1):
    """pt/mask.py:60-85 — maximal runs of mask > thresh (strict)."""
    # the reference compares a float32 tensor element with the Python float: fp32 comparison
    vals = torch.as_tensor(mask, dtype=torch.float32).tolist()
    thresh = float(torch.tensor(thresh, dtype=torch.float32))
    runs, cur = [], None
    for j, v in enumerate(vals):
        if v > thresh:
            if cur is None:
                cur = []
            cur.append(j)
        elif cur is not None:
            runs.append(cur)
            cur = None
    if cur is not None:
        runs.append(cur)
    return runs


def perturb_sequence(seq, mask, perturbation_type="freeze", snap_values=False):
    """pt/mask.py:4-56.  seq [B,C,T,H,W]; mask [T] (differentiable)."""
    if snap_values:  # :5-10, in place on the caller's mask
        with torch.no_grad():
            mask.copy_((mask > 0.5).to(mask.dtype))
    T = seq.shape[2]
    frames = []
    if perturbation_type == "freeze":  # :11-22
        prev = None
        for u in range(T):
            cur = seq[:, :, u] if u == 0 else (1 - mask[u]) * seq[:, :, u] + mask[u] * prev
            frames.append(cur)
            prev = cur
    elif perturbation_type == "reverse":  # :24-56
        frames = [seq[:, :, y] for y in range(T)]
        for run in find_submasks_from_mask(mask.detach(), 0.1):
            for u in range(len(run) // 2):
                i, j = run[u], run[-(u + 1)]
                frames[i] = (1 - mask[i]) * seq[:, :, i] + mask[i] * seq[:, :, j]
                frames[j] = (1 - mask[i]) * seq[:, :, j] + mask[i] * seq[:, :, i]
    else:
        raise ValueError(perturbation_type)
    return torch.stack(frames, dim=2)


def calc_tv_norm(mask, p=3, q=3):
    """pt/mask.py:88-100."""
    val = 0
    for u in range(1, len(mask) - 1):
        val = val + torch.abs(mask[u - 1] - mask[u]) ** p
        val = val + torch.abs(mask[u + 1] - mask[u]) ** p
    val = val ** (1 / p)
    val = val ** q
    return val


def init_mask(seq, model, batch_index, target, threshold=0.9, mode="central", mask_type="freeze",
              generator=None):
    """pt/mask.py:103-169 with the CUDA-only allocations (:131,135,158) made device-agnostic.
    Returns the raw (pre-sigmoid) mask: the first candidate that FAILS the threshold (or the last
    one tried), mapped 0 -> -5, 1 -> +5."""
    T = seq.shape[2]
    tgt = int(target[batch_index])
    if mode == "central":
        with torch.no_grad():
            frozen = seq[:, :, :1].expand(-1, -1, T, -1, -1).contiguous()
            frozen_score = model(frozen)[batch_index, tgt]
            orig_score = model(seq)[batch_index, tgt]
            new_mask = torch.ones(T)
            for i in range(1, T // 2):
                new_mask = torch.ones(T)
                new_mask[:i] = 0
                new_mask[-i:] = 0
                central = model(perturb_sequence(seq, new_mask, mask_type))[batch_index, tgt]
                ratio = (orig_score - central) / (orig_score - frozen_score)
                if ratio < threshold:
                    break
        mask = torch.where(new_mask == 0, torch.full_like(new_mask, -5.0), torch.full_like(new_mask, 5.0))
    elif mode == "random":
        mask = (torch.rand(T, generator=generator) > 0.7).float()
        mask = (mask - 0.5) * 5
        if torch.abs(mask.sum()) == 2.5 * len(mask):
            mask[8] += 0.1
    else:
        raise ValueError(mode)
    return mask.clone().requires_grad_()


def mask_search(seq, model, batch_index, target, time_mask, lam1, lam2, n_iter, mask_type="freeze", lr=0.2,
                record=None):
    """pt/FindMasksComparison_I3D_smth.py:191-216: Adam on the raw mask; the early stop never fires
    (oldLoss is never updated, :192,209).  record: optional dict of lists (loss, class, grad, mask)."""
    tgt = int(target[batch_index])
    opt = torch.optim.Adam([time_mask], lr=lr)
    class_loss = None
    for _ in range(n_iter):
        mask_clip = torch.sigmoid(time_mask)
        l1 = lam1 * torch.sum(torch.abs(mask_clip))
        tv = lam2 * calc_tv_norm(mask_clip, 3, 3)
        class_loss = model(perturb_sequence(seq, mask_clip, mask_type))[batch_index, tgt]
        loss = l1 + tv + class_loss
        opt.zero_grad()
        loss.backward()
        if record is not None:
            record.setdefault("loss", []).append(float(loss))
            record.setdefault("class", []).append(float(class_loss))
            record.setdefault("grad", []).append(time_mask.grad.detach().clone())
        opt.step()
        if record is not None:
            record.setdefault("mask", []).append(time_mask.detach().clone())
    return torch.sigmoid(time_mask).detach(), (float(class_loss) if class_loss is not None else None)
