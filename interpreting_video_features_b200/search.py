"""Batched temporal-mask search: the hot loop of pt/FindMasksComparison_I3D_smth.py:166-251
(find_masks) for B independent (clip, mask) pairs per launch sequence, the iteration captured in a
CUDA graph, clips sharded over ranks.

Per clip this computes exactly what the reference computes (init_mask 'central' pt/mask.py:121-154,
N Adam steps on loss = lam1*|s| + lam2*TV(s) + p[target], final sigmoid, freeze score = class score
of the last iteration, reverse score pt/...smth.py:234-235) — but
  * each clip of a micro-batch carries its own mask (samples are independent in eval mode, so the
    reference's full-batch forward under one mask wastes (B-1)/B of its work);
  * all T/2+1 init_mask candidates are evaluated as batched forwards with one host read-back
    (the reference syncs once per candidate, pt/mask.py:144);
  * one iteration = perturb -> I3D forward -> head -> head' -> I3D data-gradient -> perturb' ->
    loss/Adam, ~150 launches replayed from a CUDA graph (the reference issues 8 451 ATen ops).
Multi-GPU: rank r takes clips r::W; no collective inside the search; one all_gather of the result
rows at the end.
"""
import numpy as np
import torch

from . import _lib, ops


class MaskSearch:
    """Mask search over micro-batches of clips on one GPU.

    `engine` is one runner for the whole micro-batch, or a list of runners over consecutive clip groups
    (their batch sizes add up to the micro-batch).  Groups are independent (each clip owns its mask), so
    their iterations are issued on separate streams and captured as parallel branches of ONE CUDA graph: the
    small-grid kernels of one group (7x7 / 14x14 stages, head, loss) fill SMs the other group leaves idle."""

    def __init__(self, engine, lam1=0.01, lam2=0.02, lr=0.2, n_iter=300, perturb="freeze", threshold=0.9,
                 use_graph=True):
        self.engs = list(engine) if isinstance(engine, (list, tuple)) else [engine]
        self.eng = self.engs[0]
        self.lam1, self.lam2, self.lr, self.n_iter = float(lam1), float(lam2), float(lr), int(n_iter)
        self.perturb, self.threshold, self.use_graph = perturb, float(threshold), use_graph
        self.B = sum(e.B for e in self.engs)
        self.T, self.device = self.eng.T, self.eng.device
        self.slices, off = [], 0
        for e in self.engs:
            self.slices.append(slice(off, off + e.B))
            off += e.B
        B, T, dev = self.B, self.T, self.device
        self.m = torch.zeros((B, T), dtype=torch.float32, device=dev)        # raw mask (Adam parameter)
        self.sig = torch.zeros((B, T), dtype=torch.float32, device=dev)      # sigmoid(m)
        self.exp_avg = torch.zeros_like(self.m)
        self.exp_avg_sq = torch.zeros_like(self.m)
        self.step = torch.zeros(B, dtype=torch.int32, device=dev)
        self.losses = torch.zeros((B, 3), dtype=torch.float32, device=dev)
        self.graph = None
        self.launches_per_iter = None
        self._gstreams = None

    # ---- the clip groups behind one interface
    def set_input(self, x):
        for e, sl in zip(self.engs, self.slices):
            e.set_input(x[sl])

    def set_targets(self, tg):
        for e, sl in zip(self.engs, self.slices):
            e.set_targets(tg[sl])

    @_lib.on_device
    def forward(self, mask, perturb):
        """mask: None, [T] (shared) or [B,T]; returns a COPY of the [B, classes] outputs."""
        outs = []
        for e, sl in zip(self.engs, self.slices):
            mg = mask if (mask is None or mask.dim() == 1) else mask[sl]
            outs.append(e.forward(mg, perturb))
        return torch.cat(outs) if len(outs) > 1 else outs[0].clone()

    def probs(self):
        return torch.cat([e.probs for e in self.engs]) if len(self.engs) > 1 else self.eng.probs

    def dm(self):
        return torch.cat([e.dm for e in self.engs]) if len(self.engs) > 1 else self.eng.dm

    def _group_iteration(self, e, sl):
        e.forward(self.sig[sl], self.perturb)
        dm = e.backward(to_mask=True)
        ops.mask_loss_adam(self.m[sl], self.exp_avg[sl], self.exp_avg_sq[sl], dm, 0, self.lam1, self.lam2, self.lr,
                           losses=self.losses[sl], sig_out=self.sig[sl], step_dev=self.step[sl])

    # one iteration on the static buffers (what the graph captures)
    def _iteration(self):
        if len(self.engs) == 1:
            self._group_iteration(self.eng, self.slices[0])
            return
        main = torch.cuda.current_stream(self.device)
        if self._gstreams is None:
            self._gstreams = [torch.cuda.Stream(device=self.device) for _ in self.engs]
        ev = torch.cuda.Event()
        ev.record(main)
        for e, sl, gs in zip(self.engs, self.slices, self._gstreams):
            gs.wait_event(ev)
            with torch.cuda.stream(gs):
                self._group_iteration(e, sl)
        for gs in self._gstreams:
            main.wait_stream(gs)

    @_lib.on_device
    def _capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):  # warm-up outside capture (first-use attribute calls, tensor maps)
            n0 = _lib.launch_count(self.device)
            self._iteration()
            self.launches_per_iter = _lib.launch_count(self.device) - n0
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._iteration()
        self.graph = g

    @_lib.on_device
    def init_masks(self, targets, mode="central", generator=None):
        """Batched pt/mask.py:103-169; returns raw masks [B,T] on the device and the unperturbed probs."""
        B, T, dev = self.B, self.T, self.device
        idx = torch.arange(B, device=dev)
        tg = targets.to(dev).long()
        probs_orig = self.forward(None, "freeze")
        if mode == "random":
            m = (torch.rand((B, T), generator=generator) > 0.7).float()
            m = (m - 0.5) * 5
            for b in range(B):
                if abs(float(m[b].sum())) == 2.5 * T:
                    m[b, 8] += 0.1
            return m.to(dev), probs_orig
        cand = [torch.ones(T)]  # fully frozen
        for i in range(1, T // 2):
            c = torch.ones(T)
            c[:i] = 0
            c[-i:] = 0
            cand.append(c)
        scores = [probs_orig[idx, tg]]
        for c in cand:
            scores.append(self.forward(c.to(dev), self.perturb if c is not cand[0] else "freeze")[idx, tg])
        sc = torch.stack(scores).cpu().numpy()  # [2 + ncand, B]; the one host read-back of init
        orig, frozen, cen = sc[0], sc[1], sc[2:]
        raw = np.empty((B, T), dtype=np.float32)
        for b in range(B):
            chosen = None
            for k in range(cen.shape[0]):
                chosen = k
                ratio = np.float32(orig[b] - cen[k, b]) / np.float32(orig[b] - frozen[b])
                if ratio < self.threshold:
                    break
            row = np.ones(T, dtype=np.float32) if chosen is None else cand[chosen + 1].numpy()
            raw[b] = np.where(row == 0, -5.0, 5.0)
        return torch.from_numpy(raw).to(dev), probs_orig

    @_lib.on_device
    def run(self, x, targets, init="central", raw_masks=None, n_iter=None, record=None):
        """x fp32 [B,3,T,H,W] on the device; targets [B].  Returns a dict of device tensors."""
        n_iter = self.n_iter if n_iter is None else n_iter
        dev = self.device
        idx = torch.arange(self.B, device=dev)
        tg = targets.to(dev).long()
        self.set_input(x)
        self.set_targets(tg)
        if raw_masks is None:
            raw_masks, probs_orig = self.init_masks(tg, init)
        else:
            probs_orig = self.forward(None, "freeze")
        self.m.copy_(raw_masks)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step.zero_()
        ops.sigmoid(self.m, self.sig)
        if self.use_graph and self.graph is None and n_iter > 0:
            saved = [t.clone() for t in (self.m, self.sig, self.exp_avg, self.exp_avg_sq, self.step)]
            self._capture()
            for t, s in zip((self.m, self.sig, self.exp_avg, self.exp_avg_sq, self.step), saved):
                t.copy_(s)
        for _ in range(n_iter):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._iteration()
            if record is not None:
                record.setdefault("class", []).append(self.probs()[idx, tg].clone())
                record.setdefault("dm_class", []).append(self.dm().clone())
                record.setdefault("loss_reg", []).append(self.losses[:, 2].clone())
                record.setdefault("mask", []).append(self.m.clone())
        freeze_score = self.probs()[idx, tg].clone() if n_iter > 0 else probs_orig[idx, tg]
        final = self.sig.clone()
        reverse_score = self.forward(final, "reverse")[idx, tg]
        return dict(time_mask=final, raw_mask=self.m.clone(), freeze_score=freeze_score,
                    reverse_score=reverse_score, probs_orig=probs_orig, init_mask=raw_masks)


def shard_indices(n, rank, world):
    """Clip indices of `rank` (clip-parallel): r, r+W, r+2W, ..."""
    return list(range(rank, n, world))


def gather_rows(local_rows, local_idx, n_total, world, group=None):
    """all_gather per-clip result rows [n_local, k] into [n_total, k] in clip order.  Works on the
    NCCL and gloo backends; pads ranks to equal length."""
    import torch.distributed as dist
    k = local_rows.shape[1]
    per = (n_total + world - 1) // world
    pad_rows = torch.zeros((per, k), dtype=local_rows.dtype, device=local_rows.device)
    pad_idx = torch.full((per,), -1, dtype=torch.int64, device=local_rows.device)
    pad_rows[:local_rows.shape[0]] = local_rows
    pad_idx[:len(local_idx)] = torch.as_tensor(local_idx, dtype=torch.int64, device=local_rows.device)
    rows = [torch.empty_like(pad_rows) for _ in range(world)]
    idxs = [torch.empty_like(pad_idx) for _ in range(world)]
    dist.all_gather(rows, pad_rows, group=group)
    dist.all_gather(idxs, pad_idx, group=group)
    out = torch.zeros((n_total, k), dtype=local_rows.dtype, device=local_rows.device)
    for r, i in zip(rows, idxs):
        ok = i >= 0
        out[i[ok]] = r[ok]
    return out


def default_groups(micro_batch):
    """Clip groups per micro-batch: IVF_GROUPS, default 1.  Measured on C2 (8 clips): 2 groups 2.990 ms per
    step against 2.992 ms for one - the persistent convolution kernels already occupy every SM - and 4 groups
    3.36 ms, so grouping stays an option for small micro-batches of large models rather than the default."""
    import os
    g = int(os.environ.get("IVF_GROUPS", "0"))
    if g <= 0:
        g = 1
    while micro_batch % g:
        g -= 1
    return max(g, 1)


def make_engines(model, x, micro_batch, groups):
    """`groups` runners of micro_batch/groups clips each (one runner when groups == 1)."""
    if groups <= 1:
        return [model._engine(x, batch=micro_batch)]
    per = micro_batch // groups
    return [model._engine(x, batch=per, tag=g) for g in range(groups)]


def find_masks_batched(model, clips, targets, lam1=0.01, lam2=0.02, n_iter=300, perturb="freeze",
                       init="central", threshold=0.9, micro_batch=8, lr=0.2, use_graph=True, rank=0, world=1,
                       device=None, groups=None):
    """Mask search over `clips` [N,3,T,H,W] (host or device) for this rank's shard; returns a dict of
    [N, ...] tensors (gathered over ranks when torch.distributed is initialised and world > 1).
    `model` is a drop-in models.I3D_doubled[_kth].Model in eval mode."""
    device = torch.device(device if device is not None else "cuda")
    N, C, T, H, W = clips.shape
    mine = shard_indices(N, rank, world)
    searcher = None
    rows = []
    for s in range(0, len(mine), micro_batch):
        sel = mine[s:s + micro_batch]
        n_valid = len(sel)
        if n_valid < micro_batch:  # ragged tail: repeat the last clip, drop the duplicates afterwards
            sel = sel + [sel[-1]] * (micro_batch - n_valid)
        x = clips[sel].to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        tg = targets[sel]
        if searcher is None:
            engs = make_engines(model, x, micro_batch, default_groups(micro_batch) if groups is None else groups)
            # the searcher (mask/Adam buffers + the captured iteration) lives with its engines: a later call with
            # the same hyper-parameters replays the same graph instead of capturing again (~19 ms per call)
            key = (tuple(id(e) for e in engs), float(lam1), float(lam2), float(lr), perturb, float(threshold),
                   bool(use_graph))
            cache = engs[0].__dict__.setdefault("_searchers", {})
            searcher = cache.get(key)
            if searcher is None:
                cache.clear()  # one set of search buffers per engine set
                searcher = cache[key] = MaskSearch(engs, lam1, lam2, lr, n_iter, perturb, threshold, use_graph)
            searcher.n_iter = int(n_iter)  # the captured graph is one iteration: the count is free
        res = searcher.run(x, tg, init=init)
        row = torch.cat([res["time_mask"], res["freeze_score"][:, None], res["reverse_score"][:, None],
                         res["probs_orig"]], dim=1)[:n_valid]
        rows.append(row)
    ncls = model._num_classes
    local = torch.cat(rows) if rows else torch.zeros((0, T + 2 + ncls), device=device)
    import torch.distributed as dist
    indices = mine
    if world > 1 and dist.is_available() and dist.is_initialized():
        full = gather_rows(local, mine, N, world)
        indices = list(range(N))
    else:  # single rank, or a shard computed without a process group (rows follow `indices`)
        full = local
    return dict(time_mask=full[:, :T], freeze_score=full[:, T], reverse_score=full[:, T + 1],
                probs_orig=full[:, T + 2:], indices=indices)
