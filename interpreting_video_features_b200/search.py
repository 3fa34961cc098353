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
        self.m = ops.zeros((B, T), torch.float32, dev)        # raw mask (Adam parameter)
        self.sig = ops.zeros((B, T), torch.float32, dev)      # sigmoid(m)
        self.exp_avg = ops.zeros((B, T), torch.float32, dev)
        self.exp_avg_sq = ops.zeros((B, T), torch.float32, dev)
        self.step = ops.zeros((B,), torch.int32, dev)
        self.losses = ops.zeros((B, 3), torch.float32, dev)
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

    @_lib.on_device
    def gradcam_lowres(self, targets):
        """Un-normalised low-resolution Grad-CAM maps [B, T', h, w] of the clips in the static input buffers for
        the given classes (forward replayed from the engines' graphs, head backward, fused kernel)."""
        if getattr(self, "_cam_low", None) is None:
            a = self.eng.acts["Mixed_5c"]
            self._cam_low = torch.empty((self.B, a.d, a.h, a.w), dtype=torch.float32, device=self.device)
        for e, sl in zip(self.engs, self.slices):
            e.gradcam(targets[sl], None, True, cam=None, lowres=self._cam_low[sl])
        return self._cam_low

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
        """Batched pt/mask.py:103-169; returns raw masks [B,T] on the device and the unperturbed probs.
        'central': every candidate window is one batched forward whose class scores stay on the device
        (ivf_select_scores); ivf_init_mask_select then applies the reference's stopping rule per clip.  Nothing is
        read back, so the host queues the whole search behind the initialisation without waiting for it."""
        B, T, dev = self.B, self.T, self.device
        self.set_targets(targets)
        probs_orig = self.forward(None, "freeze")
        if mode == "random":
            m = (torch.rand((B, T), generator=generator) > 0.7).float()
            m = (m - 0.5) * 5
            for b in range(B):
                if abs(float(m[b].sum())) == 2.5 * T:
                    m[b, 8] += 0.1
            return m.to(dev), probs_orig
        if mode != "central":
            raise ValueError("mode must be 'central' or 'random'")
        ncand = max(T // 2, 1)  # fully frozen + the centred windows i = 1 .. T/2-1
        scores = torch.empty((1 + ncand, B), dtype=torch.float32, device=dev)
        if getattr(self, "_cand", None) is None:  # candidate masks, uploaded once per searcher
            cand = torch.ones((ncand, T))
            for i in range(1, T // 2):
                cand[i, :i] = 0
                cand[i, T - i:] = 0
            self._cand = cand.to(dev)

        def score_row(j):
            for e, sl in zip(self.engs, self.slices):
                ops.select_scores(e.probs, e._targets, scores[j, sl])

        score_row(0)
        for j in range(ncand):
            for e in self.engs:
                e.forward(self._cand[j], "freeze" if j == 0 else self.perturb)
            score_row(1 + j)
        raw = torch.empty((B, T), dtype=torch.float32, device=dev)
        self.init_choice = torch.empty(B, dtype=torch.int32, device=dev)
        ops.init_mask_select(scores, T, self.threshold, raw, self.init_choice)
        self.init_scores = scores
        return raw, probs_orig

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
        ops.fill_zero(self.exp_avg)
        ops.fill_zero(self.exp_avg_sq)
        ops.fill_zero(self.step)
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
    if groups <= 1 or not hasattr(model, "MAX_ENGINE_GEOMETRIES"):  # the ConvLSTM model keeps one engine
        return [model._engine(x, batch=micro_batch)]
    per = micro_batch // groups
    return [model._engine(x, batch=per, tag=g) for g in range(groups)]


class _Stager:
    """Micro-batches of host clips -> the engine's static input buffer without a pageable bounce: two pinned
    staging buffers filled by a host gather (clips[sel] of arbitrary indices), copied H2D on the search's
    stream; an event per buffer keeps the host from refilling it before its copy has run.  Pinned clips whose
    selection is one contiguous range skip the staging copy."""

    def __init__(self, clips, micro_batch):
        self.clips = clips
        self.on_host = not clips.is_cuda
        self.bufs, self.events, self.k = [None, None], [None, None], 0
        self.shape = (micro_batch,) + tuple(clips.shape[1:])

    def get(self, sel):
        if not self.on_host:
            return self.clips[sel]
        lo, n = sel[0], len(sel)
        if self.clips.is_pinned() and sel == list(range(lo, lo + n)):
            return self.clips[lo:lo + n]
        k = self.k = self.k ^ 1
        if self.bufs[k] is None:
            self.bufs[k] = torch.empty(self.shape, dtype=self.clips.dtype, pin_memory=True)
        elif self.events[k] is not None:
            self.events[k].synchronize()
        torch.index_select(self.clips, 0, torch.as_tensor(sel), out=self.bufs[k])
        return self.bufs[k]

    def used(self, device):
        if self.on_host and self.bufs[self.k] is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(device))
            self.events[self.k] = ev


def assemble_results(local, mine, n_total, world, t, ncls, cam_shape=None, stats=None, device=None):
    """Per-rank result rows [n_local, t + 2 + ncls (+ cam elements)] -> the result dict; with torch.distributed
    initialised and world > 1 the rows of all ranks are all_gather-ed into clip order first (NCCL on the GPUs, gloo
    in the CPU tests).  stats receives 'gather_seconds' (device-timed on CUDA) and 'gathered_bytes'."""
    import torch.distributed as dist
    indices = list(mine)
    gather_s = 0.0
    if world > 1 and dist.is_available() and dist.is_initialized():
        timed = stats is not None and local.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(local.device))
        full = gather_rows(local, mine, n_total, world)
        if timed:
            e1.record(torch.cuda.current_stream(local.device))
            e1.synchronize()
            gather_s = e0.elapsed_time(e1) * 1e-3
        indices = list(range(n_total))
    else:  # single rank, or a shard computed without a process group (rows follow `indices`)
        full = local
    if stats is not None:
        stats.update(gather_seconds=gather_s, gathered_bytes=int(full.numel() * 4))
    out = dict(time_mask=full[:, :t], freeze_score=full[:, t], reverse_score=full[:, t + 1],
               probs_orig=full[:, t + 2:t + 2 + ncls], indices=indices)
    if cam_shape is not None:
        out["cam_lowres"] = full[:, t + 2 + ncls:].reshape(-1, *cam_shape)
    return out


def find_masks_batched(model, clips, targets, lam1=0.01, lam2=0.02, n_iter=300, perturb="freeze",
                       init="central", threshold=0.9, micro_batch=8, lr=0.2, use_graph=True, rank=0, world=1,
                       device=None, groups=None, gradcam=False, n_total=None, stats=None):
    """Mask search (and, with gradcam=True, the Grad-CAM the drivers run per clip,
    pt/FindMasksComparison_I3D_smth.py:257-269) over this rank's shard of the clips; returns a dict of [N, ...]
    device tensors, gathered over ranks when torch.distributed is initialised and world > 1.

    clips: [N,3,T,H,W] host or device, fp32 0..255 or uint8 - every rank holds all N and works on r::W; or, with
    n_total given, only this rank's clips in the order of shard_indices(n_total, rank, world) (a loader that
    decodes just its shard).  targets follow clips.  `model` is a drop-in models.I3D_doubled[_kth].Model in eval
    mode.  gradcam=True adds cam_lowres [N, T', h, w]: the target class's un-normalised low-resolution map at
    Mixed_5c (upsampling and normalisation are one kernel per clip wherever the maps are consumed).
    Nothing in the loop reads back from the device: the host runs ahead of the GPU by whole micro-batches.
    stats (dict, optional) receives 'gather_seconds' (device-timed) and 'micro_batches'."""
    import torch.distributed as dist
    device = torch.device(device if device is not None else "cuda")
    N_in, C, T, H, W = clips.shape
    if n_total is None:
        N = N_in
        mine = shard_indices(N, rank, world)
        local_of = {g: g for g in mine}  # global index -> row of `clips`
    else:
        N = int(n_total)
        mine = shard_indices(N, rank, world)
        assert N_in == len(mine), "n_total: clips must hold exactly this rank's shard (%d != %d)" % (N_in, len(mine))
        local_of = {g: i for i, g in enumerate(mine)}
    stager = _Stager(clips, micro_batch)
    targets = torch.as_tensor(targets)
    searcher = None
    rows = []
    n_mb = 0
    for s in range(0, len(mine), micro_batch):
        sel = [local_of[g] for g in mine[s:s + micro_batch]]
        n_valid = len(sel)
        if n_valid < micro_batch:  # ragged tail: repeat the last clip, drop the duplicates afterwards
            sel = sel + [sel[-1]] * (micro_batch - n_valid)
        x = stager.get(sel)
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
        stager.used(device)
        cols = [res["time_mask"], res["freeze_score"][:, None], res["reverse_score"][:, None], res["probs_orig"]]
        if gradcam:  # the clips are still in the engines' static buffers
            cols.append(searcher.gradcam_lowres(tg).flatten(1))
        rows.append(torch.cat(cols, dim=1)[:n_valid])
        n_mb += 1
    ncls = getattr(model, "_num_classes", None) or model.num_classes  # I3D / CLSTM_4 attribute names
    if rows:
        local = torch.cat(rows)
    else:  # a rank without clips still takes part in the gather: it needs the row width
        cam_elems = 1
        for k in model.avg_pool.kernel_size:  # the Mixed_5c map is exactly the average pool's window (engine check)
            cam_elems *= int(k)
        local = torch.zeros((0, T + 2 + ncls + (cam_elems if gradcam else 0)), device=device)
    out = assemble_results(local, mine, N, world, T, ncls, [int(k) for k in model.avg_pool.kernel_size] if gradcam else None,
                           stats=stats, device=device)
    if stats is not None:
        stats["micro_batches"] = n_mb
    return out
