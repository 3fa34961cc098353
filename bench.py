#!/usr/bin/env python
"""bench.py — mask-search throughput of the hot path (BASELINE.json metric, config C2).

One STEP = one temporal-mask-search iteration for a micro-batch of 8 synthetic 16x224x224 clips
(BASELINE.json configs[1]: I3D Something-Something-v2, 174 classes, freeze perturbation):
perturb -> I3D forward -> head -> head' -> I3D data-gradient -> perturb' -> L1/TV loss + Adam, for 8
independent (clip, mask) pairs = 8 clip-iterations.  Metric: clip-iterations/s, whole job over all
ranks (weak scaling: every rank owns 8 clips; no collective inside the search).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode bf16|fp32]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (N > 1)

value   : inputs resident in HBM, K CUDA-graph replays timed with CUDA events (max over ranks)
e2e     : the job BASELINE.json configs[3] describes (C4), scaled to the ranks present: 128 clips per GPU as
          uint8 frames in pinned HOST memory, sharded clip-parallel, through the public API
          (search.find_masks_batched): H2D of every micro-batch, init_mask (T/2+1 batched forwards), 300
          iterations, reverse score, Grad-CAM of every clip, the NCCL all_gather of masks + scores + low-res CAMs
          (N > 1) and the D2H of the gathered result — all inside the timed region; clip-iterations/s =
          clips x 300 / wall time (max over ranks)
roofline: the tcgen05 convolution kernel: algorithmic conv FLOPs (forward + data gradient, no weight
          gradient) / summed conv-kernel time measured with CUDA events around every conv launch
cpu_baseline / --impl reference: the oracle restatement of the reference's PyTorch-CPU path (the
          reference itself is Python and cannot travel to the GPU box), all host threads.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

CLIPS, T, H, W, NCLS, N_ITER = 8, 16, 224, 224, 174, 300
METRIC, UNIT = "mask_search_clip_iterations_per_sec", "clip-iterations/s"
CONFIG = {"workload": "C2: I3D smth (174 classes) temporal-mask search, freeze, batch 8 x 3x16x224x224 "
                      "synthetic clips, lam 0.01/0.02, Adam lr 0.2; one step = one iteration for the 8 clips",
          "clips_per_gpu": CLIPS, "clip": [3, T, H, W], "iterations_per_search": N_ITER,
          "l2": "per-step working set (~1.5 GB of activations for 8 clips) exceeds the 126 MB L2"}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def state_dict():
    from interpreting_video_features_b200.pt.models import I3D_doubled
    torch.manual_seed(0)
    m = quiet(I3D_doubled.Model, NCLS, last_stride=1, stride_mod_layers="", softMax=1).eval()
    return m


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md): one nvidia-smi
    process in loop mode (-lms 50), lines stamped on arrival so only those inside [t_begin, t_end] count."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None
        self.t_begin = self.t_end = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), [c.strip() for c in line.strip().split(",")]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        rows = [r for t, r in self.rows if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t))]
        if not rows:  # region shorter than the sampling period: the nearest samples
            rows = [r for _, r in self.rows[-3:]]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active")
                                                          for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_step(model_sd, x1, raw, target, opt_state):
    """One reference mask-search iteration for ONE clip on the CPU (oracle restatement of
    pt/FindMasksComparison_I3D_smth.py:198-214 with B = 1)."""
    from oracle import i3d_oracle, mask_oracle
    tm, opt = opt_state
    mc = torch.sigmoid(tm)
    loss = 0.01 * mc.abs().sum() + 0.02 * mask_oracle.calc_tv_norm(mc, 3, 3) + \
        i3d_oracle.forward(model_sd, mask_oracle.perturb_sequence(x1, mc, "freeze"))[0, target]
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss)


def time_cpu_reference(steps, warmup):
    from oracle import synthetic
    model = state_dict()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    x1 = synthetic.uniform_clip_u8(0)[None].float()  # decoded uint8 frames -> .float(), pt/data_loader_jpg.py:27-30
    tm = torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4, requires_grad=True)
    st = (tm, torch.optim.Adam([tm], lr=0.2))
    for _ in range(warmup):
        cpu_reference_step(sd, x1, None, 3, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sd, x1, None, 3, st)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps, cores


def run_reference(args, rank):
    if rank != 0:
        return
    val, per_step, cores = time_cpu_reference(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG, sample="one clip-iteration (B=1) per step"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps of one clip-iteration, oracle port of the reference's PyTorch-CPU "
                                       "path (torch %s, %d threads)" % (args.steps, torch.__version__, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def conv_time_per_step(engs):
    tot_s, tot_n = 0.0, 0
    for eng in engs:
        s_, n_ = conv_time_one_engine(eng)
        tot_s, tot_n = tot_s + s_, tot_n + n_
    return tot_s, tot_n


def conv_time_one_engine(eng):
    """Sum of the convolution kernels' durations in one eager iteration (CUDA events around every conv
    launch on the launching stream); returns (seconds, launches)."""
    from interpreting_video_features_b200 import ops
    events = []
    orig = ops.conv3d

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        events.append((e0, e1))
        return r

    orig_split = ops.conv1x1_split

    def timed_split(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_split(*a, **k)
        e1.record()
        events.append((e0, e1))
        return r

    orig_pair = ops.conv3d_pair

    def timed_pair(*a, **k):  # one grouped launch (or two plain ones): the two 3x3x3 branches of a module
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_pair(*a, **k)
        e1.record()
        events.append((e0, e1))
        return r

    ops.conv3d = timed
    ops.conv1x1_split = timed_split
    ops.conv3d_pair = timed_pair
    streams, eng.use_streams = eng.use_streams, False  # serialised: one kernel at a time on one stream
    try:
        # hold the stream for ~60 ms so that every launch and event of the iteration is already queued when
        # the GPU starts: the event pairs then bracket kernel time, not the host's launch latency
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.06 * 1.9e9))
        eng.forward(eng._mask, eng._perturb)
        eng.backward(to_mask=True)
        torch.cuda.synchronize()
    finally:
        ops.conv3d = orig
        ops.conv1x1_split = orig_split
        ops.conv3d_pair = orig_pair
        eng.use_streams = streams
    return sum(a.elapsed_time(b) for a, b in events) * 1e-3, len(events)


def gradcam_throughput(dev, rank, world, mode, batch=32, n_clips=256, reps=5, with_cpu=False):
    """Second half of the headline metric (config C1): Grad-CAM clips/s on I3D-KTH (6 classes), native 32x120x160
    clips, through the drop-in GradCamVideo.
      value : clips resident in HBM as fp32, CAMs left on the device (forward graph + head' + fused CAM kernel);
      e2e   : GradCamVideo.stream - uint8 frames in pinned host memory in, float32 [T,H,W] maps in pinned host
              memory out, H2D / compute / D2H on three streams;
      roofline (tensor): one I3D forward = 43.05 GFLOP per clip x clips/s against the bf16 peaks;
      cpu_baseline: the oracle port of pt/grad_cam_videos.py:64-142 on the host cores (rank 0, N = 1 only);
      geometry_224: the same weights on 32x224x224 clips with the head's pool widened to [4,7,7] - throughput only
              (the KTH head is ill-formed at 224x224)."""
    import torch.distributed as dist
    from interpreting_video_features_b200.pt.grad_cam_videos import GradCamVideo
    from interpreting_video_features_b200.pt.models import I3D_doubled_kth
    from oracle import gradcam_oracle, i3d_oracle, synthetic
    torch.manual_seed(0)
    m = quiet(I3D_doubled_kth.Model, 6, last_stride=1, stride_mod_layers="", softMax=1, finalTimeLength=4)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.to(dev).eval().set_mode(mode)
    gc = GradCamVideo(model=m, target_layer_names=['Mixed_5c'], class_dict=None, use_cuda=True,
                      input_spatial_size=(160, 120), normalizePerFrame=True, archType="I3D")
    uniq = torch.stack([synthetic.uniform_clip_u8(1000 + rank * batch + i, t=32, h=120, w=160) for i in range(batch)])
    host = uniq.repeat(n_clips // batch, 1, 1, 1, 1).pin_memory()  # decoded uint8 frames, as the loaders produce them
    xd = uniq.to(dev).float()
    eng = m._engine(xd, batch=batch)
    eng.set_input(xd)
    cam_dev = torch.empty((batch, 32, 120, 160), dtype=torch.float32, device=dev)
    for _ in range(3):
        eng.gradcam(None, (120, 160), True, cam=cam_dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.gradcam(None, (120, 160), True, cam=cam_dev)
    e1.record()
    torch.cuda.synchronize()
    dev_s = e0.elapsed_time(e1) * 1e-3 / reps
    gc.stream(host, None, batch=batch)  # untimed: stream objects, staging buffers, the pinned output block (629 MB of
    # cudaHostAlloc on first use; torch's caching host allocator hands the same block back to the timed call)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    cams, outs = gc.stream(host, None, batch=batch)
    e2e_s = time.perf_counter() - t0
    assert cams.shape == (n_clips, 32, 120, 160) and bool(np_isfinite_or_nan(cams))
    if world > 1:
        t = torch.tensor([dev_s, e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
    pk, pk_src = peaks()
    gf_clip = i3d_oracle.conv_flops_per_clip(sd, (32, 120, 160)) / 1e9
    val = world * batch / dev_s
    out = {"metric": "gradcam_clips_per_sec", "unit": "clips/s", "value": val, "e2e": world * n_clips / e2e_s,
           "clips_per_gpu": n_clips, "batch": batch, "ms_per_batch_device": dev_s * 1e3,
           "h2d_bytes_per_clip": int(host[0].numel()), "d2h_bytes_per_clip": int(cams[0].size * 4),
           "roofline": {"bound": "tensor", "achieved": val / world * gf_clip / 1e3, "unit": "TFLOP/s",
                        "peak": pk["bf16_tflops"], "frac": val / world * gf_clip / 1e3 / pk["bf16_tflops"],
                        "algorithmic_gflop_per_clip": gf_clip, "peak_source": pk_src},
           "workload": "C1: I3D KTH (6 classes) Grad-CAM at Mixed_5c, 32x120x160 synthetic clips (the model's native "
                       "geometry, SURVEY fact 10), %d clips per launch sequence; e2e: %d uint8 clips from pinned host "
                       "memory through GradCamVideo.stream" % (batch, n_clips)}
    if with_cpu and rank == 0:
        x1 = uniq[:1].float()
        gradcam_oracle.gradcam_i3d(sd, x1, None, (160, 120), True, avg_pool=(4, 4, 5))
        t0 = time.perf_counter()
        for i in range(2):
            gradcam_oracle.gradcam_i3d(sd, uniq[i:i + 1].float(), None, (160, 120), True, avg_pool=(4, 4, 5))
        dt = (time.perf_counter() - t0) / 2
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "2 timed clips (32x120x160) of the oracle port of GradCamVideo after 1 warm-up"}
    # throughput-only 32x224x224 variant
    m.avg_pool.kernel_size = [4, 7, 7]
    b224 = 8
    x224 = torch.stack([synthetic.uniform_clip_u8(3000 + i, t=32, h=224, w=224) for i in range(b224)]).to(dev).float()
    eng2 = m._engine(x224, batch=b224)
    eng2.set_input(x224)
    cam2 = torch.empty((b224, 32, 224, 224), dtype=torch.float32, device=dev)
    for _ in range(3):
        eng2.gradcam(None, (224, 224), True, cam=cam2)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        eng2.gradcam(None, (224, 224), True, cam=cam2)
    e1.record()
    torch.cuda.synchronize()
    out["geometry_224"] = {"value": world * b224 / (e0.elapsed_time(e1) * 1e-3 / reps), "unit": "clips/s",
                           "clips_per_launch_sequence": b224,
                           "note": "32x224x224 clips, device resident, throughput only (no parity claim)"}
    return out


def np_isfinite_or_nan(a):
    import numpy as np
    return np.all(np.isfinite(a) | np.isnan(a))


def clstm_throughput(dev, rank, world, mode, clips_n=8, steps=10, with_cpu=False):
    """Config C3: ConvLSTM (KTH geometry, 2 layers, 5x5, conv stride 2, 32 hidden units as in the paper)
    temporal-mask search with the reverse perturbation on 32x120x160 clips: clip-iterations/s, iteration
    replayed from a CUDA graph, clips resident in HBM."""
    import torch.distributed as dist
    from interpreting_video_features_b200 import ops, search
    from interpreting_video_features_b200.pt.models import CLSTM_4
    from oracle import synthetic
    torch.manual_seed(0)
    m = quiet(CLSTM_4.Model, num_classes=6, nb_lstm_units=32, channels=3, conv_kernel_size=(5, 5), lstm_layers=2,
              step=32, conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31],
              batch_normalization=True, dropout=0.5, add_softmax=True).to(dev).eval().set_mode(mode)
    x = torch.stack([synthetic.uniform_clip_u8(2000 + rank * clips_n + i, t=32, h=120, w=160)
                     for i in range(clips_n)]).to(dev).float()  # 0..255, as pt/data_loader_kth.py:28 delivers
    eng = m._engine(x, batch=clips_n)
    ms = search.MaskSearch(eng, 0.02, 0.04, 0.2, 100, "reverse", 0.9, use_graph=True)
    ms.set_input(x)
    ms.set_targets((torch.arange(clips_n) % 6).to(dev))
    ms.m.copy_(torch.tensor([-5.] * 8 + [5.] * 16 + [-5.] * 8, device=dev).repeat(clips_n, 1))
    ops.sigmoid(ms.m, ms.sig)
    ms._capture()
    for _ in range(3):
        ms.graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ms.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([sec], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t[0])
    pk, pk_src = peaks()
    val = world * clips_n * steps / sec
    cpu = None
    if with_cpu and rank == 0:  # the oracle port of the reference loop (pt/FindMasksComparison_I3D_KTH.py:250-270), B = 1
        from oracle import clstm_oracle, mask_oracle
        sdc = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
        om = clstm_oracle.Model(sdc, hidden=32, softmax=True, num_layers=2, kernel=5, conv_stride=2,
                                effective_step=(7, 15, 23, 31))
        x1 = x[:1].cpu()
        tm = torch.tensor([-5.] * 8 + [5.] * 16 + [-5.] * 8, requires_grad=True)
        t0 = time.perf_counter()
        mask_oracle.mask_search(x1, om, 0, [0], tm, 0.02, 0.04, 2, mask_type="reverse")
        dt = (time.perf_counter() - t0) / 2
        cpu = {"value": 1.0 / dt, "unit": "clip-iterations/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "2 clip-iterations (B=1, 32x120x160, hidden 32) of the oracle port of the reference loop"}
    return {"metric": "clstm_mask_search_clip_iterations_per_sec", "unit": "clip-iterations/s",
            "value": val, "ms_per_step": sec / steps * 1e3, "clips_per_gpu": clips_n,
            "launches_per_step": int(ms.launches_per_iter),
            "algorithmic_gflop_per_clip_iteration": 76.7,
            "roofline": {"bound": "tensor", "achieved": val / world * 76.7 / 1e3, "unit": "TFLOP/s",
                         "peak": pk["bf16_tflops"], "frac": val / world * 76.7 / 1e3 / pk["bf16_tflops"],
                         "peak_source": pk_src, "note": "whole step (62 sequential recurrent steps per layer and "
                         "direction: launch and latency bound, not tensor bound)"},
            "cpu_baseline": cpu,
            "workload": "C3: ConvLSTM (6 classes, hidden 32, 2 layers) temporal-mask search, reverse perturbation, "
                        "%d synthetic 32x120x160 clips per step" % clips_n}


def train_throughput(dev, rank, world, clips_n=8, steps=3, with_cpu=False):
    """one TRAINING step of the I3D classifier (pt/train_i3d_smth.py:192-250: forward in training
    mode, CrossEntropyLoss, backward to every parameter, SGD update) on synthetic 16x224x224 clips through I3DTrainer,
    mixed precision (tcgen05 forward convolutions and data gradients, CUDA-core weight gradients) and fp32.  clips/s = clips per step / step time; algorithmic work per clip = forward +
    data gradient (all but the stem's) + weight gradient convolutions."""
    import torch.distributed as dist
    from interpreting_video_features_b200.train import I3DTrainer
    from oracle import i3d_oracle, synthetic
    model = state_dict()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.stack([synthetic.uniform_clip(5000 + rank * clips_n + i) for i in range(clips_n)]).to(dev)
    target = (torch.arange(clips_n) * 7 % NCLS).to(dev)
    def timed(mode):
        tr = I3DTrainer(sd, clips_n, (T, H, W), device=dev, optimizer="sgd", lr=1e-3, momentum=0.9, weight_decay=1e-5,
                        dropout_p=0.5, mode=mode, world=world)
        n0 = _launches(dev)
        tr.step(x, target)  # warm-up (lazy kernel loading) and the launch count of one step
        n_launch = _launches(dev) - n0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            last = tr.step(x, target)
        e1.record()
        torch.cuda.synchronize()
        s = e0.elapsed_time(e1) * 1e-3 / steps
        if world > 1:
            t = torch.tensor([s], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s = float(t[0])
            # data parallel: every rank trained on its own clips and holds the same parameters afterwards
            chk = torch.stack([p.double().abs().sum() for p in tr.params.values()]).sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert float(hi - lo) <= 1e-9 * float(hi), "ranks hold different parameters after data-parallel steps"
        return s, n_launch, float(last)

    sec32, _, loss32 = timed("fp32")
    sec, launches, loss = timed("bf16")
    fwd = i3d_oracle.conv_flops_per_clip(sd, (T, H, W))
    stem = 2.0 * 343 * 3 * 64 * (T // 2) * (H // 2) * (W // 2)
    gf_clip = (3.0 * fwd - stem) / 1e9
    val = world * clips_n / sec
    out = {"metric": "i3d_training_clips_per_sec", "unit": "clips/s", "value": val, "ms_per_step": sec * 1e3,
           "clips_per_step_per_gpu": clips_n, "launches_per_step": int(launches), "dtype": "bf16",
           "loss_after_%d_steps" % (steps + 1): float(loss),
           "algorithmic_gflop_per_clip_step": gf_clip, "achieved_tflops": val / world * gf_clip / 1e3,
           "fp32_mode": {"value": world * clips_n / sec32, "ms_per_step": sec32 * 1e3,
                         "loss_after_%d_steps" % (steps + 1): loss32,
                         "achieved_tflops": clips_n / sec32 * gf_clip / 1e3},
           "note": "mixed precision: forward convolutions and data gradients are the tcgen05 implicit GEMMs of the "
                   "interpretation path, the weight gradient is still a CUDA-core fp32-accumulate kernel and is most "
                   "of the step (tools/train_events.py); fp32_mode runs every convolution on CUDA cores",
           "data_parallel": ("%d ranks, one NCCL all-reduce of the flat gradient buffer (%.1f MB) per step, parameters "
                             "verified identical across ranks" % (world, 4e-6 * sum(v.numel() for v in sd.values()
                                                                                    if v.dtype == torch.float32))
                             if world > 1 else None),
           "workload": "f4: I3D smth (174 classes) training step, %d synthetic 16x224x224 clips per GPU, SGD(momentum "
                       "0.9, weight decay 1e-5), dropout 0.5, BatchNorm with batch statistics" % clips_n}
    if with_cpu and rank == 0:
        from oracle import train_oracle
        xc, tc = x[:2].cpu(), target[:2].cpu()
        t0 = time.perf_counter()
        _, _, g, _ = train_oracle.loss_and_grads(sd, xc, tc)
        train_oracle.sgd_step(sd, g, 1e-3, 0.9, 1e-5)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 2.0 / dt, "unit": "clips/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "1 training step of 2 clips (16x224x224) of the oracle port of the reference's "
                                         "PyTorch-CPU step, no warm-up"}
    return out


def _launches(dev):
    from interpreting_video_features_b200 import _lib
    return _lib.launch_count(dev)


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    from interpreting_video_features_b200 import _lib, search
    from oracle import i3d_oracle, synthetic  # FLOP count + synthetic clips only (not on the timed path)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    model = state_dict().to(dev).eval().set_mode(args.mode)
    sd = model.state_dict()
    conv_flops = i3d_oracle.conv_flops_per_clip({k: v for k, v in sd.items()}, (T, H, W))  # forward, 2*MAC
    clips = torch.stack([synthetic.uniform_clip(rank * CLIPS + i) for i in range(CLIPS)])
    targets = torch.randint(0, NCLS, (CLIPS,), generator=torch.Generator().manual_seed(100 + rank))

    groups = search.default_groups(CLIPS)
    engs = search.make_engines(model, clips, CLIPS, groups)
    ms = search.MaskSearch(engs, 0.01, 0.02, 0.2, N_ITER, "freeze", 0.9, use_graph=True)
    xd = clips.to(dev)
    ms.set_input(xd)
    ms.set_targets(targets.to(dev))
    raw = torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4, device=dev).repeat(CLIPS, 1)
    ms.m.copy_(raw)
    from interpreting_video_features_b200 import ops
    ops.sigmoid(ms.m, ms.sig)
    ms._capture()
    launches_per_step = ms.launches_per_iter
    for _ in range(max(args.warmup - 1, 0)):
        ms.graph.replay()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up before the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.t_begin = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        ms.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sampler.t_end = time.perf_counter()
    elapsed = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([elapsed], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.barrier()
        elapsed = float(t.item())
    sampler.stop()
    value = world * CLIPS * args.steps / elapsed

    # ---- roofline of the convolution kernel (rank 0, eager iteration with per-launch events)
    conv_s, conv_launches = conv_time_per_step(engs)
    pk, pk_src = peaks()
    flops_step = 2.0 * conv_flops * CLIPS  # forward + data gradient, no weight gradient
    achieved = flops_step / conv_s / 1e12
    # the step runs at the burst clock (1 965 MHz, no power cap on this workload: see "clocks"), so the burst cuBLAS
    # figure is the denominator; the sustained figure is reported beside it
    peak = pk["bf16_tflops"] if args.mode == "bf16" else None
    peak_sus = pk.get("bf16_tflops_sustained") if args.mode == "bf16" else None
    traffic, traffic_src = None, None
    tp = os.path.join(REPO, "profiles", "r02_step_metrics.json")
    if os.path.exists(tp) and args.mode == "bf16":  # DRAM bytes of the conv launches of one step (ncu, committed)
        traffic = json.load(open(tp))["families"]["conv"]["dram_bytes"]
        traffic_src = "profiles/r02_step_metrics.json: dram__bytes_read.sum + dram__bytes_write.sum summed over the " \
                      "conv launches of one step (8 clips), one ncu capture"
    roofline = {"bound": "tensor", "kernel": "conv_slab_kernel + conv_tc_kernel (tcgen05 implicit GEMMs, all conv launches of a step)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if peak else None, "peak_source": pk_src + " (burst bf16 GEMM, MEASURED_PEAKS.json bf16_tflops)",
                "frac_burst": (achieved / peak) if peak else None, "peak_sustained": peak_sus,
                "frac_sustained": (achieved / peak_sus) if peak_sus else None,
                "traffic": traffic, "traffic_source": traffic_src, "conv_launches_per_step": conv_launches, "conv_ms_per_step": conv_s * 1e3,
                "algorithmic_gflop_per_clip_iteration": 2.0 * conv_flops / 1e9,
                "conv_share_of_step": conv_s / (elapsed / args.steps)}

    # ---- end to end: the sharded C4 job through the public API, uint8 host clips, gather + D2H inside
    per_gpu = args.clips_per_gpu
    n_total = per_gpu * world
    mine = search.shard_indices(n_total, rank, world)
    host = torch.empty((len(mine), 3, T, H, W), dtype=torch.uint8, pin_memory=True)
    for i, gidx in enumerate(mine):
        host[i] = synthetic.uniform_clip_u8(gidx)
    tg_all = torch.randint(0, NCLS, (n_total,), generator=torch.Generator().manual_seed(7))
    tg_mine = tg_all[mine]
    # one short untimed call first: CUDA loads kernels lazily, and the init-mask / reverse-score / Grad-CAM kernels
    # have not run yet in this process; it also captures the graphs the timed call replays
    search.find_masks_batched(model, host[:CLIPS], tg_mine[:CLIPS], n_iter=2, micro_batch=CLIPS, device=dev, gradcam=True,
                              rank=rank, world=world, n_total=CLIPS * world)
    check = None
    if world > 1:  # the gathered result == each shard computed alone, bit for bit (small case: 8 clips per rank)
        small = search.find_masks_batched(model, host[:CLIPS], tg_mine[:CLIPS], n_iter=5, micro_batch=CLIPS, device=dev,
                                          gradcam=True, rank=rank, world=world, n_total=CLIPS * world)
        alone = search.find_masks_batched(model, host[:CLIPS], tg_mine[:CLIPS], n_iter=5, micro_batch=CLIPS,
                                          device=dev, gradcam=True)
        sel = search.shard_indices(CLIPS * world, rank, world)
        ok = all(torch.equal(small[k][sel], alone[k]) for k in ("time_mask", "freeze_score", "reverse_score",
                                                                 "cam_lowres"))
        t = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        check = bool(t.item() == 1.0)
        assert check, "gathered shard rows differ from the single-rank result"
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    stats = {}
    t0 = time.perf_counter()
    mb = max(CLIPS, min(args.e2e_micro_batch, per_gpu))
    if mb != CLIPS:  # untimed: engine of the other micro-batch size (tile plans, graph capture)
        search.find_masks_batched(model, host[:mb], tg_mine[:mb], n_iter=2, micro_batch=mb, device=dev, gradcam=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
    res = search.find_masks_batched(model, host, tg_mine, lam1=0.01, lam2=0.02, n_iter=N_ITER, perturb="freeze",
                                    micro_batch=mb, device=dev, gradcam=True, rank=rank, world=world,
                                    n_total=n_total, stats=stats)
    out_host = {k: res[k].cpu() for k in ("time_mask", "freeze_score", "reverse_score", "cam_lowres")}
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, stats["gather_seconds"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, stats["gather_seconds"] = float(t[0]), float(t[1])
    assert out_host["time_mask"].shape == (n_total, T) and bool(torch.isfinite(out_host["time_mask"]).all())
    d2h = sum(v.numel() * 4 for v in out_host.values())
    steps_per_rank = max(len(mine) // CLIPS, 1) * N_ITER  # 8-clip iterations per rank (the `value` leg's step)
    e2e = {"value": n_total * N_ITER / e2e_s, "unit": UNIT,
           "h2d_bytes_per_step": int(host.numel() / steps_per_rank), "d2h_bytes_per_step": int(d2h / steps_per_rank),
           "h2d_bytes_per_job_per_rank": int(host.numel()), "d2h_bytes_per_job": int(d2h), "seconds_per_job": e2e_s,
           "clips_total": n_total, "clips_per_gpu": per_gpu, "micro_batch": mb, "gather_seconds": stats["gather_seconds"],
           "gathered_bytes": stats.get("gathered_bytes"), "gather_equals_single_rank": check,
           "note": "C4-shaped job: %d uint8 clips per GPU in pinned host memory, find_masks_batched(gradcam=True): per "
                   "micro-batch of %d clips H2D + init_mask (T/2+1 forwards) + 300 iterations + reverse score + Grad-CAM; "
                   "then NCCL all_gather of masks+scores+low-res CAMs (N>1) and D2H of the gathered result; "
                   "h2d/d2h_bytes_per_step are per 8-clip iteration (the `value` leg's step); one 2-iteration call runs "
                   "untimed first (lazy kernel loading, graph capture)" % (per_gpu, mb)}

    # ---- the same conv roofline at the end-to-end leg's micro-batch (the late stages fill the machine there)
    if mb != CLIPS and args.mode == "bf16":
        engs_mb = search.make_engines(model, host[:mb], mb, 1)
        conv_s_mb, conv_n_mb = conv_time_per_step(engs_mb)
        ach_mb = 2.0 * conv_flops * mb / conv_s_mb / 1e12
        roofline["at_e2e_micro_batch"] = {
            "clips_per_step": mb, "conv_ms_per_step": conv_s_mb * 1e3, "conv_launches_per_step": conv_n_mb,
            "achieved": ach_mb, "unit": "TFLOP/s", "frac_burst": ach_mb / peak, "frac_sustained": ach_mb / peak_sus,
            "note": "same measurement (CUDA events around every conv launch of one eager, serialised iteration) on the "
                    "engine the end-to-end job runs: %d clips per launch sequence" % mb}
    gradcam = gradcam_throughput(dev, rank, world, args.mode, with_cpu=(world == 1 and not args.no_cpu)) \
        if not args.no_gradcam else None
    clstm = clstm_throughput(dev, rank, world, args.mode, with_cpu=(world == 1 and not args.no_cpu)) \
        if not args.no_clstm else None
    train = train_throughput(dev, rank, world, with_cpu=(world == 1 and not args.no_cpu)) if not args.no_train else None

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        v, per, cores = time_cpu_reference(2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2 timed clip-iterations (B=1, 16x224x224) of the oracle port of the reference's "
                         "PyTorch-CPU path after 1 warm-up, %d threads" % cores}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f32",
            "data": "synthetic", "config": CONFIG, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gradcam": gradcam, "clstm": clstm, "train": train,
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--clips-per-gpu", type=int, default=128, help="clips per GPU of the end-to-end (C4) leg")
    ap.add_argument("--e2e-micro-batch", type=int, default=int(os.environ.get("IVF_E2E_MICRO_BATCH", "128")),
                    help="clips per launch sequence in the end-to-end leg (the headline `value` stays at BASELINE's 8)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gradcam", action="store_true", help="skip the Grad-CAM clips/s leg")
    ap.add_argument("--no-clstm", action="store_true", help="skip the ConvLSTM (config C3) leg")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step (f4) leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
