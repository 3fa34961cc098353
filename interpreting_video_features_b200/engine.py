"""Native I3D runner: the forward and data-gradient passes of pt/models/I3D_doubled.py (and the
_kth variant) as a fixed sequence of libivf kernel launches over preallocated channels-last
buffers, so one mask-search iteration (pt/FindMasksComparison_I3D_smth.py:193-214) can be
captured in a CUDA graph.

What differs from the reference's execution (not from its results):
  * eval BatchNorm + ReLU live in the conv epilogue; 'same' padding is TMA zero fill; the
    Inception concat is a channel offset (pt/models/I3D_doubled.py:96-118,146);
  * only the data gradient is computed (the reference also computes 12.47 M weight gradients per
    iteration that nobody reads, pt/FindMasksComparison_I3D_smth.py:191,212-214);
  * every clip of the batch carries its own mask (the reference forwards B clips under ONE mask and
    reads one of them, :202-205);
  * bf16 mode presents the stride-2 7x7x7 stem space-to-depth: a stride-1 4x4x4 convolution over
    8*3 channels (forward) and its flipped transpose (data gradient), so every convolution of the
    network is the same stride-1 tcgen05 implicit GEMM.

  * the four branches of an Inception block (pt/models/I3D_doubled.py:121-146) are independent until
    the concat, so the program forks them onto side streams (joined before the next stage): the
    14x14 / 7x7 stages launch 49- and 7-CTA grids that cannot fill 148 SMs one at a time.  The fork
    and join are captured into the CUDA graph as parallel branches.

State-dict keys are the reference's; weights are packed once at construction.
"""
import os

import torch

from . import _lib, ops, tune
from ._lib import PFMT_NDHWC_F32, PFMT_S2D_BF16
from .ops import Act, same_pad

ENDPOINTS = ("Conv3d_1a_7x7", "MaxPool3d_2a_3x3", "Conv3d_2b_1x1", "Conv3d_2c_3x3", "MaxPool3d_3a_3x3",
             "Mixed_3b", "Mixed_3c", "MaxPool3d_4a_3x3", "Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e",
             "Mixed_4f", "MaxPool3d_5a_2x2", "Mixed_5b", "Mixed_5c")
POOLS = {"MaxPool3d_2a_3x3": ((1, 3, 3), (1, 2, 2)), "MaxPool3d_3a_3x3": ((1, 3, 3), (1, 2, 2)),
         "MaxPool3d_4a_3x3": ((3, 3, 3), (2, 2, 2)), "MaxPool3d_5a_2x2": ((2, 2, 2), (2, 2, 2))}


def strip_module_prefix(sd):
    """DataParallel checkpoints carry 'module.' (pt/utils.py:94-104)."""
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


# ---------------------------------------------------------------------------- weight packing
def _dev32(t, device):
    """A parameter as a contiguous fp32 device tensor (one H2D/D2D copy, no compute)."""
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def pack(sources, mode, dgrad=False, s2d=(1, 1, 1), ci_stride=0, co_offs=None, co_total=None, ceff_total=None,
         co_stride=1, out=None):
    """Operand matrix of a convolution kernel from one or more fp32 OIDHW (OIHW: kd = 1) device weights that
    share their input channels and kernel (ivf_pack_weights; nothing is computed by torch).

    forward operand: rows = output channels (sources stacked at co_offs), K = operand channels;
    data-gradient operand (dgrad): rows = operand channels, K = output channels (sources side by side at co_offs).
    s2d: per-axis space-to-depth factor of a stride-2 layer; ci_stride pads each parity block of operand
    channels; ceff_total pads the operand channels as a whole (the 12 -> 16 channel ConvLSTM input record).
    co_stride: output channel j of a source lands at co_off + j * co_stride (4 with co_offs 0..3 interleaves the four
    ConvLSTM gates unit-major for the fused recurrent step).
    bf16: K-major [n_pad][taps][k_pad]; fp32: tap-major [taps][k][n] (flat [taps*k, n] like the old packers).
    out: an operand matrix a previous call with the same arguments returned - refilled in place (its padding is
    already zero: no clearing pass, no allocation; the training step re-packs every weight after every update)."""
    lib = _lib.load()
    ws = [w if w.dim() == 5 else w.unsqueeze(2) for w in sources]
    dev = ws[0].device
    ci = ws[0].shape[1]
    kd, kh, kw = ws[0].shape[2:]
    for w in ws:
        assert w.dtype == torch.float32 and w.is_contiguous() and tuple(w.shape[1:]) == (ci, kd, kh, kw)
    taps = ((kd + s2d[0] - 1) // s2d[0]) * ((kh + s2d[1] - 1) // s2d[1]) * ((kw + s2d[2] - 1) // s2d[2])
    ceff = s2d[0] * s2d[1] * s2d[2] * max(ci_stride, ci)
    ceff_total = ceff if ceff_total is None else ceff_total
    assert ceff_total >= ceff
    if co_offs is None:
        co_offs, acc = [], 0
        for w in ws:
            co_offs.append(acc)
            acc += w.shape[0]
        co_total = acc if co_total is None else co_total
    assert co_total >= max(o + (w.shape[0] - 1) * co_stride + 1 for o, w in zip(co_offs, ws))
    n_total, k_total = (ceff_total, co_total) if dgrad else (co_total, ceff_total)
    bf = mode == "bf16"
    if bf:
        n_pad, k_pad = lib.ivf_conv_bf16_cout_pad(n_total), lib.ivf_conv_bf16_cin_pad(k_total)
        shape, dt = (n_pad, taps, k_pad), torch.bfloat16
    else:
        n_pad, k_pad = n_total, k_total
        shape, dt = (taps * k_pad, n_pad), torch.float32
    if out is not None:
        assert tuple(out.shape) == shape and out.dtype == dt and out.device == dev
        dst = out
    else:
        dst = torch.empty(shape, dtype=dt, device=dev)
    h = _lib.handle(dev)
    for i, (w, off) in enumerate(zip(ws, co_offs)):
        d = _lib.PackDesc()
        d.co, d.ci, d.kd, d.kh, d.kw = w.shape[0], ci, kd, kh, kw
        d.ci_stride = ci_stride
        d.s2d_d, d.s2d_h, d.s2d_w = s2d
        d.dgrad = (1 if bf else 2) if dgrad else 0
        d.layout = _lib.PACK_KMAJOR if bf else _lib.PACK_TAPMAJOR
        d.dtype = _lib.IVF_BF16 if bf else _lib.IVF_F32
        d.n_pad, d.k_pad = n_pad, k_pad
        d.n_off, d.k_off = (0, off) if dgrad else (off, 0)
        d.n_stride, d.k_stride = (1, co_stride) if dgrad else (co_stride, 1)
        d.zero_first = int(i == 0 and out is None)
        _lib.check(lib.ivf_pack_weights(h, d, _lib.ptr(w), _lib.ptr(dst), _lib.stream_ptr(dev)), "ivf_pack_weights")
    return dst


def pack_fwd(w, mode):
    return pack([w], mode)


def pack_dgrad(w, mode):
    """fp32: [taps][co][ci] for the transposed gather; bf16: flipped taps, roles swapped."""
    return pack([w], mode, dgrad=True)


def pack_dgrad_two_sources(w_first, w_second):
    """Data-gradient weights of 1x1x1 units that share their input, reduced in ONE GEMM over
    [dz of the first unit | dz of the second]: K-major [ci][K], the first unit's K padded to whole 64-channel
    stages (ivf_conv3d_split), zeros in the padding."""
    k1 = (w_first.shape[0] + 63) // 64 * 64
    return pack([w_first, w_second], "bf16", dgrad=True, co_offs=[0, k1], co_total=k1 + w_second.shape[0])


def bn_fold(sd, prefix, device, eps, scale, shift, with_conv_bias=False):
    """eval BatchNorm of `prefix` folded into scale/shift (views of the caller's buffers); units without a
    BatchNorm get scale 1 / shift 0; with_conv_bias adds scale * conv bias to the shift."""
    lib = _lib.load()
    c = scale.numel()
    bias = sd.get(prefix + ".conv3d.bias") if with_conv_bias else None
    bias = _dev32(bias, device) if bias is not None else None
    if prefix + ".bn.weight" in sd:
        g, b, mu, var = (_dev32(sd["%s.bn.%s" % (prefix, k)], device)
                         for k in ("weight", "bias", "running_mean", "running_var"))
    else:
        g = b = mu = var = None
    _lib.check(lib.ivf_bn_fold(_lib.handle(device), _lib.ptr(g), _lib.ptr(b), _lib.ptr(mu), _lib.ptr(var), eps,
                               _lib.ptr(bias), c, _lib.ptr(scale), _lib.ptr(shift), _lib.stream_ptr(device)),
               "ivf_bn_fold")
    return [g, b, mu, var, bias]  # kept alive by the caller until the stream has consumed them


class SplitConvOp:
    """A 1x1x1 convolution with two destinations or two sources (ops.conv1x1_split) as a program item."""

    def __init__(self, x, w, out, **kw):
        self.x, self.w, self.out, self.kw = x, w, out, kw

    def __call__(self):
        ops.conv1x1_split(self.x, self.w, self.out, **self.kw)


class ConvOp:
    """One convolution launch of a program, kept as data so that its tile plan can be measured (tune.py)."""

    def __init__(self, x, w, out, kernel, stride, pf, **kw):
        self.x, self.w, self.out, self.kernel, self.stride, self.pf, self.kw = x, w, out, kernel, stride, pf, kw
        self.plan = None

    def desc(self, plan=None):
        kw = self.kw
        return ops.conv_desc(self.x, self.out, self.kernel, self.stride, self.pf, kw.get("flags", 0), kw.get("scale"),
                             kw.get("acc_in"), kw.get("mask"), kw.get("transposed", 0), plan=plan)

    def launch(self, plan):
        ops.conv3d(self.x, self.w, self.out, self.kernel, self.stride, self.pf, plan=plan, **self.kw)

    def __call__(self):
        self.launch(self.plan)

    def tune(self, device, min_ms=None):
        self.plan = tune.best_plan(self.desc, self.launch, device, min_ms=min_ms)


class PairConvOp:
    """Two independent ConvOps issued through ivf_conv3d_pair: one grouped launch of the halo-slab kernel where it can
    (the two 3x3x3 branches of an Inception module), otherwise two launches on the same lane."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    @staticmethod
    def _args(op):
        return dict(x=op.x, w=op.w, out=op.out, kernel=op.kernel, stride=op.stride, pad_front=op.pf, **op.kw)

    def __call__(self):
        ops.conv3d_pair(self._args(self.a), self._args(self.b))


class Unit:
    """Unit3D (pt/models/I3D_doubled.py:43-118): conv weights packed both ways + folded BN.  `prefix` may be a
    list: 1x1x1 units that read the same input, fused along the output channels into one GEMM."""

    def __init__(self, sd, prefix, stride, mode, device, s2d=False):
        prefixes = [prefix] if isinstance(prefix, str) else list(prefix)
        ws = [_dev32(sd[p + ".conv3d.weight"], device) for p in prefixes]
        self.cout, self.cin = sum(w.shape[0] for w in ws), ws[0].shape[1]
        self.kernel = tuple(ws[0].shape[2:])
        self.stride = tuple(stride)
        self.s2d = s2d
        self.scale = torch.empty(self.cout, dtype=torch.float32, device=device)
        self.shift = torch.empty(self.cout, dtype=torch.float32, device=device)
        keep, off = [], 0
        for p, w in zip(prefixes, ws):  # BatchNorm3d(eps=0.001), pt/models/I3D_doubled.py:75
            c = w.shape[0]
            keep.append(bn_fold(sd, p, device, 1e-3, self.scale[off:off + c], self.shift[off:off + c]))
            off += c
        if len(prefixes) == 1 and prefixes[0] + ".conv3d.bias" in sd:  # the logits unit (no BN, bias)
            self.shift_with_bias = torch.empty_like(self.shift)
            keep.append(bn_fold(sd, prefixes[0], device, 1e-3, torch.empty_like(self.scale), self.shift_with_bias,
                                with_conv_bias=True))
        else:
            self.shift_with_bias = self.shift
        if s2d:
            assert mode == "bf16" and self.stride == (2, 2, 2)
            win = (self.kernel[0] + 1) // 2
            self.kernel_eff, self.stride_eff, self.cin_eff = (win,) * 3, (1, 1, 1), 8 * self.cin
            f = (2, 2, 2)
        else:
            self.kernel_eff, self.stride_eff, self.cin_eff = self.kernel, self.stride, self.cin
            f = (1, 1, 1)
        self.w_fwd = pack(ws, mode, s2d=f)
        self.w_dgrad = pack(ws, mode, dgrad=True, s2d=f)
        self._src = (ws, keep)  # the fp32 sources stay referenced: the pack kernels read them asynchronously


class I3DEngine:
    def __init__(self, state_dict, batch, clip, mode="bf16", softmax=True, avg_pool=(2, 7, 7),
                 stride_mods=None, device=None, in_channels=3):
        assert mode in ("bf16", "fp32")
        self.mode = mode
        self.dtype = torch.bfloat16 if mode == "bf16" else torch.float32
        self.device = torch.device(device if device is not None else "cuda")
        _lib.handle(self.device)  # fails loudly without the extension / a GPU
        sd = strip_module_prefix(state_dict)
        self.B, (self.T, self.H, self.W) = batch, clip
        self.C = in_channels
        self.softmax = bool(softmax)
        self.generation = 0  # bumped by everything that overwrites activations or dprobs
        stride_mods = stride_mods or {}
        dev = self.device
        B = batch

        def strides_of(name, default):
            return stride_mods.get(name, default)

        # ---- stem input format
        s1 = strides_of("Conv3d_1a_7x7", (2, 2, 2))
        if mode == "bf16":
            if s1 != (2, 2, 2) or self.T % 2 or self.H % 2 or self.W % 2 or in_channels > 4:
                raise _lib.IvfError("bf16 mode needs the stride-(2,2,2) stem on even T/H/W; use mode='fp32'")
            self.in_fmt = PFMT_S2D_BF16
            self.xin = Act.empty(B, self.T // 2, self.H // 2, self.W // 2, 32, torch.bfloat16, dev, zero=True)
            self.g_xin = Act.empty(B, self.T // 2, self.H // 2, self.W // 2, 32, torch.bfloat16, dev, zero=True)
        else:
            self.in_fmt = PFMT_NDHWC_F32
            self.xin = Act.empty(B, self.T, self.H, self.W, in_channels, torch.float32, dev)
            self.g_xin = Act.empty(B, self.T, self.H, self.W, in_channels, torch.float32, dev)

        # programs: (lane, fn) launches plus ("fork",)/("join",) markers; lane 0 is the caller's stream,
        # lanes 1-3 are side streams used between a fork and its join
        self.fwd_ops, self.bwd_ops = [], []
        self._lane = 0
        self.use_streams = os.environ.get("IVF_STREAMS", "1") != "0"
        self.pool_premask = os.environ.get("IVF_POOL_PREMASK", "1") != "0"
        # the two 3x3x3 branches of an Inception module as one grouped launch (bf16: ivf_conv3d_pair), opt-in:
        # measured in situ the grouped launch saves 7-13 us per pair on the 14x14 / 7x7 stages at 8 clips (the summed
        # convolution time drops 2.08 -> 2.02 ms), but both branches then sit on ONE lane and the step loses the
        # overlap of lanes 1 and 2: 2.059 ms without, 2.134 ms with (IVF_PAIR_MAX_PIXELS=1000: 2.097 ms)
        self.pair_branches = mode == "bf16" and os.environ.get("IVF_PAIR_BRANCHES", "0") != "0"
        self.pair_max_pixels = int(os.environ.get("IVF_PAIR_MAX_PIXELS", "8192"))
        self._side = None
        self.acts = {}  # endpoint -> Act (forward output)
        # each stage: dict(out=Act, scale=tensor|None (None: pool-type output), gout=Act)
        stages = []

        def new_act(n, d, h, w, c, dtype=None, zero=False):
            return Act.empty(n, d, h, w, c, dtype or self.dtype, dev, zero)

        # ------------------------------------------------------------------ builders
        def add_unit_fwd(unit, x, out, xin_is_s2d=False):
            if xin_is_s2d:
                pf = (1, 1, 1)
                xa = Act(x.buf, x.n, x.d, x.h, x.w, x.ld, 0, unit.cin_eff)
            else:
                pf = tuple(same_pad(sz, k, s)[0] for sz, k, s in zip((x.d, x.h, x.w), unit.kernel, unit.stride))
                xa = x
            self.fwd_ops.append((self._lane, ConvOp(xa, unit.w_fwd, out, unit.kernel_eff, unit.stride_eff, pf,
                                                    flags=_lib.EP_RELU, scale=unit.scale, shift=unit.shift)))

        def add_unit_bwd(unit, dz_out, x_shape_act, g_in, acc_in=None, mask=None, mask_scale=None,
                         xin_is_s2d=False):
            """g_in (+)= dgrad(dz_out); g_in is an Act shaped like the unit's input."""
            if self.mode == "fp32":
                pf = tuple(same_pad(sz, k, s)[0] for sz, k, s in
                           zip((x_shape_act.d, x_shape_act.h, x_shape_act.w), unit.kernel, unit.stride))
                self.bwd_ops.append((self._lane, ConvOp(dz_out, unit.w_dgrad, g_in, unit.kernel, unit.stride, pf,
                                                        acc_in=acc_in, mask=mask, mask_scale=mask_scale,
                                                        transposed=1)))
            else:
                if xin_is_s2d:
                    pf = tuple(k - 1 - 1 for k in unit.kernel_eff)
                    gi = Act(g_in.buf, g_in.n, g_in.d, g_in.h, g_in.w, g_in.ld, 0, unit.cin_eff)
                else:
                    assert unit.stride == (1, 1, 1)
                    pf = tuple(k - 1 - same_pad(sz, k, 1)[0] for sz, k in
                               zip((x_shape_act.d, x_shape_act.h, x_shape_act.w), unit.kernel))
                    gi = g_in
                self.bwd_ops.append((self._lane, ConvOp(dz_out, unit.w_dgrad, gi, unit.kernel_eff, (1, 1, 1), pf,
                                                        acc_in=acc_in, mask=mask, mask_scale=mask_scale)))

        def out_dims(x, k, s):
            return tuple(same_pad(sz, kk, ss)[2] for sz, kk, ss in zip((x.d, x.h, x.w), k, s))

        # ------------------------------------------------------------------ build the network
        self.units = {}
        x = self.xin
        prev = None  # previous stage record
        for name in ENDPOINTS:
            if name.startswith("Conv3d"):
                stride = strides_of(name, (2, 2, 2)) if name == "Conv3d_1a_7x7" else (1, 1, 1)
                first = name == "Conv3d_1a_7x7"
                unit = Unit(sd, name, stride, mode, dev, s2d=(first and mode == "bf16"))
                self.units[name] = unit
                if first and mode == "bf16":
                    od, oh, ow = x.d, x.h, x.w
                else:
                    od, oh, ow = out_dims(x, unit.kernel, unit.stride)
                out = new_act(B, od, oh, ow, unit.cout)
                add_unit_fwd(unit, x, out, xin_is_s2d=(first and mode == "bf16"))
                stages.append(dict(kind="unit", name=name, unit=unit, x=x, out=out, scale=unit.scale,
                                   gout=out.like(), first=first))
            elif name.startswith("MaxPool"):
                k, s0 = POOLS[name]
                s = strides_of(name, s0)
                pads = tuple(same_pad(sz, kk, ss)[0] for sz, kk, ss in zip((x.d, x.h, x.w), k, s))
                od, oh, ow = out_dims(x, k, s)
                out = new_act(B, od, oh, ow, x.c)
                am = torch.empty((out.pixels, x.c), dtype=torch.uint8, device=dev)
                # ReLU' bit mask of the pool's input, written by the forward kernel: the backward then reads one
                # byte per 8 channels instead of the producer's bf16 output (bf16 stride-2 pools only)
                bits = None
                if (mode == "bf16" and x.c % 8 == 0 and min(s[1:]) >= 2 and stages and stages[-1]["scale"] is not None
                        and os.environ.get("IVF_POOL_BITS", "0") != "0"):
                    bits = torch.empty((x.pixels, x.c // 8), dtype=torch.uint8, device=dev)
                # every pool of the network reads ReLU outputs (units, Inception concats, pools of those)
                self.fwd_ops.append((0, lambda x=x, out=out, am=am, k=k, s=s, pads=pads, bits=bits:
                                     ops.maxpool3d_fwd(x, out, am, k, s, pads, relu_bits=bits, nonneg=True)))
                stages.append(dict(kind="pool", name=name, x=x, out=out, scale=None, gout=out.like(), argmax=am,
                                   k=k, s=s, pads=pads, bits=bits))
            else:  # Inception module
                rec = self._build_inception(sd, name, x, new_act, add_unit_fwd)
                stages.append(rec)
            x = stages[-1]["out"]
            self.acts[name] = x
        self.stages = stages

        # ---- head (pt/models/I3D_doubled.py:313-333,360-371)
        feat = x
        if (feat.d, feat.h, feat.w) != tuple(avg_pool):
            raise _lib.IvfError(
                "Mixed_5c map %s is not covered by avg_pool %s: the reference's squeeze() would not give "
                "[B, classes] here (SURVEY fact 10)" % ((feat.d, feat.h, feat.w), tuple(avg_pool)))
        wl = sd["logits.conv3d.weight"].detach().to(device=dev, dtype=torch.float32)
        self.num_classes = wl.shape[0]
        self.w_logits = wl.reshape(self.num_classes, -1).contiguous()
        self.b_logits = sd["logits.conv3d.bias"].detach().to(device=dev, dtype=torch.float32).contiguous()
        self.logits = ops.zeros((B, self.num_classes), torch.float32, dev)
        self.probs = ops.zeros((B, self.num_classes), torch.float32, dev)
        self.dprobs = ops.zeros((B, self.num_classes), torch.float32, dev)
        self.head_ws = ops.head_workspace(B, feat.c, self.num_classes, dev)  # this engine's partial logits
        self.fwd_ops.append((0, lambda: ops.head_fwd(feat, self.w_logits, self.b_logits, self.softmax, self.probs,
                                                     self.logits, workspace=self.head_ws)))
        # Grad-CAM reads the raw (unmasked) gradient w.r.t. Mixed_5c in fp32
        self.g_feat_raw = feat.like(torch.float32)

        # ---- backward program
        self._add_unit_bwd = add_unit_bwd
        self._emit_head_bwd()
        for i in range(len(stages) - 1, -1, -1):
            self._emit_stage_bwd(i)

        # ---- mask-search state; x is a static buffer so a captured graph stays valid across batches
        self.x = ops.zeros((B, in_channels, self.T, self.H, self.W), torch.float32, dev)
        self.dm = ops.zeros((B, self.T), torch.float32, dev)
        self.zero_mask = ops.zeros((B, self.T), torch.float32, dev)
        if mode == "bf16" and tune.enabled():  # measured tile plans for the slab convolutions (cached per shape)
            with torch.cuda.device(dev):
                for item in self.fwd_ops + self.bwd_ops:
                    if isinstance(item[0], int) and isinstance(item[1], ConvOp):
                        item[1].tune(dev)
                torch.cuda.synchronize(dev)

    # -------------------------------------------------------------------------------------
    def _emit_head_bwd(self):
        last = self.stages[-1]
        self.bwd_ops.append((0, lambda: ops.head_bwd(last["gout"], self.w_logits, self.softmax, self.probs,
                                                     self.dprobs, mask=last["out"], mask_scale=last["scale"])))

    def _emit_stage_bwd(self, i, raw_out=None, premask=True):
        """Append the data-gradient launches of stage i to self.bwd_ops.  Normally its input gradient goes to the
        previous stage's `gout` with that stage's ReLU'/BN' applied; raw_out (an fp32 Act shaped like the previous
        stage's output) receives the UNMASKED gradient w.r.t. that output instead - Grad-CAM's `grads_val`.

        Max-pool after a ReLU unit (the four stage pools): the pool routes each window's gradient to its arg-max
        element only, and that element is positive exactly when the POOLED value is (inputs are ReLU outputs, the
        padding is zero).  So the producer's ReLU'/BN' is applied one stage earlier, by the pool's CONSUMER, with the
        pooled output as the mask (an eighth / a quarter of the pool's input), and the pool's own backward is pure
        routing: it no longer reads the producer's output.  premask=False keeps the unmasked `gout` (needed when the
        pool's backward feeds a raw Grad-CAM gradient)."""
        stages, mode = self.stages, self.mode
        st = stages[i]
        prv = stages[i - 1] if i > 0 else None

        def pre(j):  # stage j is a pool whose consumer applies the ReLU'/BN' of the pool's producer
            return (self.pool_premask and j >= 1 and stages[j]["kind"] == "pool"
                    and stages[j - 1]["scale"] is not None)

        if raw_out is not None:
            g_in, mask, mscale = raw_out, None, None
        elif prv is None:
            g_in, mask, mscale = self.g_xin, None, None
        elif premask and pre(i - 1):
            g_in, mask, mscale = prv["gout"], prv["out"], stages[i - 2]["scale"]
        elif st["kind"] == "pool" and pre(i):
            g_in, mask, mscale = prv["gout"], None, None  # already applied to this pool's gout by its consumer
        else:
            g_in = prv["gout"]
            mask = prv["out"] if prv["scale"] is not None else None
            mscale = prv["scale"]
        if st["kind"] == "unit":
            self._add_unit_bwd(st["unit"], st["gout"], st["x"], g_in, mask=mask, mask_scale=mscale,
                               xin_is_s2d=(st["first"] and mode == "bf16"))
        elif st["kind"] == "pool":
            self.bwd_ops.append((0, lambda st=st, g_in=g_in, mask=mask, mscale=mscale:
                                 ops.maxpool3d_bwd(st["gout"], st["argmax"], g_in, st["k"], st["s"], st["pads"],
                                                   mask=mask, mask_scale=mscale,
                                                   relu_bits=st["bits"] if mask is not None else None)))
        else:
            self._build_inception_bwd(st, g_in, mask, mscale, self._add_unit_bwd)

    def raw_gradient_program(self, layer):
        """(program, fp32 Act): the backward launches from the head down to the consumer of endpoint `layer`, the
        last of them writing the unmasked gradient of sum(dprobs * probs) w.r.t. that endpoint's output - what the
        reference's hook on the layer's output records (pt/pytorch-grad-cam/grad-cam.py:23-54).  Built on first use
        per layer; the activations and the masked `gout` buffers of the later stages are the engine's own."""
        cache = self.__dict__.setdefault("_raw_progs", {})
        if layer in cache:
            return cache[layer]
        names = [st["name"] for st in self.stages]
        if layer not in names:
            raise _lib.IvfError("unknown endpoint %r (have %s)" % (layer, names))
        idx = names.index(layer)
        if idx == len(names) - 1:
            raise _lib.IvfError("the last endpoint's raw gradient comes from head_grad_raw()")
        raw = self.stages[idx]["out"].like(torch.float32)
        saved, self.bwd_ops, self._lane = self.bwd_ops, [], 0
        try:
            self._emit_head_bwd()
            for i in range(len(self.stages) - 1, idx, -1):
                # the raw gradient behind a stage pool needs that pool's gout WITHOUT the producer's ReLU'/BN'
                self._emit_stage_bwd(i, raw_out=raw if i == idx + 1 else None, premask=(i != idx + 2))
            prog = self.bwd_ops
        finally:
            self.bwd_ops = saved
        cache[layer] = (prog, raw)
        return cache[layer]

    def _build_inception(self, sd, name, x, new_act, add_unit_fwd):
        mode, dev = self.mode, self.device
        u = {b: Unit(sd, "%s.%s" % (name, b), (1, 1, 1), mode, dev) for b in ("b0", "b1a", "b1b", "b2a", "b2b", "b3b")}
        for b, unit in u.items():
            self.units["%s.%s" % (name, b)] = unit
        c0, c1, c2, c3, c4, c5 = (u[b].cout for b in ("b0", "b1a", "b1b", "b2a", "b2b", "b3b"))
        cout = c0 + c2 + c4 + c5
        out = new_act(x.n, x.d, x.h, x.w, cout)
        # the 1x1x1 units that read x: b1a and b2a always share ONE GEMM (t12 = [t1 | t2]); on the tensor-core
        # path b0 joins them (its channels go straight into the concat buffer: ivf_conv3d_split), and backward
        # is likewise one data-gradient GEMM over the concatenated K
        fuse_b0 = mode == "bf16" and os.environ.get("IVF_FUSE_B0", "1") != "0"
        group = ("b0", "b1a", "b2a") if fuse_b0 else ("b1a", "b2a")
        fused = Unit(sd, ["%s.%s" % (name, b) for b in group], (1, 1, 1), mode, dev)
        if fuse_b0:
            w0 = _dev32(sd["%s.b0.conv3d.weight" % name], dev)
            k1 = (w0.shape[0] + 63) // 64 * 64
            w1, w2 = (_dev32(sd["%s.%s.conv3d.weight" % (name, b)], dev) for b in ("b1a", "b2a"))
            fused.w_dgrad = pack([w0, w1, w2], "bf16", dgrad=True, co_offs=[0, k1, k1 + w1.shape[0]],
                                 co_total=k1 + w1.shape[0] + w2.shape[0])
            fused._src2 = (w0, w1, w2)
        t12 = new_act(x.n, x.d, x.h, x.w, c1 + c3)
        t1, t2 = t12.slice(0, c1), t12.slice(c1, c3)
        t3 = new_act(x.n, x.d, x.h, x.w, x.c)
        am = torch.empty((x.pixels, x.c), dtype=torch.uint8, device=dev)
        k3 = (3, 3, 3)
        pads = tuple(same_pad(sz, 3, 1)[0] for sz in (x.d, x.h, x.w))
        self.fwd_ops.append(("fork",))
        self._lane = 3
        self.fwd_ops.append((3, lambda: ops.maxpool3d_fwd(x, t3, am, k3, (1, 1, 1), pads, nonneg=True)))
        add_unit_fwd(u["b3b"], t3, out.slice(c0 + c2 + c4, c5))
        self._lane = 0
        if fuse_b0:
            self.fwd_ops.append((0, SplitConvOp(x, fused.w_fwd, out.slice(0, c0), out2=t12, flags=_lib.EP_RELU,
                                                scale=fused.scale, shift=fused.shift)))
        else:
            add_unit_fwd(fused, x, t12)
        if self.pair_branches and x.pixels <= self.pair_max_pixels:  # b1b, b2b as one grouped launch on lane 1
            self.fwd_ops.append(("fork", (1,)))
            self._lane = 1
            add_unit_fwd(u["b1b"], t1, out.slice(c0, c2))
            add_unit_fwd(u["b2b"], t2, out.slice(c0 + c2, c4))
            (_, op_b) = self.fwd_ops.pop()
            (_, op_a) = self.fwd_ops.pop()
            self.fwd_ops.append((1, PairConvOp(op_a, op_b)))
        else:
            self.fwd_ops.append(("fork", (1, 2)))  # the 3x3x3 branches start once their bottlenecks exist
            self._lane = 1
            add_unit_fwd(u["b1b"], t1, out.slice(c0, c2))
            self._lane = 2
            add_unit_fwd(u["b2b"], t2, out.slice(c0 + c2, c4))
        self._lane = 0
        if not fuse_b0:
            add_unit_fwd(u["b0"], x, out.slice(0, c0))
        self.fwd_ops.append(("join",))
        scale = torch.empty(cout, dtype=torch.float32, device=dev)  # BN scales of the concat (ReLU'/BN' masks)
        off = 0
        for b in ("b0", "b1b", "b2b", "b3b"):
            scale[off:off + u[b].cout].copy_(u[b].scale)  # contiguous same-dtype slices: cudaMemcpyAsync
            off += u[b].cout
        g_t12 = t12.like()
        return dict(kind="inception", name=name, units=u, fused=fused, fuse_b0=fuse_b0, x=x, out=out, scale=scale,
                    gout=out.like(),
                    t1=t1, t2=t2, t3=t3, t12=t12, argmax=am, pads=pads, g_t12=g_t12,
                    g_t1=g_t12.slice(0, c1), g_t2=g_t12.slice(c1, c3), g_t3=t3.like(),
                    g_x32=x.like(torch.float32))

    def _build_inception_bwd(self, st, g_in, mask, mscale, add_unit_bwd):
        u, x, dz = st["units"], st["x"], st["gout"]
        c0, c2, c4, c5 = u["b0"].cout, u["b1b"].cout, u["b2b"].cout, u["b3b"].cout
        # Four lanes: the two 3x3x3 branch tails (ReLU'/BN' of b1a/b2a fused, written side by side into
        # g_t12), b0' (starts the fp32 sum over x's consumers) and the pool branch b3b' -> pool' (added in
        # place once b0' is done).  After the join ONE data-gradient GEMM over the concatenated K of the two bottlenecks adds
        # the sum and applies the producer's ReLU'/BN'.  (Summation order per element is fixed: b0, pool, GEMM.)
        acc = st["g_x32"]
        fuse_b0 = st["fuse_b0"]
        self.bwd_ops.append(("fork",))
        self._lane = 1
        add_unit_bwd(u["b1b"], dz.slice(c0, c2), st["t1"], st["g_t1"], mask=st["t1"], mask_scale=u["b1a"].scale)
        paired = self.pair_branches and x.pixels <= self.pair_max_pixels
        self._lane = 1 if paired else 2
        add_unit_bwd(u["b2b"], dz.slice(c0 + c2, c4), st["t2"], st["g_t2"], mask=st["t2"], mask_scale=u["b2a"].scale)
        if paired:  # the two branch tails as one grouped launch
            (_, op_b) = self.bwd_ops.pop()
            (_, op_a) = self.bwd_ops.pop()
            self.bwd_ops.append((1, PairConvOp(op_a, op_b)))
        self._lane = 3
        add_unit_bwd(u["b3b"], dz.slice(c0 + c2 + c4, c5), st["t3"], st["g_t3"])
        if fuse_b0:
            # pool' starts the fp32 sum; b0' is part of the data-gradient GEMM after the join (K = [dz of b0 | g_t12])
            self.bwd_ops.append((3, lambda: ops.maxpool3d_bwd(st["g_t3"], st["argmax"], acc, (3, 3, 3), (1, 1, 1),
                                                              st["pads"])))
            self.bwd_ops.append(("join",))
            self._lane = 0
            self.bwd_ops.append((0, SplitConvOp(dz.slice(0, c0), st["fused"].w_dgrad, g_in, x2=st["g_t12"],
                                                acc_in=acc, mask=mask, mask_scale=mscale)))
            return
        self._lane = 0
        add_unit_bwd(u["b0"], dz.slice(0, c0), x, acc)
        self.bwd_ops.append(("sync", 0, 3))  # pool' adds into the sum b0' started
        self.bwd_ops.append((3, lambda: ops.maxpool3d_bwd(st["g_t3"], st["argmax"], acc, (3, 3, 3), (1, 1, 1),
                                                          st["pads"], acc_in=acc)))
        self.bwd_ops.append(("join",))
        add_unit_bwd(st["fused"], st["g_t12"], x, g_in, acc_in=acc, mask=mask, mask_scale=mscale)

    # ------------------------------------------------------------------------------------- running
    def set_input(self, x):
        """x: [B,3,T,H,W] in the loader's layout, host or device, fp32 (0..255) or uint8 (frames as decoded: they
        cross PCIe as bytes and are converted on the device); copied into the static buffer."""
        assert tuple(x.shape) == (self.B, self.C, self.T, self.H, self.W), (x.shape,)
        if x.dtype == torch.uint8:
            if getattr(self, "x_u8", None) is None:
                self.x_u8 = torch.empty(self.x.shape, dtype=torch.uint8, device=self.device)
            self.x_u8.copy_(x, non_blocking=True)
            ops.u8_to_f32(self.x_u8, self.x)
        else:
            self.x.copy_(x, non_blocking=True)

    def _run(self, prog):
        """Issue a program on the current stream; forked lanes go to side streams (also under CUDA-graph
        capture, where the event waits become graph dependencies)."""
        if not self.use_streams:
            for item in prog:
                if isinstance(item[0], int):
                    item[1]()
            return
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            # IVF_LANE_PRIO="p1,p2,p3": CUDA stream priorities of the three side lanes (lower = more urgent)
            prio = [int(v) for v in os.environ.get("IVF_LANE_PRIO", "0,0,0").split(",")]
            self._side = [torch.cuda.Stream(device=self.device, priority=prio[i]) for i in range(3)]
        for item in prog:
            if item[0] == "fork":  # ("fork",) all side lanes, ("fork", (lanes...)) only those
                ev = torch.cuda.Event()
                ev.record(main)
                for i, sd in enumerate(self._side):
                    if len(item) == 1 or (i + 1) in item[1]:
                        sd.wait_event(ev)
            elif item[0] == "sync":  # ("sync", src lane, dst lane): dst waits for src's work so far
                lanes = [main] + self._side
                ev = torch.cuda.Event()
                ev.record(lanes[item[1]])
                lanes[item[2]].wait_event(ev)
            elif item[0] == "join":
                for sd in self._side:
                    main.wait_stream(sd)
            elif item[0] == 0:
                item[1]()
            else:
                with torch.cuda.stream(self._side[item[0] - 1]):
                    item[1]()

    @_lib.on_device
    def forward(self, mask=None, perturb="freeze"):
        """Perturb (mask: sigmoid-ed values [T] or [B,T]; None = unperturbed clip) and run the network.
        Returns the [B, classes] probability (or logit) buffer — a live buffer, not a copy."""
        self._mask, self._perturb = (self.zero_mask if mask is None else mask), perturb
        self.generation += 1  # the activations of any earlier forward are gone (autograd nodes check this)
        ops.perturb_fwd(self.x, self._mask, perturb, self.in_fmt, self.xin.buf)
        self._run(self.fwd_ops)
        return self.probs

    @_lib.on_device
    def forward_graphed(self):
        """forward(None) - the unperturbed clip in the static input buffer - replayed from a CUDA graph captured
        on first use.  The eager forward is ~60 launches at ~35 us of host time each: launch bound for Grad-CAM,
        which runs one forward per call."""
        if getattr(self, "_fwd_graph", None) is None:
            self.forward(None)  # first use outside the capture: lazy kernel loading, tensor maps
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward(None)
            self._fwd_graph = g
        self.generation += 1
        self._fwd_graph.replay()
        return self.probs

    @_lib.on_device
    def backward(self, to_mask=True):
        """Data-gradient pass from self.dprobs; returns d(sum dprobs*probs)/dmask [B,T] (live buffer)."""
        self._run(self.bwd_ops)
        if to_mask:
            ops.perturb_bwd(self.x, self._mask, self._perturb, self.in_fmt, self.g_xin.buf, self.dm)
        return self.dm

    @_lib.on_device
    def set_targets(self, targets):
        """dprobs = one-hot(targets): the class_loss of pt/FindMasksComparison_I3D_smth.py:205."""
        self.generation += 1
        self._targets = ops.as_int32_targets(targets, self.device)
        ops.one_hot(self._targets, self.dprobs)

    @_lib.on_device
    def gradcam(self, indices=None, out_hw=None, per_frame=True, cam=None, lowres=None, graphed=True,
                layer="Mixed_5c"):
        """Grad-CAM of the clips in the static input buffer (pt/grad_cam_videos.py:64-142): one forward, the head's
        backward only, one fused kernel.  indices None = each clip's arg-max class (:70-71, taken on the device).
        out_hw=(H, W) writes the upsampled, normalised map into `cam` [B, T, H, W] (allocated when None);
        `lowres` [B, T', h, w] receives the map before upsampling (what a multi-GPU job gathers).  Returns
        (cam or None, lowres or None, probs copy)."""
        if layer not in self.acts:
            raise _lib.IvfError("unknown Grad-CAM target layer %r (endpoints: %s)" % (layer, list(self.acts)))
        probs = self.forward_graphed() if graphed else self.forward(None)
        out = probs.clone()
        if indices is None:
            if getattr(self, "_argmax_buf", None) is None:
                self._argmax_buf = torch.empty(self.B, dtype=torch.int32, device=self.device)
            ops.argmax_rows(probs, self._argmax_buf)
            self.generation += 1
            self._targets = self._argmax_buf
            ops.one_hot(self._targets, self.dprobs)
        else:
            self.set_targets(indices)
        if layer == self.stages[-1]["name"]:
            grad = self.head_grad_raw()  # the head's backward alone reaches the last endpoint
        else:  # any earlier endpoint: the data-gradient pass down to that layer's consumer, unmasked at the end
            prog, grad = self.raw_gradient_program(layer)
            self._run(prog)
        act = self.acts[layer]
        step = self.T // act.d  # pt/grad_cam_videos.py:112-113
        if out_hw is not None and cam is None:
            cam = torch.empty((self.B, act.d * step, out_hw[0], out_hw[1]), dtype=torch.float32, device=self.device)
        h_out, w_out = out_hw if out_hw is not None else (act.h, act.w)
        ops.gradcam(act, grad, step, h_out, w_out, per_frame, cam, cam_lowres=lowres)
        return cam, lowres, out

    @_lib.on_device
    def head_grad_raw(self):
        """fp32 gradient of sum(dprobs*probs) w.r.t. Mixed_5c (no ReLU mask): Grad-CAM's `grads_val`."""
        return ops.head_bwd(self.g_feat_raw, self.w_logits, self.softmax, self.probs, self.dprobs)
