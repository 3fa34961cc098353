"""Shared implementation behind the drop-in models.I3D_doubled / models.I3D_doubled_kth modules.

The nn.Module tree, constructor arguments, parameter names and registration order mirror the
reference (pt/models/I3D_doubled.py:43-118 Unit3D, :8-40 MaxPool3dSamePadding, :121-146
InceptionModule, :149-380 Model; state-dict keys of SURVEY §3.4) so checkpoints load unchanged and
code that walks `model._modules` (pt/pytorch-grad-cam/grad-cam.py:23-54) keeps working — but no
torch operator computes anything: `Model.forward` runs the whole network through the libivf engine
(one autograd node), and each child module, when called on its own, runs its libivf kernels
through a per-op autograd node.  CUDA only; eval mode only (the path is interpretation).
"""
import os

import torch
import torch.nn as nn

from ... import _lib, ops
from ...engine import I3DEngine
from ...ops import Act, same_pad

_INCEPTION_SPECS = (
    ("Mixed_3b", 192, (64, 96, 128, 16, 32, 32)),
    ("Mixed_3c", 256, (128, 128, 192, 32, 96, 64)),
    ("MaxPool3d_4a_3x3", (3, 3, 3), (2, 2, 2)),
    ("Mixed_4b", 480, (192, 96, 208, 16, 48, 64)),
    ("Mixed_4c", 512, (160, 112, 224, 24, 64, 64)),
    ("Mixed_4d", 512, (128, 128, 256, 24, 64, 64)),
    ("Mixed_4e", 512, (112, 144, 288, 32, 64, 64)),
    ("Mixed_4f", 528, (256, 160, 320, 32, 128, 128)),
    ("MaxPool3d_5a_2x2", (2, 2, 2), (2, 2, 2)),
    ("Mixed_5b", 832, (256, 160, 320, 32, 128, 128)),
    ("Mixed_5c", 832, (384, 192, 384, 48, 128, 128)),
)


def default_mode():
    """bf16 tensor-core path unless IVF_FP32=1 (the 1e-4 parity mode)."""
    return "fp32" if os.environ.get("IVF_FP32", "0") == "1" else "bf16"


def _require_cuda(x, what):
    if not x.is_cuda:
        raise _lib.IvfError("%s: input is on %s; the native path runs on a B200 only (no CPU fallback)"
                            % (what, x.device))


# ------------------------------------------------------------------ per-op autograd nodes
def _to_act(x, dtype):
    """NCDHW tensor -> channels-last Act (zero-copy when already channels_last_3d in `dtype`)."""
    n, c, d, h, w = x.shape
    buf = x.permute(0, 2, 3, 4, 1).to(dtype).contiguous()
    return Act(buf, n, d, h, w, c, 0, c)


def _from_act(a):
    """channels-last Act -> NCDHW-shaped view (channels_last_3d strides), fp32 for the caller."""
    return a.tensor().permute(0, 4, 1, 2, 3).float()


class _UnitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod):
        from ...engine import Unit
        mode = mod.ivf_mode
        dt = torch.bfloat16 if mode == "bf16" else torch.float32
        unit = mod._packed(mode, x.device)
        if mode == "bf16" and unit.stride != (1, 1, 1):
            raise _lib.IvfError("calling a strided Unit3D on its own needs IVF_FP32=1 (the bf16 path presents "
                                "the stem space-to-depth inside Model.forward)")
        xa = _to_act(x, dt)
        geo = [same_pad(s, k, st) for s, k, st in zip(x.shape[2:], unit.kernel, unit.stride)]
        out = Act.empty(xa.n, geo[0][2], geo[1][2], geo[2][2], unit.cout, dt, x.device)
        pf = tuple(g[0] for g in geo)
        flags = _lib.EP_RELU if mod._activation_fn is not None else 0
        ops.conv3d(xa, unit.w_fwd, out, unit.kernel, unit.stride, pf, flags=flags, scale=unit.scale,
                   shift=unit.shift_with_bias)
        ctx.mod, ctx.unit, ctx.xa, ctx.out, ctx.pf, ctx.relu = mod, unit, xa, out, pf, bool(flags)
        return _from_act(out)

    @staticmethod
    def backward(ctx, gy):
        unit, xa, out = ctx.unit, ctx.xa, ctx.out
        dt = out.buf.dtype
        g = _to_act(gy, dt)
        # ReLU'/BN' of this unit: identity pool backward with the mask epilogue
        dz = out.like()
        am = torch.zeros((out.pixels, out.c), dtype=torch.uint8, device=gy.device)
        if ctx.relu:
            ops.maxpool3d_bwd(g, am, dz, (1, 1, 1), (1, 1, 1), (0, 0, 0), mask=out, mask_scale=unit.scale)
        elif ctx.mod._use_batch_norm:
            ops.maxpool3d_bwd(g, am, dz, (1, 1, 1), (1, 1, 1), (0, 0, 0), mask=_ones_like(out),
                              mask_scale=unit.scale)
        else:
            dz = g
        gx = Act.empty(xa.n, xa.d, xa.h, xa.w, xa.c, dt, gy.device)
        if dt == torch.float32:
            ops.conv3d(dz, unit.w_dgrad, gx, unit.kernel, unit.stride, ctx.pf, transposed=1)
        else:
            pf = tuple(k - 1 - p for k, p in zip(unit.kernel, ctx.pf))
            ops.conv3d(dz, unit.w_dgrad, gx, unit.kernel, (1, 1, 1), pf)
        return _from_act(gx), None


def _ones_like(a):
    return Act(torch.ones_like(a.buf), a.n, a.d, a.h, a.w, a.ld, a.coff, a.c)


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, stride, mode):
        dt = torch.bfloat16 if mode == "bf16" else torch.float32
        xa = _to_act(x, dt)
        geo = [same_pad(s, k, st) for s, k, st in zip(x.shape[2:], kernel, stride)]
        out = Act.empty(xa.n, geo[0][2], geo[1][2], geo[2][2], xa.c, dt, x.device)
        am = torch.empty((out.pixels, xa.c), dtype=torch.uint8, device=x.device)
        pf = tuple(g[0] for g in geo)
        ops.maxpool3d_fwd(xa, out, am, kernel, stride, pf)
        ctx.saved = (xa, am, kernel, stride, pf)
        return _from_act(out)

    @staticmethod
    def backward(ctx, gy):
        xa, am, kernel, stride, pf = ctx.saved
        g = _to_act(gy, xa.buf.dtype)
        gx = xa.like()
        ops.maxpool3d_bwd(g, am, gx, kernel, stride, pf)
        return _from_act(gx), None, None, None


# ------------------------------------------------------------------ modules
class MaxPool3dSamePadding(nn.MaxPool3d):
    """TF-'same' zero-padded max-pool (pt/models/I3D_doubled.py:8-40)."""

    ivf_mode = None

    def compute_pad(self, dim, s):
        return sum(same_pad(s, self.kernel_size[dim], self.stride[dim])[:2])

    def forward(self, x):
        _require_cuda(x, "MaxPool3dSamePadding")
        return _PoolFn.apply(x, tuple(self.kernel_size), tuple(self.stride), self.ivf_mode or default_mode())


class Unit3D(nn.Module):
    """conv3d (dynamic 'same' pad) -> BatchNorm3d(eps 1e-3) -> activation
    (pt/models/I3D_doubled.py:43-118)."""

    ivf_mode = None

    def __init__(self, in_channels, output_channels, kernel_shape=(1, 1, 1), stride=(1, 1, 1), padding=0,
                 activation_fn=torch.nn.functional.relu, use_batch_norm=True, use_bias=False, name="unit_3d"):
        super().__init__()
        self._output_channels = output_channels
        self._kernel_shape = kernel_shape
        self._stride = stride
        self._use_batch_norm = use_batch_norm
        self._activation_fn = activation_fn
        self._use_bias = use_bias
        self.name = name
        self.padding = padding
        self.conv3d = nn.Conv3d(in_channels, output_channels, kernel_size=kernel_shape, stride=stride,
                                padding=0, bias=use_bias)
        if use_batch_norm:
            self.bn = nn.BatchNorm3d(output_channels, eps=0.001, momentum=0.01)
        self._pack_cache = {}

    def compute_pad(self, dim, s):
        return sum(same_pad(s, self._kernel_shape[dim], self._stride[dim])[:2])

    def _packed(self, mode, device):
        from ...engine import Unit
        ver = tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())
        key = (mode, str(device))
        hit = self._pack_cache.get(key)
        if hit is None or hit[0] != ver:
            sd = {"u." + k: v for k, v in self.state_dict().items()}
            unit = Unit(sd, "u", tuple(self._stride), mode, device)  # shift_with_bias folds a conv bias
            hit = (ver, unit)
            self._pack_cache[key] = hit
        return hit[1]

    def forward(self, x):
        _require_cuda(x, "Unit3D")
        if self.training and self._use_batch_norm:
            raise _lib.IvfError("Unit3D: training-mode BatchNorm is outside the interpretation hot path; "
                                "call model.eval()")
        if self._activation_fn not in (None, torch.nn.functional.relu):
            raise _lib.IvfError("Unit3D: only ReLU / no activation have a native epilogue")
        if self.ivf_mode is None:
            self.ivf_mode = default_mode()
        return _UnitFn.apply(x, self)


class InceptionModule(nn.Module):
    """pt/models/I3D_doubled.py:121-146."""

    def __init__(self, in_channels, out_channels, name):
        super().__init__()
        oc = out_channels
        self.b0 = Unit3D(in_channels, oc[0], [1, 1, 1], name=name + "/Branch_0/Conv3d_0a_1x1")
        self.b1a = Unit3D(in_channels, oc[1], [1, 1, 1], name=name + "/Branch_1/Conv3d_0a_1x1")
        self.b1b = Unit3D(oc[1], oc[2], [3, 3, 3], name=name + "/Branch_1/Conv3d_0b_3x3")
        self.b2a = Unit3D(in_channels, oc[3], [1, 1, 1], name=name + "/Branch_2/Conv3d_0a_1x1")
        self.b2b = Unit3D(oc[3], oc[4], [3, 3, 3], name=name + "/Branch_2/Conv3d_0b_3x3")
        self.b3a = MaxPool3dSamePadding(kernel_size=[3, 3, 3], stride=(1, 1, 1), padding=0)
        self.b3b = Unit3D(in_channels, oc[5], [1, 1, 1], name=name + "/Branch_3/Conv3d_0b_1x1")
        self.name = name

    def forward(self, x):
        return torch.cat([self.b0(x), self.b1b(self.b1a(x)), self.b2b(self.b2a(x)), self.b3b(self.b3a(x))],
                         dim=1)


class _ModelFn(torch.autograd.Function):
    """The whole network as one autograd node over the engine's preallocated buffers."""

    @staticmethod
    def forward(ctx, x, model):
        eng = model._engine(x)
        eng.set_input(x.detach().contiguous())
        out = eng.forward(None).clone()
        # the engine bumps `generation` on everything that overwrites its activations or dprobs (forward, graph
        # replay, set_targets): a backward issued after, e.g., a Grad-CAM call on the same engine is refused
        ctx.eng, ctx.gen, ctx.shape = eng, eng.generation, x.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        eng = ctx.eng
        if eng.generation != ctx.gen:
            raise _lib.IvfError("backward through a forward whose activations were overwritten by a later "
                                "forward of the same model/geometry (the native model keeps one activation set)")
        eng.dprobs.copy_(gout)
        eng.generation += 1  # dprobs overwritten: a second backward through the same node needs a new forward
        eng.backward(to_mask=False)
        b, c, t, h, w = ctx.shape
        g = eng.g_xin.buf
        if eng.mode == "bf16":  # space-to-depth record -> NCDHW
            g = g.view(b, t // 2, h // 2, w // 2, 32)[..., :8 * c].float()
            g = g.view(b, t // 2, h // 2, w // 2, 2, 2, 2, c).permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(b, c, t, h, w)
        else:
            g = g.permute(0, 4, 1, 2, 3).contiguous()
        return g, None


class I3DBase(nn.Module):
    """Inception-v1 I3D with the reference's constructor surface (pt/models/I3D_doubled.py:186-335)."""

    MAX_ENGINE_GEOMETRIES = 3

    VALID_ENDPOINTS = ("Conv3d_1a_7x7", "MaxPool3d_2a_3x3", "Conv3d_2b_1x1", "Conv3d_2c_3x3",
                       "MaxPool3d_3a_3x3", "Mixed_3b", "Mixed_3c", "MaxPool3d_4a_3x3", "Mixed_4b", "Mixed_4c",
                       "Mixed_4d", "Mixed_4e", "Mixed_4f", "MaxPool3d_5a_2x2", "Mixed_5b", "Mixed_5c", "Logits",
                       "Predictions")

    def _init_i3d(self, num_classes, spatial_squeeze, final_endpoint, name, in_channels, dropout_keep_prob,
                  last_stride, stride_mod_layers, softMax, lastRelu, pool_time, pool_hw):
        if final_endpoint not in self.VALID_ENDPOINTS:
            raise ValueError("Unknown final endpoint %s" % final_endpoint)
        if final_endpoint != "Logits":
            raise NotImplementedError("the native model is built up to 'Logits' only")
        if stride_mod_layers is None:  # argparse default; the reference raises TypeError here (SURVEY bug 5)
            stride_mod_layers = ""
        self._num_classes = num_classes
        self._spatial_squeeze = spatial_squeeze
        self._final_endpoint = final_endpoint
        self.logits = None
        self.softMax = softMax
        self.lastRelu = lastRelu
        self.sm = nn.Softmax(dim=1)
        self.ivf_mode = default_mode()
        self._engines = {}
        mods = stride_mod_layers.split(",") if isinstance(stride_mod_layers, str) and stride_mod_layers else \
            list(stride_mod_layers or [])
        self._stride_mods = {}

        def tstride(ep):
            s = last_stride if ep in mods else 2
            if ep in mods:
                self._stride_mods[ep] = (s, 2, 2)
            return s

        ep = {}
        ep["Conv3d_1a_7x7"] = Unit3D(in_channels, 64, [7, 7, 7], stride=(tstride("Conv3d_1a_7x7"), 2, 2),
                                     padding=(3, 3, 3), name=name + "Conv3d_1a_7x7")
        ep["MaxPool3d_2a_3x3"] = MaxPool3dSamePadding(kernel_size=[1, 3, 3], stride=(1, 2, 2), padding=0)
        ep["Conv3d_2b_1x1"] = Unit3D(64, 64, [1, 1, 1], name=name + "Conv3d_2b_1x1")
        ep["Conv3d_2c_3x3"] = Unit3D(64, 192, [3, 3, 3], padding=1, name=name + "Conv3d_2c_3x3")
        ep["MaxPool3d_3a_3x3"] = MaxPool3dSamePadding(kernel_size=[1, 3, 3], stride=(1, 2, 2), padding=0)
        for spec in _INCEPTION_SPECS:
            if spec[0].startswith("Mixed"):
                ep[spec[0]] = InceptionModule(spec[1], list(spec[2]), name + spec[0])
            else:
                ep[spec[0]] = MaxPool3dSamePadding(kernel_size=list(spec[1]),
                                                   stride=(tstride(spec[0]), spec[2][1], spec[2][2]), padding=0)
        self.end_points = ep
        if not mods:
            kt = pool_time
        else:
            kt = int(pool_time * ((2 / last_stride) ** len(mods)))
        self.avg_pool = nn.AvgPool3d(kernel_size=[kt, pool_hw[0], pool_hw[1]], stride=(1, 1, 1))
        self.dropout = nn.Dropout(dropout_keep_prob)
        last_actf = torch.nn.functional.relu if lastRelu == "relu" else None
        self.logits = Unit3D(1024, num_classes, [1, 1, 1], activation_fn=last_actf, use_batch_norm=False,
                             use_bias=True, name="logits")
        self.build()

    def replace_logits(self, num_classes):
        self._num_classes = num_classes
        self.logits = Unit3D(1024, num_classes, [1, 1, 1], activation_fn=None, use_batch_norm=False,
                             use_bias=True, name="logits")
        self._engines = {}

    def build(self):
        for k, m in self.end_points.items():
            self.add_module(k, m)

    # -- native engine management
    def _version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def _engine(self, x, batch=None, tag=0):
        """The native runner for this geometry (cached).  `tag` distinguishes several engines of the same
        geometry that are alive together (clip groups of the batched mask search)."""
        b, c, t, h, w = x.shape
        device = next(self.parameters()).device  # the model's GPU; x may still be on the host
        if device.type != "cuda":
            raise _lib.IvfError("the native I3D needs the model on a CUDA device (model.cuda())")
        key = (batch or b, c, t, h, w, self.ivf_mode, str(device), tag)
        ver = self._version()
        hit = self._engines.get(key)
        if hit is None or hit[0] != ver:
            if self.training:
                raise _lib.IvfError("the native I3D runs eval-mode BatchNorm only; call model.eval()")
            # pt/models/I3D_doubled.py:321-326: `if leaky: leaky_relu` is followed by `if relu: relu else: None`,
            # so "leaky" ends up with NO activation upstream (and so here); only "relu" activates the logits,
            # and the native head has no ReLU epilogue after the logits.
            if self.lastRelu == "relu":
                raise _lib.IvfError("lastRelu='relu' (ReLU on the logits) has no native head epilogue")
            sd = {k: v for k, v in self.state_dict().items()}
            eng = I3DEngine(sd, batch or b, (t, h, w), mode=self.ivf_mode, softmax=bool(self.softMax),
                            avg_pool=tuple(self.avg_pool.kernel_size), stride_mods=self._stride_mods,
                            device=device, in_channels=c)
            hit = (ver, eng)
            # a small LRU over geometries: the reference driver alternates model(input_var) at B = batch_size
            # with Grad-CAM at B = 1 for every clip (pt/FindMasksComparison_I3D_smth.py:176,266); rebuilding an
            # engine per switch (weight packing, buffers, graphs) would cost far more than the kernels.  Engines
            # of older weights go first; clip groups of one geometry (same key up to `tag`) count as one entry.
            self._engines = {k: v for k, v in self._engines.items() if v[0] == ver}
            geos = []
            for k in self._engines:
                if k[:-1] not in geos:
                    geos.append(k[:-1])
            while len(geos) >= self.MAX_ENGINE_GEOMETRIES and geos[0] != key[:-1]:
                drop = geos.pop(0)
                self._engines = {k: v for k, v in self._engines.items() if k[:-1] != drop}
        else:
            self._engines.pop(key)  # re-insert: most recently used last
        self._engines[key] = hit
        return hit[1]

    def set_mode(self, mode):
        assert mode in ("bf16", "fp32")
        self.ivf_mode = mode
        for m in self.modules():
            if isinstance(m, (Unit3D, MaxPool3dSamePadding)):
                m.ivf_mode = mode
        return self

    def forward(self, x):
        _require_cuda(x, "I3D Model")
        if not self._spatial_squeeze:
            raise _lib.IvfError("spatial_squeeze=False is not supported by the native head")
        return _ModelFn.apply(x, self)

    def extract_features(self, x):
        """pt/models/I3D_doubled.py:382-388: endpoints then avg_pool (per-op native kernels)."""
        for end_point in self.VALID_ENDPOINTS:
            if end_point in self.end_points:
                x = self._modules[end_point](x)
        return self.avg_pool(x)  # stock nn.AvgPool3d child, kept for the reference's surface
