"""Oracle (test infrastructure): Inception-v1 I3D forward restated functionally from
pt/models/I3D_doubled.py and pt/models/I3D_doubled_kth.py over a reference-keyed state dict.
CPU/any device, fp32, plain torch.nn.functional — differentiable through autograd.
"""
import math

import torch
import torch.nn.functional as F

ENDPOINTS = ("Conv3d_1a_7x7", "MaxPool3d_2a_3x3", "Conv3d_2b_1x1", "Conv3d_2c_3x3", "MaxPool3d_3a_3x3",
             "Mixed_3b", "Mixed_3c", "MaxPool3d_4a_3x3", "Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e",
             "Mixed_4f", "MaxPool3d_5a_2x2", "Mixed_5b", "Mixed_5c")
POOLS = {"MaxPool3d_2a_3x3": ((1, 3, 3), (1, 2, 2)), "MaxPool3d_3a_3x3": ((1, 3, 3), (1, 2, 2)),
         "MaxPool3d_4a_3x3": ((3, 3, 3), (2, 2, 2)), "MaxPool3d_5a_2x2": ((2, 2, 2), (2, 2, 2))}


def _same_pad(x, kernel, stride):
    """pt/models/I3D_doubled.py:77-106 (Unit3D) / :9-38 (MaxPool3dSamePadding)."""
    pads = []
    for size, k, s in zip(x.shape[2:], kernel, stride):
        total = max(k - s, 0) if size % s == 0 else max(k - (size % s), 0)
        pads.append((total // 2, total - total // 2))
    (tf, tb), (hf, hb), (wf, wb) = pads
    return F.pad(x, (wf, wb, hf, hb, tf, tb))


# --- optional bf16 quantisation points ---------------------------------------------------------
# quant=True restates the SAME network as the mixed-precision computation the bf16 tensor-core path
# is specified to perform: conv weights and every stored activation rounded to bf16 (fp32
# accumulation, fp32 BatchNorm affine, fp32 head), and in the backward pass every stored gradient
# tensor rounded to bf16 (gradient w.r.t. each conv output after ReLU'/BN', w.r.t. pool outputs and
# w.r.t. the network input); sums over the consumers of a tensor stay fp32.  A random deep ReLU
# network amplifies ANY perturbation of its features (measured: 0.4 % after the stem, >50 % at
# Mixed_5c for bf16 rounding alone), so bf16 results are checked against this matched-rounding
# restatement, and against the plain fp32 oracle only where the north star says so.
class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _q(x):
    """round to bf16 in the forward pass, identity gradient"""
    return x + (x.bfloat16().float() - x).detach()


def quant_input(x):
    """the bf16 operand the stem convolution reads, with a bf16-stored input gradient"""
    return _q(_RoundGrad.apply(x))


def unit3d(sd, prefix, x, stride=(1, 1, 1), relu=True, quant=False, force=None):
    """pt/models/I3D_doubled.py:83-118: pad -> conv3d -> BatchNorm3d(eval, eps 1e-3) -> ReLU.
    force: optional {prefix: bool mask} of imposed ReLU decisions (see features())."""
    w = sd[prefix + ".conv3d.weight"]
    b = sd.get(prefix + ".conv3d.bias")
    if quant:
        w = w.bfloat16().float()
    x = F.conv3d(_same_pad(x, w.shape[2:], stride), w, b, stride=stride)
    if quant:
        x = _RoundGrad.apply(x)
    if prefix + ".bn.weight" in sd:
        x = F.batch_norm(x, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"],
                         sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], training=False, eps=1e-3)
    if relu and force is not None and prefix in force:
        x = x * force[prefix].to(x.dtype)  # the imposed activation pattern instead of this evaluation's own
    else:
        x = F.relu(x) if relu else x
    return _q(x) if quant else x


def forced_pool(x, kernel, stride, idx):
    """Max-pool with IMPOSED routing: every output takes the window element number idx (scan order
    (kd*KH + kh)*KW + kw over the zero-padded window) instead of its own arg-maximum."""
    xp = _same_pad(x, kernel, stride)
    u = xp.unfold(2, kernel[0], stride[0]).unfold(3, kernel[1], stride[1]).unfold(4, kernel[2], stride[2])
    u = u.reshape(*u.shape[:5], -1)
    return u.gather(-1, idx.long().unsqueeze(-1)).squeeze(-1)


def maxpool_same(x, kernel, stride, quant=False, force_idx=None):
    if force_idx is not None:
        y = forced_pool(x, kernel, stride, force_idx)
    else:
        y = F.max_pool3d(_same_pad(x, kernel, stride), kernel, stride)
    return _RoundGrad.apply(y) if quant else y


def inception(sd, name, x, quant=False, force=None):
    """pt/models/I3D_doubled.py:121-146."""
    fi = None if force is None else force.get(name + ".b3a")
    b0 = unit3d(sd, name + ".b0", x, quant=quant, force=force)
    b1 = unit3d(sd, name + ".b1b", unit3d(sd, name + ".b1a", x, quant=quant, force=force), quant=quant, force=force)
    b2 = unit3d(sd, name + ".b2b", unit3d(sd, name + ".b2a", x, quant=quant, force=force), quant=quant, force=force)
    b3 = unit3d(sd, name + ".b3b", maxpool_same(x, (3, 3, 3), (1, 1, 1), quant, fi), quant=quant, force=force)
    return torch.cat([b0, b1, b2, b3], dim=1)


def features(sd, x, upto="Mixed_5c", stride_mods=None, quant=False, force=None, start_after=None):
    """force: optional dict of IMPOSED decisions - {unit prefix: bool ReLU mask [N,C,D,H,W]} and
    {pool name (Inception branch pools: '<module>.b3a'): window index [N,C,od,oh,ow]}.  With every decision
    imposed the network is a fixed linear map of its input, so two evaluations that agree on the decisions
    agree on the gradient up to rounding: the tests use this to separate the bf16 path's arithmetic error
    from the re-routing that flipped ReLU / arg-max decisions cause."""
    stride_mods = stride_mods or {}
    outs = {}
    if quant and start_after is None:
        x = quant_input(x)
    skipping = start_after is not None  # x is the output of endpoint `start_after`: run the rest of the network
    for name in ENDPOINTS:
        if skipping:
            skipping = name != start_after
            continue
        if name == "Conv3d_1a_7x7":
            x = unit3d(sd, name, x, stride_mods.get(name, (2, 2, 2)), quant=quant, force=force)
        elif name.startswith("Conv3d"):
            x = unit3d(sd, name, x, quant=quant, force=force)
        elif name.startswith("MaxPool"):
            k, s = POOLS[name]
            x = maxpool_same(x, k, stride_mods.get(name, s), quant, None if force is None else force.get(name))
        else:
            x = inception(sd, name, x, quant, force)
        outs[name] = x
        if name == upto:
            break
    return x, outs


def head(sd, feat, avg_pool=(2, 7, 7), softmax=True):
    """pt/models/I3D_doubled.py:360-371: avg_pool(stride 1) -> dropout(eval) -> logits(+bias) ->
    squeeze(3).squeeze(3).squeeze() -> [None] if 1-D -> softmax(dim=1)."""
    x = F.avg_pool3d(feat, avg_pool, stride=(1, 1, 1))
    x = F.conv3d(x, sd["logits.conv3d.weight"], sd["logits.conv3d.bias"])
    logits = x.squeeze(3).squeeze(3).squeeze()
    if logits.dim() < 2:
        logits = logits[None, :]
    return F.softmax(logits, dim=1) if softmax else logits


def forward(sd, x, avg_pool=(2, 7, 7), softmax=True, stride_mods=None, quant=False, force=None):
    feat, _ = features(sd, x, stride_mods=stride_mods, quant=quant, force=force)
    return head(sd, feat, avg_pool, softmax)


class Model:
    """Callable wrapper so the mask oracle can call model(x) like the reference does."""

    def __init__(self, sd, avg_pool=(2, 7, 7), softmax=True, quant=False, stride_mods=None):
        self.sd, self.avg_pool, self.softmax, self.quant, self.stride_mods = sd, avg_pool, softmax, quant, stride_mods

    def __call__(self, x):
        return forward(self.sd, x, self.avg_pool, self.softmax, stride_mods=self.stride_mods, quant=self.quant)


def conv_flops_per_clip(sd, clip_shape):
    """2*MACs of every convolution for one clip (SURVEY §3.4's GF column), by shape propagation."""
    t, h, w = clip_shape
    total = 0

    def conv(prefix, dims, stride=(1, 1, 1)):
        nonlocal total
        wt = sd[prefix + ".conv3d.weight"]
        out = tuple(int(math.ceil(d / s)) for d, s in zip(dims, stride))
        total += 2 * out[0] * out[1] * out[2] * wt.shape[0] * wt.shape[1] * wt.shape[2] * wt.shape[3] * wt.shape[4]
        return out

    dims = (t, h, w)
    for name in ENDPOINTS:
        if name == "Conv3d_1a_7x7":
            dims = conv(name, dims, (2, 2, 2))
        elif name.startswith("Conv3d"):
            dims = conv(name, dims)
        elif name.startswith("MaxPool"):
            dims = tuple(int(math.ceil(d / s)) for d, s in zip(dims, POOLS[name][1]))
        else:
            for b in ("b0", "b1a", "b1b", "b2a", "b2b", "b3b"):
                conv(name + "." + b, dims)
    return total


# ---------------------------------------------------------------------------------------------
# "Sharpened" deterministic initialisation: with PyTorch-default random weights the
# class gradient w.r.t. the mask is ~1e-9, five orders below the regulariser's, so trajectory/IoU
# tests would pass with a broken conv backward.  calibrate_and_sharpen() rewrites a state dict so
# that (1) every BatchNorm's running stats equal the batch statistics of `x` (activations stay
# O(1) through the depth) and (2) the logits layer is scaled so max softmax prob ~= target_prob.
def calibrate_and_sharpen(sd, x, avg_pool=(2, 7, 7), target_prob=0.5):
    sd = {k: v.clone() for k, v in sd.items()}

    def unit_cal(prefix, inp, stride=(1, 1, 1)):
        w = sd[prefix + ".conv3d.weight"]
        z = F.conv3d(_same_pad(inp, w.shape[2:], stride), w, None, stride=stride)
        sd[prefix + ".bn.running_mean"] = z.mean(dim=(0, 2, 3, 4))
        sd[prefix + ".bn.running_var"] = z.var(dim=(0, 2, 3, 4), unbiased=False)
        return unit3d(sd, prefix, inp, stride)

    with torch.no_grad():
        for name in ENDPOINTS:
            if name == "Conv3d_1a_7x7":
                x = unit_cal(name, x, (2, 2, 2))
            elif name.startswith("Conv3d"):
                x = unit_cal(name, x)
            elif name.startswith("MaxPool"):
                k, s = POOLS[name]
                x = maxpool_same(x, k, s)
            else:
                b0 = unit_cal(name + ".b0", x)
                b1 = unit_cal(name + ".b1b", unit_cal(name + ".b1a", x))
                b2 = unit_cal(name + ".b2b", unit_cal(name + ".b2a", x))
                b3 = unit_cal(name + ".b3b", maxpool_same(x, (3, 3, 3), (1, 1, 1)))
                x = torch.cat([b0, b1, b2, b3], dim=1)
        logits = head(sd, x, avg_pool, softmax=False)
        lo, hi = 0.0, 1e6
        for _ in range(80):  # bisection on the logit scale
            mid = 0.5 * (lo + hi)
            p = F.softmax(mid * logits, dim=1).max(dim=1)[0].mean().item()
            if p < target_prob:
                lo = mid
            else:
                hi = mid
        alpha = 0.5 * (lo + hi)
        sd["logits.conv3d.weight"] = sd["logits.conv3d.weight"] * alpha
        sd["logits.conv3d.bias"] = sd["logits.conv3d.bias"] * alpha
    return sd


def sharpen_head_only(sd, x, avg_pool=(2, 7, 7), target_prob=0.5, stride_mods=None):
    """Default-initialised trunk (well conditioned under bf16 rounding: features agree with fp32 to < 1e-2 up to
    Mixed_5c) with a head that makes the class term matter: the logit bias centres the logits over the clips `x`
    and the logits layer is scaled until the mean top probability is target_prob.  Unlike
    calibrate_and_sharpen the BatchNorm statistics stay at their initial values."""
    sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        feat, _ = features(sd, x, stride_mods=stride_mods)
        sd["logits.conv3d.bias"] = torch.zeros_like(sd["logits.conv3d.bias"])
        lg = head(sd, feat, avg_pool, softmax=False)
        sd["logits.conv3d.bias"] = -lg.mean(0)
        lg = lg - lg.mean(0)
        lo, hi = 0.0, 1e12
        for _ in range(100):
            mid = 0.5 * (lo + hi)
            p = F.softmax(mid * lg, dim=1).max(dim=1)[0].mean().item()
            if p < target_prob:
                lo = mid
            else:
                hi = mid
        alpha = 0.5 * (lo + hi)
        sd["logits.conv3d.weight"] = sd["logits.conv3d.weight"] * alpha
        sd["logits.conv3d.bias"] = sd["logits.conv3d.bias"] * alpha
    return sd
