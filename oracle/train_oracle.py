"""Oracle (test infrastructure, never imported by the product): one TRAINING step of the I3D classifier restated
functionally over a reference-keyed state dict - pt/train_i3d_smth.py:192-250 (model.train(); output = model(input);
loss = CrossEntropyLoss(output, target); loss.backward(); optimizer.step()) on the model of
pt/models/I3D_doubled.py (Unit3D :83-118 with BatchNorm3d(eps=1e-3, momentum=0.01) in training mode, the head
:360-371 with its dropout).  Plain torch autograd on the CPU, fp32 (or fp64 with dtype=torch.float64).

Pinned against the unmodified reference by oracle/pin_train_step.py, which also writes tests/golden/i3d_train.npz.
"""
import torch
import torch.nn.functional as F

from .i3d_oracle import ENDPOINTS, POOLS, _RoundGrad, _q, _same_pad

BN_EPS, BN_MOMENTUM = 1e-3, 0.01


def sample_index(key, numel, n=64):
    """seeded positions of the entries of tensor `key` that tests/golden/i3d_train.npz stores (a full gradient set
    is 49 MB): shared by oracle/pin_train_step.py and the tests"""
    g = torch.Generator().manual_seed(sum(key.encode()) * 7919 + numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def is_param(key):
    return key.endswith((".conv3d.weight", ".conv3d.bias", ".bn.weight", ".bn.bias"))


def _unit(p, buf, prefix, x, stride=(1, 1, 1), probe=None, quant=False):
    """quant: the rounding points of the mixed-precision step (train.py, mode 'bf16') - convolution weights, the
    stored convolution output z and the stored BatchNorm/ReLU output y rounded to bf16, and in the backward pass the
    stored dz (the gradient both the weight gradient and the data gradient read) rounded to bf16; BatchNorm
    statistics are taken from the rounded z, gradients summed over consumers stay fp32."""
    w = p[prefix + ".conv3d.weight"]
    x = F.conv3d(_same_pad(x, w.shape[2:], stride), _q(w) if quant else w, None, stride=stride)
    if quant:
        x = _q(_RoundGrad.apply(x))
    if probe is not None:
        probe[prefix + ":z"] = x
    x = F.batch_norm(x, buf[prefix + ".bn.running_mean"], buf[prefix + ".bn.running_var"], p[prefix + ".bn.weight"],
                     p[prefix + ".bn.bias"], training=True, momentum=BN_MOMENTUM, eps=BN_EPS)
    return _q(F.relu(x)) if quant else F.relu(x)


def _inception(p, buf, name, x, probe=None, quant=False):
    b0 = _unit(p, buf, name + ".b0", x, probe=probe, quant=quant)
    b1 = _unit(p, buf, name + ".b1b", _unit(p, buf, name + ".b1a", x, probe=probe, quant=quant), probe=probe, quant=quant)
    b2 = _unit(p, buf, name + ".b2b", _unit(p, buf, name + ".b2a", x, probe=probe, quant=quant), probe=probe, quant=quant)
    t3 = F.max_pool3d(_same_pad(x, (3, 3, 3), (1, 1, 1)), (3, 3, 3), (1, 1, 1))
    if probe is not None:
        probe[name + ".b3a:y"] = t3
    b3 = _unit(p, buf, name + ".b3b", t3, probe=probe, quant=quant)
    return torch.cat([b0, b1, b2, b3], dim=1)


def loss_and_grads(sd, x, target, avg_pool=(2, 7, 7), drop=None, dtype=torch.float32, probe=None, quant=False):
    """One forward/backward in training mode.  drop: optional [B, 1024] dropout mask already scaled by 1/keep
    (the reference draws it from torch's RNG; parity tests pass it in or disable dropout).
    Returns loss (float), logits [B, classes], grads {parameter key: tensor}, buffers {running stat key: tensor}
    as nn.BatchNorm3d leaves them after the step.  probe: optional dict that receives intermediate tensors
    ('<unit>:z' raw convolution outputs, '<module>.b3a:y' branch-pool outputs, '<endpoint>:y') and, under
    '<name>:grad', the loss gradient with respect to each of them (debugging aid of the tests)."""
    p = {k: v.detach().clone().to(dtype).requires_grad_() for k, v in sd.items() if is_param(k)}
    buf = {k: v.detach().clone().to(dtype) for k, v in sd.items() if ".bn.running_" in k}
    h = x.to(dtype)
    if quant:
        h = _q(h)  # the stem's tensor-core operand is the bf16 copy of the clip
    for name in ENDPOINTS:
        if name == "Conv3d_1a_7x7":
            h = _unit(p, buf, name, h, (2, 2, 2), probe=probe, quant=quant)
        elif name.startswith("Conv3d"):
            h = _unit(p, buf, name, h, probe=probe, quant=quant)
        elif name.startswith("MaxPool"):
            k, s = POOLS[name]
            h = F.max_pool3d(_same_pad(h, k, s), k, s)
        else:
            h = _inception(p, buf, name, h, probe=probe, quant=quant)
        if probe is not None:
            probe[name + ":y"] = h
    pooled = F.avg_pool3d(h, avg_pool, stride=(1, 1, 1))
    if drop is not None:
        pooled = pooled * drop.to(dtype).view(pooled.shape[0], -1, 1, 1, 1)
    logits = F.conv3d(pooled, p["logits.conv3d.weight"], p["logits.conv3d.bias"]).squeeze(3).squeeze(3).squeeze()
    if logits.dim() < 2:
        logits = logits[None, :]
    loss = F.cross_entropy(logits, target)
    keys = list(p)
    pk = list(probe) if probe is not None else []
    all_g = torch.autograd.grad(loss, [p[k] for k in keys] + [probe[k] for k in pk])
    grads = dict(zip(keys, all_g[:len(keys)]))
    for k, g in zip(pk, all_g[len(keys):]):
        probe[k + ":grad"] = g
        probe[k] = probe[k].detach()
    return float(loss.detach()), logits.detach(), grads, buf


def sgd_step(sd, grads, lr, momentum=0.0, weight_decay=0.0, state=None):
    """torch.optim.SGD semantics (no dampening, no Nesterov) on a copy of the parameters; state: {key: buffer}."""
    state = {} if state is None else state
    out = {}
    for k, g in grads.items():
        w = sd[k].detach().to(g.dtype)
        g = g + weight_decay * w
        if momentum:
            state[k] = g.clone() if k not in state else momentum * state[k] + g
            g = state[k]
        out[k] = w - lr * g
    return out, state


def adam_step(sd, grads, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, state=None, step=1):
    """torch.optim.Adam semantics (L2 weight decay added to the gradient)."""
    state = {} if state is None else state
    out = {}
    for k, g in grads.items():
        w = sd[k].detach().to(g.dtype)
        g = g + weight_decay * w
        m, v = state.get(k, (torch.zeros_like(g), torch.zeros_like(g)))
        m = betas[0] * m + (1 - betas[0]) * g
        v = betas[1] * v + (1 - betas[1]) * g * g
        state[k] = (m, v)
        c1, c2 = 1 - betas[0] ** step, 1 - betas[1] ** step
        out[k] = w - lr / c1 * m / (v.sqrt() / c2 ** 0.5 + eps)
    return out, state
