"""Native training step of the I3D classifier - what pt/train_i3d_smth.py:192-250 does per
batch: model.train(); output = model(input); loss = CrossEntropyLoss(output, target); loss.backward();
optimizer.step() - as a fixed program of libivf launches over channels-last buffers:

  forward   per Unit3D (pt/models/I3D_doubled.py:83-118): ivf_conv3d (raw output z) -> ivf_bn_train_fwd
            (batch statistics, running statistics updated, ReLU) into the unit's slice of the Inception concat;
            max-pools as in the interpretation path; ivf_head_train_fwd (average pool, dropout, logits, loss);
  backward  ivf_head_train_bwd, then per unit in reverse: ivf_bn_train_bwd (dgamma, dbeta, dz) ->
            ivf_conv3d_wgrad (dW) -> ivf_conv3d data gradient, accumulated over the consumers of a tensor;
  update    ivf_optim_step per parameter (torch.optim.SGD / Adam semantics), weights re-packed for the next step.

Two modes.  "fp32" (the reference trains in fp32): every tensor fp32, convolutions and data gradients on the fp32
implicit-GEMM kernel.  "bf16" (mixed precision): activations and convolution operands bf16 - forward convolutions and
data gradients are the tcgen05 implicit GEMMs of the interpretation path (the stem reads its space-to-depth copy of
the clips) - while parameters, BatchNorm statistics, every gradient that is summed over several consumers, the weight
gradients and the optimizer state stay fp32.  The weight gradient is a CUDA-core kernel in both modes.
Parameters are fp32 device tensors updated IN PLACE: built from a drop-in model (`I3DTrainer.from_model`) they are the
model's own parameter storage, so the model sees every step.  torch is used for allocation only.
"""
import torch

from . import _lib, ops
from ._lib import PFMT_NDHWC_F32, PFMT_S2D_BF16
from .engine import ENDPOINTS, POOLS, pack, strip_module_prefix
from .ops import Act, same_pad

BN_EPS, BN_MOMENTUM = 1e-3, 0.01  # nn.BatchNorm3d(eps=0.001, momentum=0.01), pt/models/I3D_doubled.py:75


def _is_param(key):
    return key.endswith((".conv3d.weight", ".conv3d.bias", ".bn.weight", ".bn.bias"))


class _TrainUnit:
    """One Unit3D in training mode: parameters (views of the trainer's fp32 tensors), packed operands, buffers.
    x_conv: what the forward convolution reads (bf16 stem: the space-to-depth copy), x: the unit's input as the
    weight gradient reads it."""

    def __init__(self, tr, prefix, stride, x, y, x_conv=None):
        P, dev = tr.params, tr.device
        self.mode, self.prefix, self.stride, self.x, self.y = tr.mode, prefix, tuple(stride), x, y
        self.w = P[prefix + ".conv3d.weight"]
        self.gamma, self.beta = P[prefix + ".bn.weight"], P[prefix + ".bn.bias"]
        self.rmean, self.rvar = tr.buffers[prefix + ".bn.running_mean"], tr.buffers[prefix + ".bn.running_var"]
        self.cout, self.cin = self.w.shape[0], self.w.shape[1]
        self.kernel = tuple(self.w.shape[2:])
        self.pf = tuple(same_pad(sz, k, s)[0] for sz, k, s in zip((x.d, x.h, x.w), self.kernel, self.stride))
        assert y.c == self.cout and x.c == self.cin
        self.s2d = x_conv is not None  # bf16 stem (pt/models/I3D_doubled.py:233-235): stride 2 as a stride-1 4x4x4
        if self.s2d:                   # convolution over the 2x2x2 space-to-depth record, as in engine.Unit
            assert self.stride == (2, 2, 2) and tr.mode == "bf16"
            win = (self.kernel[0] + 1) // 2
            self.conv_x = Act(x_conv.buf, x_conv.n, x_conv.d, x_conv.h, x_conv.w, x_conv.ld, 0, 8 * self.cin)
            self.conv_kernel, self.conv_stride, self.conv_pf, self.f = (win,) * 3, (1, 1, 1), (1, 1, 1), (2, 2, 2)
        else:
            self.conv_x, self.conv_kernel, self.conv_stride, self.conv_pf, self.f = x, self.kernel, self.stride, self.pf, (1, 1, 1)
        # data gradient: fp32 = transposed gather with the forward pads; bf16 = stride-1 convolution with flipped taps
        self.dgrad_pf = self.pf if tr.mode == "fp32" else tuple(k - 1 - p for k, p in zip(self.kernel, self.pf))
        self.z = Act.empty(y.n, y.d, y.h, y.w, self.cout, tr.adt, dev)  # raw convolution output, then dz
        self.save_mean = torch.empty(self.cout, dtype=torch.float32, device=dev)
        self.save_rstd = torch.empty(self.cout, dtype=torch.float32, device=dev)
        self.ws = torch.empty(2 * self.cout, dtype=torch.float64, device=dev)
        self.repack()

    w_fwd = w_dgrad = None

    def repack(self):
        self.w_fwd = pack([self.w], self.mode, s2d=self.f, out=self.w_fwd)
        self.w_dgrad = None if self.s2d else pack([self.w], self.mode, dgrad=True, out=self.w_dgrad)

    def dgrad(self, dz, gx, acc):
        if self.mode == "fp32":
            ops.conv3d(dz, self.w_dgrad, gx, self.kernel, self.stride, self.dgrad_pf, acc_in=gx if acc else None,
                       transposed=1)
        else:
            assert self.stride == (1, 1, 1)
            ops.conv3d(dz, self.w_dgrad, gx, self.kernel, (1, 1, 1), self.dgrad_pf, acc_in=gx if acc else None)


class I3DTrainer:
    def __init__(self, state_dict, batch, clip, avg_pool=(2, 7, 7), device=None, optimizer="sgd", lr=0.01,
                 momentum=0.9, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, dropout_p=0.0, seed=0, in_channels=3,
                 share_storage=False, mode="fp32", world=1):
        """state_dict: the reference-keyed parameters and BatchNorm buffers (pt/models/I3D_doubled.py).  dropout_p is
        what the reference hands to nn.Dropout (its `dropout_keep_prob` argument, pt/models/I3D_doubled.py:319).
        share_storage: fp32 contiguous tensors of state_dict that already live on the device are updated in place
        instead of copied (from_model).  world > 1: data parallel over the ranks of torch.distributed's default group
        (one process per GPU, each with its own clips): the parameter gradients - one flat fp32 buffer - are summed with
        ONE all-reduce per step and averaged inside the update kernel; BatchNorm statistics stay per rank, as under
        the reference's nn.DataParallel replicas (pt/train_i3d_smth.py:58)."""
        if optimizer not in ("sgd", "adam"):
            raise _lib.IvfError("optimizer must be 'sgd' or 'adam' (pt/train_i3d_smth.py:131-138)")
        if mode not in ("fp32", "bf16"):
            raise _lib.IvfError("mode must be 'fp32' or 'bf16'")
        self.mode, self.adt = mode, (torch.float32 if mode == "fp32" else torch.bfloat16)
        self.device = dev = torch.device(device if device is not None else "cuda")
        _lib.handle(dev)  # fails loudly without the extension / a GPU
        sd = strip_module_prefix(state_dict)
        self.B, (self.T, self.H, self.W), self.C = batch, clip, in_channels
        self.opt, self.lr, self.momentum, self.betas, self.eps, self.wd = optimizer, lr, momentum, betas, eps, weight_decay
        self.dropout_p, self.seed, self.step_count = float(dropout_p), int(seed), 0

        def shared(t):
            return share_storage and t.is_cuda and t.device == dev and t.dtype == torch.float32 and t.is_contiguous()

        def own(t):
            if shared(t):
                return t.detach()
            return t.detach().to(device=dev, dtype=torch.float32).contiguous().clone()

        # tensors of the caller that this trainer writes in place: their autograd version counters are bumped after
        # every update so that whoever caches by version (the drop-in model's engines) sees the change
        self._shared = [v for k, v in sd.items() if (_is_param(k) or ".bn.running_" in k) and shared(v)]
        self.params = {k: own(v) for k, v in sd.items() if _is_param(k)}
        self.buffers = {k: own(v) for k, v in sd.items() if ".bn.running_" in k}
        # all parameter gradients are views of ONE flat buffer (16-byte aligned pieces): one all-reduce for data
        # parallelism, one table-driven launch for the update
        offs, total = {}, 0
        for k, v in self.params.items():
            offs[k] = total
            total += (v.numel() + 3) // 4 * 4
        self.flat_grads = ops.zeros((total,), torch.float32, dev)
        self.grads = {k: self.flat_grads[offs[k]:offs[k] + v.numel()].view(v.shape) for k, v in self.params.items()}
        self.state1 = {k: ops.zeros(v.shape, torch.float32, dev) for k, v in self.params.items()}
        self.state2 = {k: ops.zeros(v.shape, torch.float32, dev) for k, v in self.params.items()} \
            if optimizer == "adam" else {}
        keys = list(self.params)
        self.world = int(world)
        self._table = ops.optim_table([self.params[k] for k in keys], [self.grads[k] for k in keys],
                                      [self.state1[k] for k in keys], [self.state2.get(k) for k in keys], dev)

        B = batch
        self.x = ops.zeros((B, in_channels, self.T, self.H, self.W), torch.float32, dev)
        self.zero_mask = ops.zeros((B, self.T), torch.float32, dev)
        self.xin = Act.empty(B, self.T, self.H, self.W, in_channels, torch.float32, dev)
        self.xin_s2d = None
        if mode == "bf16":  # the stem's tensor-core operand (the weight gradient reads xin itself)
            if self.T % 2 or self.H % 2 or self.W % 2 or in_channels > 4:
                raise _lib.IvfError("bf16 mode needs even T/H/W and at most 4 input channels; use mode='fp32'")
            self.xin_s2d = Act.empty(B, self.T // 2, self.H // 2, self.W // 2, 32, torch.bfloat16, dev, zero=True)
        self.fwd, tape = [], []  # launches of the forward pass; records the backward pass is derived from
        self.units = []

        def new_act(n, d, h, w, c):
            return Act.empty(n, d, h, w, c, self.adt, dev)

        def out_dims(x, k, s):
            return tuple(same_pad(sz, kk, ss)[2] for sz, kk, ss in zip((x.d, x.h, x.w), k, s))

        def add_unit(prefix, x, y, stride=(1, 1, 1), x_conv=None):
            u = _TrainUnit(self, prefix, stride, x, y, x_conv)
            self.units.append(u)
            self.fwd.append(lambda u=u: ops.conv3d(u.conv_x, u.w_fwd, u.z, u.conv_kernel, u.conv_stride, u.conv_pf))
            self.fwd.append(lambda u=u: ops.bn_train_fwd(u.z, u.gamma, u.beta, BN_EPS, BN_MOMENTUM, u.rmean, u.rvar,
                                                         u.save_mean, u.save_rstd, u.ws, u.y, relu=True))
            tape.append(("unit", u))

        def add_pool(x, y, k, s):
            pads = tuple(same_pad(sz, kk, ss)[0] for sz, kk, ss in zip((x.d, x.h, x.w), k, s))
            am = torch.empty((y.pixels, x.c), dtype=torch.uint8, device=dev)
            self.fwd.append(lambda: ops.maxpool3d_fwd(x, y, am, k, s, pads, nonneg=True))  # inputs are ReLU outputs
            tape.append(("pool", x, y, am, k, s, pads))

        x = self.xin
        for name in ENDPOINTS:
            if name.startswith("Conv3d"):
                stride = (2, 2, 2) if name == "Conv3d_1a_7x7" else (1, 1, 1)
                w = self.params[name + ".conv3d.weight"]
                od, oh, ow = out_dims(x, tuple(w.shape[2:]), stride)
                y = new_act(B, od, oh, ow, w.shape[0])
                add_unit(name, x, y, stride, x_conv=self.xin_s2d if name == "Conv3d_1a_7x7" else None)
            elif name.startswith("MaxPool"):
                k, s = POOLS[name]
                y = new_act(B, *out_dims(x, k, s), x.c)
                add_pool(x, y, k, s)
            else:  # InceptionModule, pt/models/I3D_doubled.py:121-146: cat([b0, b1b(b1a), b2b(b2a), b3b(pool)])
                co = {b: self.params["%s.%s.conv3d.weight" % (name, b)].shape[0]
                      for b in ("b0", "b1a", "b1b", "b2a", "b2b", "b3b")}
                y = new_act(x.n, x.d, x.h, x.w, co["b0"] + co["b1b"] + co["b2b"] + co["b3b"])
                t1, t2 = new_act(x.n, x.d, x.h, x.w, co["b1a"]), new_act(x.n, x.d, x.h, x.w, co["b2a"])
                t3 = new_act(x.n, x.d, x.h, x.w, x.c)
                add_unit(name + ".b0", x, y.slice(0, co["b0"]))
                add_unit(name + ".b1a", x, t1)
                add_unit(name + ".b1b", t1, y.slice(co["b0"], co["b1b"]))
                add_unit(name + ".b2a", x, t2)
                add_unit(name + ".b2b", t2, y.slice(co["b0"] + co["b1b"], co["b2b"]))
                add_pool(x, t3, (3, 3, 3), (1, 1, 1))
                add_unit(name + ".b3b", t3, y.slice(co["b0"] + co["b1b"] + co["b2b"], co["b3b"]))
            x = y
        feat = x
        if (feat.d, feat.h, feat.w) != tuple(avg_pool):
            raise _lib.IvfError("Mixed_5c map %s is not covered by avg_pool %s (one pooled position per clip is what "
                                "CrossEntropyLoss needs, pt/models/I3D_doubled.py:360-371)"
                                % ((feat.d, feat.h, feat.w), tuple(avg_pool)))
        self.feat = feat
        wl = self.params["logits.conv3d.weight"]
        self.num_classes, cf = wl.shape[0], feat.c
        self.w_logits = wl.view(self.num_classes, cf)  # [classes, 1024, 1, 1, 1] is the same memory
        self.b_logits = self.params["logits.conv3d.bias"]
        self.pooled = torch.empty((B, cf), dtype=torch.float32, device=dev)
        self.drop = torch.empty((B, cf), dtype=torch.float32, device=dev) if self.dropout_p > 0 else None
        self.logits = torch.empty((B, self.num_classes), dtype=torch.float32, device=dev)
        self.dlogits = torch.empty((B, self.num_classes), dtype=torch.float32, device=dev)
        self.loss = torch.empty((1,), dtype=torch.float32, device=dev)
        self.target = torch.zeros((B,), dtype=torch.int32, device=dev)

        # ---- backward program: the tape in reverse.  A tensor consumed several times (an Inception module's input)
        # gets its gradient from the consumer that runs FIRST in the backward pass, the others accumulate onto it.
        self.bwd = []
        grad_of, written = {}, set()

        def grad(a):  # gradient buffer of a whole forward buffer, and the slice matching `a`
            g = grad_of.get(id(a.buf))
            if g is None:  # fp32 in both modes: gradients are summed over consumers and feed the BatchNorm sums
                g = grad_of[id(a.buf)] = torch.empty(a.buf.shape, dtype=torch.float32, device=dev)
            return Act(g, a.n, a.d, a.h, a.w, a.ld, a.coff, a.c)

        def contribute(a):  # -> (gradient Act, accumulate?)
            acc = id(a.buf) in written
            written.add(id(a.buf))
            return grad(a), acc

        self.grad_act = grad  # forward Act -> its fp32 gradient Act (the tests read the backward pass link by link)
        self.tape = tape
        gfeat = grad(feat)
        written.add(id(feat.buf))
        self.bwd.append(lambda: ops.head_train_bwd(self.dlogits, self.pooled, self.drop, self.w_logits,
                                                   self.grads["logits.conv3d.weight"], self.grads["logits.conv3d.bias"],
                                                   gfeat))
        for rec in reversed(tape):
            if rec[0] == "unit":
                u = rec[1]
                gy = grad(u.y)
                dz = u.z  # in place: the raw output is not needed once xhat has been formed
                self.bwd.append(lambda u=u, gy=gy, dz=dz: ops.bn_train_bwd(
                    gy, u.y, u.z, u.gamma, u.save_mean, u.save_rstd, u.ws, dz, self.grads[u.prefix + ".bn.weight"],
                    self.grads[u.prefix + ".bn.bias"]))
                if u.s2d:  # the stem in mixed precision: from the space-to-depth record its forward read
                    self.bwd.append(lambda u=u, dz=dz: ops.conv3d_wgrad_s2d(
                        u.conv_x, dz, self.grads[u.prefix + ".conv3d.weight"], u.conv_kernel, u.conv_pf))
                else:
                    self.bwd.append(lambda u=u, dz=dz: ops.conv3d_wgrad(
                        u.x, dz, self.grads[u.prefix + ".conv3d.weight"], u.kernel, u.stride, u.pf))
                if u.x is not self.xin:  # the clip needs no gradient
                    gx, acc = contribute(u.x)
                    self.bwd.append(lambda u=u, dz=dz, gx=gx, acc=acc: u.dgrad(dz, gx, acc))
            else:
                _, px, py, am, k, s, pads = rec
                gy = grad(py)
                gx, acc = contribute(px)
                self.bwd.append(lambda gy=gy, am=am, gx=gx, k=k, s=s, pads=pads, acc=acc: ops.maxpool3d_bwd(
                    gy, am, gx, k, s, pads, acc_in=gx if acc else None))

    # ------------------------------------------------------------------------------------------------
    @classmethod
    def from_model(cls, model, batch, clip, **kw):
        """Trainer over a drop-in model's OWN parameter storage (model on the GPU, fp32): `step` updates the model."""
        m = model.module if hasattr(model, "module") else model
        sd = {k: v for k, v in m.state_dict(keep_vars=True).items()}
        dev = next(m.parameters()).device
        return cls(sd, batch, clip, device=dev, share_storage=True, **kw)

    def state_dict(self):
        """Parameters and running statistics as the reference keys them (num_batches_tracked counts the steps).  The
        tensors are the trainer's own (live, updated in place by every step): clone them to keep a snapshot."""
        out = dict(self.params)
        out.update(self.buffers)
        for k in list(self.buffers):
            if k.endswith("running_mean"):
                out[k[:-len("running_mean")] + "num_batches_tracked"] = torch.tensor(self.step_count)
        return out

    @_lib.on_device
    def forward_backward(self, x, target, drop=None, after_forward=None):
        """One forward and backward pass in training mode; returns the loss as a device tensor [1] (no sync).
        x: [B,3,T,H,W] fp32 (host or device), target: class indices [B]; drop: optional scaled dropout mask
        [B, 1024] to use instead of drawing one; after_forward: optional callable run between the passes (the tests
        snapshot the raw convolution outputs there: the backward pass overwrites them with dz)."""
        assert tuple(x.shape) == (self.B, self.C, self.T, self.H, self.W), (tuple(x.shape),)
        self.x.copy_(x, non_blocking=True)
        self.target.copy_(ops.as_int32_targets(target, self.device), non_blocking=True)
        if self.xin_s2d is None:  # layout change only (mask 0): channels-last fp32, or the stem's bf16 space-to-depth record
            ops.perturb_fwd(self.x, self.zero_mask, "freeze", PFMT_NDHWC_F32, self.xin.buf)
        else:
            ops.perturb_fwd(self.x, self.zero_mask, "freeze", PFMT_S2D_BF16, self.xin_s2d.buf)
        for op in self.fwd:
            op()
        if after_forward is not None:
            after_forward()
        d = None
        if drop is not None:
            d = drop.to(device=self.device, dtype=torch.float32).contiguous()
        elif self.drop is not None:
            d = ops.dropout_mask(self.drop, self.dropout_p, self.seed * 1000003 + self.step_count)
        self._drop_used = d
        ops.head_train_fwd(self.feat, d, self.w_logits, self.b_logits, self.target, self.pooled, self.logits,
                           self.dlogits, self.loss)
        saved, self.drop = self.drop, d  # the backward closure reads self.drop
        try:
            for op in self.bwd:
                op()
        finally:
            self.drop = saved
        return self.loss

    @_lib.on_device
    def apply_update(self):
        """optimizer.step(): every parameter in place, then the convolution operands re-packed."""
        self.step_count += 1
        kind = self.opt
        b1 = self.momentum if kind == "sgd" else self.betas[0]
        b2 = 0.0 if kind == "sgd" else self.betas[1]
        if self.world > 1:  # the one exchange of a data-parallel step: NCCL sum of every rank's gradients
            import torch.distributed as dist
            dist.all_reduce(self.flat_grads)
        ops.optim_step_multi(kind, self._table, self.lr, b1, b2, self.eps, self.wd, self.step_count,
                             grad_scale=1.0 / self.world)
        for u in self.units:
            u.repack()
        for t in self._shared:
            torch.autograd.graph.increment_version(t)

    def step(self, x, target, drop=None):
        """pt/train_i3d_smth.py:208-226 for one batch: forward, loss, backward, optimizer step.  Returns the loss."""
        loss = self.forward_backward(x, target, drop)
        self.apply_update()
        return loss
