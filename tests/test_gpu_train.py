"""GPU (-m gpu): the training step (pt/train_i3d_smth.py:192-250) - its kernels one by one
against torch autograd on the CPU, and the whole step (forward, CrossEntropyLoss, backward, optimizer update) against
the oracle restatement and the golden vectors taken from the unmodified reference (oracle/pin_train_step.py).

Tolerances: fp32 arithmetic with different summation orders - 1e-4 on losses / logits / statistics, 1e-3 relative to
the tensor's norm on gradients (a weight gradient sums up to 8e5 products per entry)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import GOLD, i3d_state_dict, quiet, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from interpreting_video_features_b200 import _lib
    _lib.handle()
    return torch.device("cuda")


def to_act(x, dtype, dev, pad_c=0, coff=0):
    """NCDHW cpu tensor -> channels-last Act on the device, optionally as a slice of a wider buffer."""
    from interpreting_video_features_b200.ops import Act
    n, c, d, h, w = x.shape
    ld = c + pad_c
    buf = torch.zeros((n, d, h, w, ld), dtype=dtype)
    buf[..., coff:coff + c] = x.permute(0, 2, 3, 4, 1).to(dtype)
    return Act(buf.to(dev), n, d, h, w, ld, coff, c)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_train_forward_backward(dev, dtype):
    """BatchNorm3d(eps 1e-3, momentum 0.01) with batch statistics + ReLU on a channel slice of a wider buffer:
    output, running statistics, dz, dgamma, dbeta against torch autograd."""
    from interpreting_video_features_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, c, dhw = 3, 40, (2, 5, 7)
    z = (torch.randn((n, c) + dhw, generator=g) * 2 + 0.5).to(dtype).float()
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.2
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    dy = torch.randn((n, c) + dhw, generator=g).to(dtype).float()
    zr, gr, br = z.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y_ref = F.relu(F.batch_norm(zr, rm_ref, rv_ref, gr, br, training=True, momentum=0.01, eps=1e-3))
    dz_ref, dg_ref, db_ref = torch.autograd.grad(y_ref, [zr, gr, br], dy)

    za = to_act(z, dtype, dev, pad_c=8, coff=8)
    ya = to_act(torch.zeros_like(z), dtype, dev, pad_c=24, coff=16)
    dya = to_act(dy, dtype, dev, pad_c=24, coff=16)
    dza = to_act(torch.zeros_like(z), dtype, dev)
    gm, bt, rmd, rvd = gamma.to(dev), beta.to(dev), rm.to(dev), rv.to(dev)
    sm, sr = torch.empty(c, device=dev), torch.empty(c, device=dev)
    ws = torch.empty(2 * c, dtype=torch.float64, device=dev)
    ops.bn_train_fwd(za, gm, bt, 1e-3, 0.01, rmd, rvd, sm, sr, ws, ya, relu=True)
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel_err(ya.ncdhw().cpu(), y_ref.detach()) < tol
    assert float(ya.buf[..., :16].float().abs().max()) == 0.0  # the neighbouring channels are untouched
    assert rel_err(rmd.cpu(), rm_ref) < 1e-5 and rel_err(rvd.cpu(), rv_ref) < 1e-5
    mean = z.transpose(0, 1).reshape(c, -1).mean(1)
    assert rel_err(sm.cpu(), mean) < 1e-5
    # backward with the ReLU mask taken from the forward output
    if dtype == torch.bfloat16:  # use the reference's y as the mask source so that ties round the same way
        ya = to_act(y_ref.detach(), dtype, dev, pad_c=24, coff=16)
    dgd, dbd = torch.empty(c, device=dev), torch.empty(c, device=dev)
    ops.bn_train_bwd(dya, ya, za, gm, sm, sr, ws, dza, dgd, dbd)
    assert rel_err(dza.ncdhw().cpu(), dz_ref) < (1e-4 if dtype == torch.float32 else 1e-2)
    assert rel_err(dgd.cpu(), dg_ref) < 1e-4 and rel_err(dbd.cpu(), db_ref) < 1e-4


CASES = [  # (n, cin, cout, dhw, kernel, stride)
    (2, 3, 64, (8, 20, 22), (7, 7, 7), (2, 2, 2)),    # the stem: narrow input, stride 2, asymmetric pads
    (2, 24, 40, (4, 9, 10), (3, 3, 3), (1, 1, 1)),    # an Inception branch tail
    (2, 72, 136, (3, 6, 5), (1, 1, 1), (1, 1, 1)),    # a bottleneck GEMM, ragged channel tiles
    (1, 16, 8, (2, 7, 9), (1, 3, 3), (1, 2, 2)),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
def test_conv_weight_gradient(dev, case, dtype):
    """ivf_conv3d_wgrad against autograd's weight gradient of the 'same'-padded convolution
    (pt/models/I3D_doubled.py:77-113), input and output gradient as channel slices of wider buffers."""
    from interpreting_video_features_b200 import ops
    from interpreting_video_features_b200.ops import same_pad
    n, cin, cout, dhw, k, s = case
    g = torch.Generator().manual_seed(1)
    x = torch.randn((n, cin) + dhw, generator=g).to(dtype).float()
    w = torch.randn((cout, cin) + k, generator=g, requires_grad=True)
    pads = [same_pad(sz, kk, ss) for sz, kk, ss in zip(dhw, k, s)]
    xp = F.pad(x, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    y = F.conv3d(xp, w, stride=s)
    assert tuple(y.shape[2:]) == tuple(p[2] for p in pads)
    dz = torch.randn(y.shape, generator=g).to(dtype).float()
    (dw_ref,) = torch.autograd.grad(y, w, dz)
    xa = to_act(x, dtype, dev, pad_c=8, coff=0)
    dza = to_act(dz, dtype, dev, pad_c=8, coff=8)
    dw = torch.full(w.shape, 7.0, device=dev)  # overwritten, not accumulated
    ops.conv3d_wgrad(xa, dza, dw, k, s, tuple(p[0] for p in pads))
    assert rel_err(dw.cpu(), dw_ref) < 1e-5


def test_head_train_and_loss(dev):
    """average pool + dropout mask + logits + CrossEntropyLoss(mean) and their gradients
    (pt/models/I3D_doubled.py:360-371, pt/train_i3d_smth.py:124-127)."""
    from interpreting_video_features_b200 import ops
    g = torch.Generator().manual_seed(2)
    b, c, dhw, ncls = 3, 96, (2, 2, 3), 17
    feat = torch.randn((b, c) + dhw, generator=g)
    w = (torch.randn((ncls, c), generator=g) * 0.3).requires_grad_()
    bias = torch.randn(ncls, generator=g).requires_grad_()
    drop = (torch.rand((b, c), generator=g) > 0.5).float() * 2.0
    target = torch.tensor([3, 16, 0])
    fr = feat.clone().requires_grad_()
    pooled = fr.mean(dim=(2, 3, 4)) * drop
    logits = pooled @ w.t() + bias
    loss = F.cross_entropy(logits, target)
    dfeat_ref, dw_ref, db_ref = torch.autograd.grad(loss, [fr, w, bias])
    fa = to_act(feat, torch.float32, dev)
    pd, lg, dl = torch.empty((b, c), device=dev), torch.empty((b, ncls), device=dev), torch.empty((b, ncls), device=dev)
    ls = torch.empty(1, device=dev)
    wd, bd, dd = w.detach().to(dev), bias.detach().to(dev), drop.to(dev)
    ops.head_train_fwd(fa, dd, wd, bd, target.to(dev, torch.int32), pd, lg, dl, ls)
    assert abs(float(ls) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    assert rel_err(lg.cpu(), logits.detach()) < 1e-5
    dwd, dbd, dfa = torch.empty_like(wd), torch.empty_like(bd), to_act(torch.zeros_like(feat), torch.float32, dev)
    ops.head_train_bwd(dl, pd, dd, wd, dwd, dbd, dfa)
    assert rel_err(dwd.cpu(), dw_ref) < 1e-5 and rel_err(dbd.cpu(), db_ref) < 1e-5
    assert rel_err(dfa.ncdhw().cpu(), dfeat_ref) < 1e-5


@pytest.mark.parametrize("kind", ["sgd", "sgd_plain", "adam"])
def test_optimizer_step_equals_torch(dev, kind):
    """ivf_optim_step against torch.optim.SGD(momentum, weight_decay) / Adam(weight_decay) for three steps
    (pt/train_i3d_smth.py:131-138)."""
    from interpreting_video_features_b200 import ops
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) for _ in range(3)]
    pr = p0.clone().requires_grad_()
    if kind == "adam":
        opt = torch.optim.Adam([pr], lr=0.008, weight_decay=1e-5)
    else:
        opt = torch.optim.SGD([pr], lr=0.01, momentum=0.9 if kind == "sgd" else 0.0, weight_decay=1e-4)
    p = p0.to(dev)
    s1, s2 = torch.zeros(1000, device=dev), torch.zeros(1000, device=dev)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        if kind == "adam":
            ops.optim_step("adam", p, gr.to(dev), s1, s2, 0.008, 0.9, 0.999, 1e-8, 1e-5, i + 1)
        else:
            ops.optim_step("sgd", p, gr.to(dev), s1, None, 0.01, 0.9 if kind == "sgd" else 0.0, 0.0, 0.0, 1e-4, i + 1)
        assert rel_err(p.cpu(), pr.detach()) < 1e-6


def test_dropout_mask(dev):
    from interpreting_video_features_b200 import ops
    a, b = torch.empty(1 << 16, device=dev), torch.empty(1 << 16, device=dev)
    ops.dropout_mask(a, 0.5, 11)
    ops.dropout_mask(b, 0.5, 11)
    assert torch.equal(a, b) and set(a.unique().tolist()) == {0.0, 2.0}
    assert abs(float((a == 0).float().mean()) - 0.5) < 0.01
    ops.dropout_mask(b, 0.25, 12)
    assert not torch.equal(a, b) and abs(float((b == 0).float().mean()) - 0.25) < 0.01


# 16x96x96: Mixed_4 maps of 4x6x6, Mixed_5 maps of 2x3x3.  On a 16x64x64 clip the Mixed_4 maps are 4x4x4 and the
# 3x3x3 stride-1 branch pool returns the SAME vector at most positions; the 1x1x1 convolution after it then produces
# exactly equal outputs at neighbouring positions and the following stride-2 pool routes its gradient among exact
# ties - to whichever position an evaluation's last-bit rounding favours (measured there: Mixed_4f.b3b's dz 2.7e-2
# away from the fp64 oracle's choice, everything above it at 2e-5 ... 1.6e-4).  A degenerate geometry, not arithmetic.
SMALL = dict(clip=(16, 96, 96), avg_pool=(2, 3, 3))


def noise_check(mine, theirs, what):
    """mine / theirs: per-tensor distances to the fp64 gradient of the kernels and of an fp32 torch evaluation.
    fp32 on this network is noisy whoever computes it.  (a) Cancellation: the loss reads the AVERAGE-POOLED top
    feature map, so the gradient reaching Mixed_5c is the same number at every position, and every BatchNorm
    backward subtracts the per-channel mean of what it receives.  (b) Decisions: among millions of ReLU / max-pool
    decisions a few sit within rounding of a tie and fall differently in every evaluation; ONE flipped decision moves
    the gradients of its Inception branch by 0.2-0.6 % and everything below it (measured per unit against fp64,
    tools/debug_train.py: torch's fp32 run jumps at Mixed_5c.b1b, the kernels at Mixed_5b.b1b, both sit at 4e-3 from
    Mixed_4d down; at 16x224x224 torch's fp32 gradients are 1.3e-2 (median) / 2.4e-2 (worst) from fp64).
    Both runs are draws of the same noise - and the kernels' draw changes from run to run (atomic summation order) -
    so they are held to it statistically: the median distance at most twice torch's + 5e-3 (one flip near the top of
    the network moves every tensor below it by that much), the worst tensor at most three times torch's worst + 1e-2.
    A wrong term, pad or scale shows up as tens of per cent; every single launch is held to 1e-4 by
    test_train_step_link_by_link."""
    mine, theirs = np.asarray(mine), np.asarray(theirs)
    assert np.median(mine) <= 2.0 * np.median(theirs) + 5e-3, (what, float(np.median(mine)), float(np.median(theirs)))
    assert mine.max() <= 3.0 * theirs.max() + 1e-2, (what, int(np.argmax(mine)), float(mine.max()), float(theirs.max()))


def grad_check(trainer, g64, g32, what):
    keys = list(g64)
    noise_check([rel_err(trainer.grads[k].cpu(), g64[k]) for k in keys], [rel_err(g32[k], g64[k]) for k in keys], what)


@pytest.mark.parametrize("optimizer", ["sgd", "adam"])
def test_train_step_small_vs_oracle(dev, optimizer):
    """Two whole training steps on a small geometry against the oracle: loss, logits, every parameter gradient,
    running statistics and the updated parameters."""
    from interpreting_video_features_b200.train import I3DTrainer
    from oracle import synthetic, train_oracle
    sd, _ = quiet(i3d_state_dict, 174)
    x = synthetic.clips(2, t=SMALL['clip'][0], h=SMALL['clip'][1], w=SMALL['clip'][2])
    target = torch.tensor([5, 77])
    kw = dict(lr=0.01, momentum=0.9, weight_decay=1e-5) if optimizer == "sgd" else dict(lr=0.008, weight_decay=1e-5)
    tr = I3DTrainer(sd, 2, SMALL["clip"], avg_pool=SMALL["avg_pool"], device=dev, optimizer=optimizer, **kw)
    state = None
    for step in (1, 2):
        cur = {k: v.cpu() for k, v in tr.state_dict().items()}  # the trainer's own state: both sides start each step equal
        l64, logits64, g64, buf64 = train_oracle.loss_and_grads(cur, x, target, avg_pool=SMALL["avg_pool"],
                                                                dtype=torch.float64)
        _, _, g32, _ = train_oracle.loss_and_grads(cur, x, target, avg_pool=SMALL["avg_pool"])
        loss = tr.forward_backward(x, target)
        assert abs(float(loss) - l64) < 1e-4 * max(1.0, abs(l64)), (step, float(loss), l64)
        assert rel_err(tr.logits.cpu(), logits64) < 1e-4
        grad_check(tr, g64, g32, "step %d" % step)
        own = {k: v.clone() for k, v in tr.grads.items()}  # the update is checked on the trainer's own gradients
        tr.apply_update()
        if optimizer == "sgd":
            new, state = train_oracle.sgd_step(cur, {k: v.cpu() for k, v in own.items()}, 0.01, 0.9, 1e-5, state)
        else:
            new, state = train_oracle.adam_step(cur, {k: v.cpu() for k, v in own.items()}, 0.008, weight_decay=1e-5,
                                                state=state, step=step)
        out = tr.state_dict()
        for k, v in buf64.items():
            assert rel_err(out[k].cpu(), v) < 1e-4, k
        for k, v in new.items():
            upd = (v - cur[k]).norm() + 1e-12
            assert float((out[k].cpu() - v).norm() / upd) < 1e-3, (k, float((out[k].cpu() - v).norm() / upd))
    assert int(tr.state_dict()["Conv3d_1a_7x7.bn.num_batches_tracked"]) == 2


def test_train_step_full_geometry_vs_reference_golden(dev):
    """Two SGD steps on two 16x224x224 clips against tests/golden/i3d_train.npz (oracle/pin_train_step.py): the
    UNMODIFIED reference's loss, logits, gradients and updated parameters, and the fp64 oracle's gradients of the
    same states.  Gradients are stored as norm + 64 sampled entries per tensor; the kernels must be as close to the
    fp64 values as the reference's own fp32 run is (noise_check)."""
    from interpreting_video_features_b200.train import I3DTrainer
    from oracle import synthetic
    from oracle.train_oracle import sample_index
    gold = np.load(os.path.join(GOLD, "i3d_train.npz"))
    sd, _ = quiet(i3d_state_dict, 174)
    x = synthetic.clips(2)
    target = torch.as_tensor(gold["target"])
    tr = I3DTrainer(sd, 2, (16, 224, 224), device=dev, optimizer="sgd", lr=float(gold["lr"]),
                    momentum=float(gold["momentum"]), weight_decay=float(gold["weight_decay"]))
    step = 1
    loss = tr.forward_backward(x, target)
    ref_loss = float(gold["loss_1"])
    assert abs(float(loss) - ref_loss) < 2e-4 * max(1.0, abs(ref_loss)), (float(loss), ref_loss)
    assert rel_err(tr.logits.cpu(), gold["logits_1"]) < 2e-4
    mine, theirs, mine_n, theirs_n = [], [], [], []
    for k, g in tr.grads.items():
        n64, nref = float(gold["g64norm_%d/%s" % (step, k)]), float(gold["gnorm_%d/%s" % (step, k)])
        mine_n.append(abs(float(g.norm()) - n64) / n64)
        theirs_n.append(abs(nref - n64) / n64)
        s64, sref = gold["g64samp_%d/%s" % (step, k)], gold["gsamp_%d/%s" % (step, k)]
        got = g.flatten()[sample_index(k, g.numel()).to(dev)].cpu().numpy()
        scale = max(np.linalg.norm(s64), n64 / np.sqrt(g.numel()))  # samples of a sparse tensor may all be ~0
        mine.append(np.linalg.norm(got - s64) / scale)
        theirs.append(np.linalg.norm(sref - s64) / scale)
    noise_check(mine, theirs, "sampled entries")
    noise_check(mine_n, theirs_n, "norms")
    tr.apply_update()
    out = tr.state_dict()
    # the state after the update: the CHANGE since initialisation within the gradient's own noise of the
    # reference's change (the update kernel itself is exact: test_optimizer_step_equals_torch and the small case)
    for k in list(tr.params) + list(tr.buffers):
        idx = sample_index(k, out[k].numel())
        ref, init = gold["psamp_1/%s" % k], sd[k].flatten()[idx].numpy()
        got = out[k].flatten()[idx.to(dev)].cpu().numpy()
        assert np.linalg.norm(got - ref) <= 6e-2 * np.linalg.norm(ref - init) + 1e-6 * (np.linalg.norm(ref) + 1.0), \
            (k, float(np.linalg.norm(got - ref)), float(np.linalg.norm(ref - init)))
    # second step from the trainer's OWN updated state (the reference's full state is not stored): the loss fell by
    # 0.91 in one step, so a gradient direction that is 1-2 % noise moves the next loss by ~2 % of that fall
    loss2 = tr.forward_backward(x, target)
    ref1, ref2 = float(gold["loss_1"]), float(gold["loss_2"])
    assert abs(float(loss2) - ref2) < 6e-2 * abs(ref1 - ref2), (float(loss2), ref2)


def test_trainer_updates_the_dropin_model_in_place(dev):
    """I3DTrainer.from_model works on the drop-in model's own parameter storage: after a step the model's
    state dict holds the updated parameters and running statistics, and its eval forward uses them."""
    from interpreting_video_features_b200.train import I3DTrainer
    from oracle import i3d_oracle, synthetic, train_oracle
    sd, model = quiet(i3d_state_dict, 174)
    model = model.to(dev).train().set_mode("fp32")
    model.avg_pool.kernel_size = list(SMALL["avg_pool"])
    x = synthetic.clips(2, t=SMALL['clip'][0], h=SMALL['clip'][1], w=SMALL['clip'][2])
    target = torch.tensor([1, 2])
    tr = I3DTrainer.from_model(model, 2, SMALL["clip"], avg_pool=SMALL["avg_pool"], optimizer="sgd", lr=0.05, momentum=0.0)
    tr.step(x, target)
    _, _, g_or, buf_or = train_oracle.loss_and_grads(sd, x, target, avg_pool=SMALL["avg_pool"], dtype=torch.float64)
    new, _ = train_oracle.sgd_step(sd, g_or, 0.05)
    new = {k: v.float() for k, v in new.items()}
    buf_or = {k: v.float() for k, v in buf_or.items()}
    msd = model.state_dict()
    for k in ("Conv3d_1a_7x7.conv3d.weight", "Mixed_4c.b1b.bn.weight", "logits.conv3d.bias"):
        upd = (new[k] - sd[k]).norm()
        assert float((msd[k].cpu() - new[k]).norm() / upd) < 6e-2, k  # the gradient's own fp32 noise (see noise_check)
    assert rel_err(msd["Mixed_5c.b3b.bn.running_var"].cpu(), buf_or["Mixed_5c.b3b.bn.running_var"]) < 1e-4
    new_sd = dict(sd)
    new_sd.update(new)
    new_sd.update(buf_or)
    with torch.no_grad():
        p_model = model.eval()(x.to(dev))
        p_or = i3d_oracle.forward(new_sd, x, SMALL["avg_pool"])
    assert rel_err(p_model.cpu(), p_or) < 1e-3


def _ncdhw(act):
    return act.tensor().permute(0, 4, 1, 2, 3).float().cpu()


def _conv_same(x, w, stride):
    from oracle.i3d_oracle import _same_pad
    return F.conv3d(_same_pad(x, w.shape[2:], stride), w, stride=stride)


# the second geometry has the odd maps of the KTH configuration (60x80 -> 30x40 -> 15x20 -> 8x10 -> 4x5 -> 2x3):
# asymmetric 'same' pads at every strided pool (pt/models/I3D_doubled.py:9-38)
@pytest.mark.parametrize("geom", [SMALL, dict(clip=(16, 60, 80), avg_pool=(2, 2, 3))], ids=["96x96", "60x80"])
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_train_step_link_by_link(dev, mode, geom):
    """Every launch of one training step against torch evaluated on the step's OWN stored operands (teacher
    forcing).  With bf16 rounding points the step is chaotic end to end - the oracle's matched-rounding restatement
    evaluated in fp32 and in fp64 gives gradients that differ by 0.76 (median, relative) on this geometry, a flipped
    rounding doubles per layer - so the mixed-precision mode cannot be held to any end-to-end gradient; what CAN be
    held exactly is each link: convolution, BatchNorm forward and backward, weight gradient, data gradient (one and
    several consumers, incl. the branch pool's routing).  Tolerances: bf16 storage rounds at 2^-9 = 2e-3."""
    from interpreting_video_features_b200.train import I3DTrainer
    from oracle import synthetic
    from oracle.i3d_oracle import _same_pad
    bf = mode == "bf16"
    tol_store = 6e-3 if bf else 2e-5   # a stored tensor against its fp32 recomputation
    sd, _ = quiet(i3d_state_dict, 174)
    x = synthetic.clips(2, t=geom['clip'][0], h=geom['clip'][1], w=geom['clip'][2])
    target = torch.tensor([5, 77])
    tr = I3DTrainer(sd, 2, geom["clip"], avg_pool=geom["avg_pool"], device=dev, optimizer="sgd", mode=mode)
    zs = {}
    tr.forward_backward(x, target, after_forward=lambda: zs.update({u.prefix: _ncdhw(u.z) for u in tr.units}))
    torch.cuda.synchronize()
    q = (lambda t: t.bfloat16().float()) if bf else (lambda t: t)
    units = {u.prefix: u for u in tr.units}
    checked = 0
    for name, u in units.items():
        w = q(sd[name + ".conv3d.weight"].float())
        xin = q(x) if name == "Conv3d_1a_7x7" else _ncdhw(u.x)
        # forward convolution and BatchNorm(batch statistics) + ReLU
        z_ref = _conv_same(xin, w, u.stride)
        assert rel_err(zs[name], z_ref) < tol_store, (name, "conv", rel_err(zs[name], z_ref))
        zz = zs[name].clone().requires_grad_()
        g, b = sd[name + ".bn.weight"].float().clone().requires_grad_(), sd[name + ".bn.bias"].float().clone().requires_grad_()
        y_ref = F.relu(F.batch_norm(zz, None, None, g, b, training=True, eps=1e-3))
        y = _ncdhw(u.y)
        assert rel_err(y, y_ref.detach()) < tol_store, (name, "bn", rel_err(y, y_ref.detach()))
        # BatchNorm backward from the step's own gradient of y; the ReLU mask is the stored y's
        gy = _ncdhw(tr.grad_act(u.y))
        y_m = F.batch_norm(zz, None, None, g, b, training=True, eps=1e-3) * (y > 0).float()
        dz_ref, dg_ref, db_ref = torch.autograd.grad(y_m, [zz, g, b], gy)
        dz = _ncdhw(u.z)
        assert rel_err(dz, dz_ref) < tol_store, (name, "bn'", rel_err(dz, dz_ref))
        assert rel_err(tr.grads[name + ".bn.weight"].cpu(), dg_ref) < 1e-3, (name, "dgamma")
        assert rel_err(tr.grads[name + ".bn.bias"].cpu(), db_ref) < 1e-3, (name, "dbeta")
        # weight gradient from the stored dz and the unit's input
        xw = xin  # (the stem's tensor-core weight gradient rounds the fp32 clip to bf16 while staging it)
        ww = w.clone().requires_grad_()
        (dw_ref,) = torch.autograd.grad(_conv_same(xw, ww, u.stride), ww, dz)
        assert rel_err(tr.grads[name + ".conv3d.weight"].cpu(), dw_ref) < 1e-4, (name, "wgrad")
        checked += 1
    assert checked == 57
    # data gradients.  One consumer: t1 = b1a's output feeds b1b only.
    for mod in ("Mixed_3b", "Mixed_4d", "Mixed_5c"):
        for a, b_ in (("b1a", "b1b"), ("b2a", "b2b")):
            ub = units["%s.%s" % (mod, b_)]
            w = q(sd["%s.%s.conv3d.weight" % (mod, b_)].float())
            t = _ncdhw(ub.x).requires_grad_()
            (g_ref,) = torch.autograd.grad(_conv_same(t, w, (1, 1, 1)), t, _ncdhw(ub.z))
            got = _ncdhw(tr.grad_act(units["%s.%s" % (mod, a)].y))
            assert rel_err(got, g_ref) < 1e-4, (mod, b_, "dgrad", rel_err(got, g_ref))
        # Four consumers: the module's input gets b0' + b1a' + b2a' + pool'(b3b'), the pool routed by ITS arg-max
        xm = _ncdhw(units[mod + ".b0"].x).requires_grad_()
        total = 0
        for br in ("b0", "b1a", "b2a"):
            w = q(sd["%s.%s.conv3d.weight" % (mod, br)].float())
            total = total + (_conv_same(xm, w, (1, 1, 1)) * _ncdhw(units["%s.%s" % (mod, br)].z)).sum()
        w3 = q(sd[mod + ".b3b.conv3d.weight"].float())
        t3 = F.max_pool3d(_same_pad(xm, (3, 3, 3), (1, 1, 1)), (3, 3, 3), (1, 1, 1))
        total = total + (_conv_same(t3, w3, (1, 1, 1)) * _ncdhw(units[mod + ".b3b"].z)).sum()
        (gx_ref,) = torch.autograd.grad(total, xm)
        got = _ncdhw(tr.grad_act(units[mod + ".b0"].x))
        assert rel_err(got, gx_ref) < 1e-4, (mod, "input gradient", rel_err(got, gx_ref))


def test_mixed_precision_training_tracks_fp32(dev):
    """Five SGD steps on eight 16x224x224 clips: the mixed-precision step (tcgen05 forward convolutions and data
    gradients) lowers the loss like the fp32 step does - measured 5.36 -> 5.07 against 5.34 -> 5.09."""
    from interpreting_video_features_b200.train import I3DTrainer
    from oracle import synthetic
    sd, _ = quiet(i3d_state_dict, 174)
    x = torch.stack([synthetic.uniform_clip(5000 + i) for i in range(8)]).to(dev)
    target = torch.arange(8) % 174
    hist = {}
    for mode in ("fp32", "bf16"):
        tr = I3DTrainer(sd, 8, (16, 224, 224), device=dev, optimizer="sgd", lr=1e-3, momentum=0.9, mode=mode)
        hist[mode] = [float(tr.step(x, target)) for _ in range(5)]
        del tr
    assert hist["fp32"][-1] < hist["fp32"][0] - 0.15 and hist["bf16"][-1] < hist["bf16"][0] - 0.15, hist
    assert abs(hist["bf16"][0] - hist["fp32"][0]) < 0.05 * hist["fp32"][0], hist
    assert abs(hist["bf16"][-1] - hist["fp32"][-1]) < 0.05 * hist["fp32"][-1], hist


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_optimizer_one_launch_equals_per_tensor(dev, kind):
    """ivf_optim_step_multi (every parameter chunk in one launch, gradient scaled by 1 / world first) gives what
    ivf_optim_step gives tensor by tensor."""
    from interpreting_video_features_b200 import ops
    g = torch.Generator().manual_seed(5)
    sizes = [10, 5000, 4096, 12289]
    ps = [torch.randn(n, generator=g).to(dev) for n in sizes]
    gs = [torch.randn(n, generator=g).to(dev) for n in sizes]
    a1, a2 = [torch.zeros(n, device=dev) for n in sizes], [torch.zeros(n, device=dev) for n in sizes]
    qs, b1, b2 = [p.clone() for p in ps], [t.clone() for t in a1], [t.clone() for t in a2]
    table = ops.optim_table(ps, gs, a1, a2 if kind == "adam" else [None] * 4, dev)
    assert table.shape == (1 + 2 + 1 + 4, 5)
    for step in (1, 2, 3):
        ops.optim_step_multi(kind, table, 0.01, 0.9, 0.999, 1e-8, 1e-4, step, grad_scale=0.5)
        for q, gr, s1, s2 in zip(qs, gs, b1, b2):
            ops.optim_step(kind, q, gr * 0.5, s1, s2 if kind == "adam" else None, 0.01, 0.9, 0.999 if kind == "adam" else 0.0,
                           1e-8, 1e-4, step)
        for p, q in zip(ps, qs):
            assert rel_err(p.cpu(), q.cpu()) < 1e-6
