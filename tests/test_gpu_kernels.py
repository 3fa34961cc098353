"""GPU (-m gpu): every libivf kernel through the C ABI against the oracle / plain torch fp32 on the
same seeded inputs.  fp32 kernels: 1e-4; bf16 tensor-core kernels: 1e-2 relative (the north star's
tolerances), bit-exact for index work (max-pool argmax)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import GOLD, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from interpreting_video_features_b200 import _lib
    _lib.handle()  # raises if libivf.so is missing: the CUDA path is the only path
    return torch.device("cuda")


def to_act(x, dtype):
    from interpreting_video_features_b200.ops import Act
    n, c, d, h, w = x.shape
    return Act(x.permute(0, 2, 3, 4, 1).to(dtype).contiguous(), n, d, h, w, c, 0, c)


def ref_conv(x, w, stride, pf, out_dhw):
    """fp32 CPU reference of the generalised conv with explicit front pads and given output size."""
    pads = []
    for size, k, s, p, o in zip(x.shape[2:], w.shape[2:], stride, pf, out_dhw):
        back = (o - 1) * s + k - size - p
        pads.append((p, back))
    (tf, tb), (hf, hb), (wf, wb) = pads
    xp = F.pad(x, (wf, max(wb, 0), hf, max(hb, 0), tf, max(tb, 0)))
    if wb < 0:
        xp = xp[..., :wb]
    if hb < 0:
        xp = xp[..., :hb, :]
    if tb < 0:
        xp = xp[:, :, :tb]
    return F.conv3d(xp, w, stride=stride)


# ----------------------------------------------------------------------------- TMA im2col probe
@pytest.mark.parametrize("cfg", [
    dict(n=2, c=64, dhw=(4, 6, 10), k=(3, 3, 3), pf=(1, 1, 1), m0=0, tap=0),
    dict(n=2, c=64, dhw=(4, 6, 10), k=(3, 3, 3), pf=(1, 1, 1), m0=128, tap=13),
    dict(n=2, c=64, dhw=(4, 6, 10), k=(3, 3, 3), pf=(1, 1, 1), m0=384, tap=26),
    dict(n=3, c=32, dhw=(2, 7, 7), k=(3, 3, 3), pf=(1, 1, 1), m0=128, tap=5),
    dict(n=2, c=16, dhw=(3, 5, 9), k=(1, 1, 1), pf=(0, 0, 0), m0=128, tap=0),
    dict(n=1, c=32, dhw=(4, 8, 8), k=(4, 4, 4), pf=(1, 1, 1), m0=128, tap=63),
    dict(n=1, c=24, dhw=(4, 8, 8), k=(4, 4, 4), pf=(2, 2, 2), m0=0, tap=0),
])
def test_im2col_tma_tile(dev, cfg):
    """What the conv kernel's TMA stages for (tile m0, filter tap) == the im2col definition."""
    from interpreting_video_features_b200 import _lib, ops
    n, c, (d, h, w) = cfg["n"], cfg["c"], cfg["dhw"]
    g = torch.Generator().manual_seed(0)
    x = torch.randint(-8, 9, (n, c, d, h, w), generator=g).float()
    xa = to_act(x.to(dev), torch.bfloat16)
    k, pf = cfg["k"], cfg["pf"]
    tile = ops.probe_im2col(xa, k, (1, 1, 1), pf, (d, h, w), cfg["m0"], cfg["tap"], 0).float().cpu()
    kch = _lib.load().ivf_conv_bf16_kchunk(c)
    kw_i = cfg["tap"] % k[2]
    kh_i = (cfg["tap"] // k[2]) % k[1]
    kd_i = cfg["tap"] // (k[2] * k[1])
    want = torch.zeros(128, kch)
    M = n * d * h * w
    for r in range(128):
        m = cfg["m0"] + r
        if m >= M:
            continue
        ow, t = m % w, m // w
        oh, t = t % h, t // h
        od, nn = t % d, t // d
        zd, zh, zw = od - pf[0] + kd_i, oh - pf[1] + kh_i, ow - pf[2] + kw_i
        if 0 <= zd < d and 0 <= zh < h and 0 <= zw < w:
            want[r, :min(c, kch)] = x[nn, :kch, zd, zh, zw]
    rows_ok = [r for r in range(128) if cfg["m0"] + r < M]
    bad = [r for r in rows_ok if not torch.equal(tile[r], want[r])]
    assert not bad, "im2col rows differ: %s\n got %s\n want %s" % (bad[:8], tile[bad[0]][:8], want[bad[0]][:8])


# ----------------------------------------------------------------------------- convolution
CONV_CASES = [
    # n, cin, cout, dhw, kernel, stride, note
    (2, 64, 64, (3, 9, 10), (1, 1, 1), (1, 1, 1)),
    (1, 64, 192, (4, 12, 12), (3, 3, 3), (1, 1, 1)),
    (2, 16, 48, (2, 7, 7), (3, 3, 3), (1, 1, 1)),
    (2, 24, 64, (2, 7, 7), (3, 3, 3), (1, 1, 1)),
    (1, 96, 208, (4, 14, 14), (3, 3, 3), (1, 1, 1)),
    (1, 112, 288, (2, 7, 7), (3, 3, 3), (1, 1, 1)),
    (2, 832, 384, (2, 7, 7), (1, 1, 1), (1, 1, 1)),
    (1, 480, 16, (4, 14, 14), (1, 1, 1), (1, 1, 1)),
]


# wide stride-1 maps: the bf16 path serves these with the halo-slab kernel (csrc/conv_slab.cu); the cases
# cover 1/2/3/4 accumulators per tile, a ragged last row tile, a 32-channel tail chunk (96 = 64 + 32), a
# 16-channel operand, two N tiles (cout 192), the 4x4x4 stem geometry and a 2-D 5x5 (ConvLSTM) kernel
SLAB_CASES = [
    (1, 64, 192, (3, 10, 56), (3, 3, 3), (1, 1, 1)),
    (2, 96, 128, (2, 9, 28), (3, 3, 3), (1, 1, 1)),
    (1, 64, 32, (2, 30, 30), (3, 3, 3), (1, 1, 1)),
    (1, 16, 16, (2, 24, 24), (3, 3, 3), (1, 1, 1)),
    (1, 32, 64, (3, 6, 40), (4, 4, 4), (1, 1, 1)),
    (3, 32, 128, (1, 30, 40), (1, 5, 5), (1, 1, 1)),
    (1, 128, 96, (1, 5, 120), (3, 3, 3), (1, 1, 1)),
]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES + SLAB_CASES)
def test_conv_forward_bn_relu(dev, mode, case):
    from interpreting_video_features_b200 import _lib, engine, ops
    from interpreting_video_features_b200.ops import Act, same_pad
    n, cin, cout, dhw, k, s = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn((n, cin) + dhw, generator=g)
    w = torch.randn((cout, cin) + k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    if mode == "bf16":
        x, w = x.bfloat16().float(), w.bfloat16().float()
    geo = [same_pad(sz, kk, ss) for sz, kk, ss in zip(dhw, k, s)]
    pf, out_dhw = tuple(g_[0] for g_ in geo), tuple(g_[2] for g_ in geo)
    ref = F.relu(ref_conv(x, w, s, pf, out_dhw) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    xa = to_act(x.to(dev), dt)
    # write into a channel slice of a wider buffer (the Inception concat path)
    wide = Act.empty(n, *out_dhw, cout + 16, dt, dev, zero=True)
    out = wide.slice(8, cout)
    ops.conv3d(xa, engine.pack_fwd(w.to(dev), mode), out, k, s, pf, flags=_lib.EP_RELU, scale=scale.to(dev),
               shift=shift.to(dev))
    torch.cuda.synchronize()
    got = out.ncdhw().cpu()
    tol = 1e-2 if mode == "bf16" else 1e-4
    assert rel_err(got, ref) < tol, rel_err(got, ref)
    # neighbours of the slice untouched
    assert float(wide.tensor()[..., :8].float().abs().max()) == 0.0
    assert float(wide.tensor()[..., 8 + cout:].float().abs().max()) == 0.0


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES[:6] + SLAB_CASES)
def test_conv_data_gradient_with_fused_mask_and_accumulate(dev, mode, case):
    """dX = (acc + conv_dgrad(dZ)) * 1[y_prev>0] * scale_prev — what the backward pass launches."""
    from interpreting_video_features_b200 import engine, ops
    from interpreting_video_features_b200.ops import same_pad
    n, cin, cout, dhw, k, s = case
    g = torch.Generator().manual_seed(7 + hash(case) % 1000)
    x = torch.randn((n, cin) + dhw, generator=g, requires_grad=True)
    w = torch.randn((cout, cin) + k, generator=g) / (cout * k[0] * k[1] * k[2]) ** 0.5
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    if mode == "bf16":
        w = w.bfloat16().float()
    geo = [same_pad(sz, kk, ss) for sz, kk, ss in zip(dhw, k, s)]
    pf, out_dhw = tuple(g_[0] for g_ in geo), tuple(g_[2] for g_ in geo)
    y = ref_conv(x, w, s, pf, out_dhw)
    dz = torch.randn(y.shape, generator=g)
    if mode == "bf16":
        dz = dz.bfloat16().float()
    (gx,) = torch.autograd.grad(y, x, dz)
    y_prev = torch.randn(x.shape, generator=g)
    acc = torch.randn(x.shape, generator=g) * 0.1
    sc_prev = torch.rand(cin, generator=g) + 0.5
    ref = (gx + acc) * (y_prev > 0).float() * sc_prev.view(1, -1, 1, 1, 1)
    dza = to_act(dz.to(dev), dt)
    ya = to_act(y_prev.to(dev), dt)
    acc_a = to_act(acc.to(dev), torch.float32)
    gxa = to_act(torch.zeros_like(y_prev).to(dev), dt)
    wd = engine.pack_dgrad(w.to(dev), mode)
    if mode == "fp32":
        ops.conv3d(dza, wd, gxa, k, s, pf, acc_in=acc_a, mask=ya, mask_scale=sc_prev.to(dev), transposed=1)
    else:
        pfd = tuple(kk - 1 - p for kk, p in zip(k, pf))
        ops.conv3d(dza, wd, gxa, k, (1, 1, 1), pfd, acc_in=acc_a, mask=ya, mask_scale=sc_prev.to(dev))
    torch.cuda.synchronize()
    tol = 1e-2 if mode == "bf16" else 1e-4
    assert rel_err(gxa.ncdhw().cpu(), ref) < tol
    if mode == "bf16" and case in SLAB_CASES:
        # the same data gradient with two output depths stacked along the MMA's N where the layer allows it (a
        # request the kernel may decline: odd depth, no front padding, N not in whole TMEM blocks)
        gxa.buf.zero_()
        ops.conv3d(dza, wd, gxa, k, (1, 1, 1), pfd, acc_in=acc_a, mask=ya, mask_scale=sc_prev.to(dev),
                   plan=(0, 0, 0, 0, 0, 2))
        assert rel_err(gxa.ncdhw().cpu(), ref) < tol


@pytest.mark.parametrize("case", [(1, 64, 64, (3, 10, 56), (3, 3, 3)), (2, 32, 32, (2, 9, 28), (3, 3, 3)),
                                  (1, 32, 64, (3, 6, 40), (4, 4, 4)), (1, 96, 16, (2, 14, 14), (3, 3, 3)),
                                  (1, 32, 64, (4, 6, 40), (4, 4, 4)), (2, 64, 96, (6, 7, 14), (3, 3, 3))])
def test_conv_slab_every_tile_plan(dev, case):
    """Every tile plan the halo-slab kernel accepts for a layer (kw-merge 1..4, 1..4 accumulators per tile, single
    and double buffered TMEM, single CTAs and cta_group::2 pairs, 1..N tiles, one or two stacked output depths)
    gives the same convolution: each within 1e-2 of the fp32 reference (bf16 operands) and within 2e-3 of the cost
    model's own plan."""
    from interpreting_video_features_b200 import _lib, engine, ops, tune
    from interpreting_video_features_b200.ops import Act, same_pad
    n, cin, cout, dhw, k = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn((n, cin) + dhw, generator=g).bfloat16().float()
    w = (torch.randn((cout, cin) + k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5).bfloat16().float()
    geo = [same_pad(sz, kk, 1) for sz, kk in zip(dhw, k)]
    pf = tuple(q[0] for q in geo)
    ref = ref_conv(x, w, (1, 1, 1), pf, dhw)
    xa = to_act(x.to(dev), torch.bfloat16)
    wp = engine.pack_fwd(w.to(dev), "bf16")
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    out = Act.empty(n, *dhw, cout, torch.bfloat16, dev)

    def make_desc(plan):
        return ops.conv_desc(xa, out, k, (1, 1, 1), pf, plan=plan)

    plans = tune.candidates(make_desc, sm)
    assert len(plans) >= 6, plans
    assert {p[3] for p in plans} == {1, 2} and len({p[0] for p in plans}) >= 2 and len({p[2] for p in plans}) == 2
    if dhw[0] % 2 == 0 and cout % 32 == 0:  # depth stacking: even depth, TMEM blocks of whole 32 columns
        assert {p[5] for p in plans} == {1, 2}, plans
        assert {p[3] for p in plans if p[5] == 2} == {1, 2}, plans  # stacked, single CTAs and pairs
    ops.conv3d(xa, wp, out, k, (1, 1, 1), pf)
    base = out.ncdhw().cpu()
    assert rel_err(base, ref) < 1e-2
    for plan in plans:
        out.buf.zero_()
        ops.conv3d(xa, wp, out, k, (1, 1, 1), pf, plan=plan)
        got = out.ncdhw().cpu()
        assert rel_err(got, ref) < 1e-2, (plan, rel_err(got, ref))
        assert rel_err(got, base) < 2e-3, (plan, rel_err(got, base))


@pytest.mark.parametrize("case", [((4, 14, 14), 96, 208, 24, 64), ((8, 28, 28), 96, 128, 16, 32),
                                  ((2, 7, 7), 160, 320, 32, 128)])
def test_conv3d_pair_grouped_launch(dev, case):
    """ivf_conv3d_pair: the two 3x3x3 branches of an Inception module as ONE grouped launch of the halo-slab kernel
    (CTAs split between the two problems) give what two separate convolutions give - forward with BN+ReLU into
    channel slices of one concat buffer, and the data gradients with the fused ReLU'/BN' mask."""
    from interpreting_video_features_b200 import _lib, engine, ops
    from interpreting_video_features_b200.ops import Act
    dhw, ci1, co1, ci2, co2 = case
    n, k, pf = 2, (3, 3, 3), (1, 1, 1)
    g = torch.Generator().manual_seed(3)
    before = _lib.launch_count(dev)

    def mk(cin, cout):
        x = torch.randn((n, cin) + dhw, generator=g).bfloat16().float()
        w = (torch.randn((cout, cin) + k, generator=g) / (cin * 27) ** 0.5).bfloat16().float()
        sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
        return x, w, sc, sh

    xa, wa, sca, sha = mk(ci1, co1)
    xb, wb, scb, shb = mk(ci2, co2)
    concat = Act.empty(n, *dhw, co1 + co2 + 16, torch.bfloat16, dev, zero=True)
    oa, ob = concat.slice(8, co1), concat.slice(8 + co1, co2)
    argsf = [dict(x=to_act(x.to(dev), torch.bfloat16), w=engine.pack_fwd(w.to(dev), "bf16"), out=o, kernel=k,
                  stride=(1, 1, 1), pad_front=pf, flags=_lib.EP_RELU, scale=sc.to(dev), shift=sh.to(dev))
             for x, w, sc, sh, o in ((xa, wa, sca, sha, oa), (xb, wb, scb, shb, ob))]
    n0 = _lib.launch_count(dev)
    ops.conv3d_pair(*argsf)
    assert _lib.launch_count(dev) - n0 == 1, "the pair must go out as one grouped launch"
    for x, w, sc, sh, o in ((xa, wa, sca, sha, oa), (xb, wb, scb, shb, ob)):
        ref = F.relu(ref_conv(x, w, (1, 1, 1), pf, dhw) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))
        assert rel_err(o.ncdhw().cpu(), ref) < 1e-2
    assert float(concat.tensor()[..., :8].float().abs().max()) == 0.0
    assert float(concat.tensor()[..., 8 + co1 + co2:].float().abs().max()) == 0.0
    # data gradients with the producers' ReLU'/BN' masks, side by side in one buffer (g_t12 of the engine)
    gt = Act.empty(n, *dhw, ci1 + ci2, torch.bfloat16, dev, zero=True)
    argsb, refs = [], []
    for (x, w, cin, cout, off) in ((xa, wa, ci1, co1, 0), (xb, wb, ci2, co2, ci1)):
        dz = torch.randn((n, cout) + dhw, generator=g).bfloat16().float()
        xr = x.clone().requires_grad_()
        (gx,) = torch.autograd.grad(ref_conv(xr, w, (1, 1, 1), pf, dhw), xr, dz)
        msc = torch.rand(cin, generator=g) + 0.5
        refs.append((gx * (x > 0).float() * msc.view(1, -1, 1, 1, 1), off, cin))
        argsb.append(dict(x=to_act(dz.to(dev), torch.bfloat16), w=engine.pack_dgrad(w.to(dev), "bf16"),
                          out=gt.slice(off, cin), kernel=k, stride=(1, 1, 1), pad_front=pf,
                          mask=to_act(x.to(dev), torch.bfloat16), mask_scale=msc.to(dev)))
    n0 = _lib.launch_count(dev)
    ops.conv3d_pair(*argsb)
    assert _lib.launch_count(dev) - n0 == 1
    full = gt.ncdhw().cpu()
    for ref, off, cin in refs:
        assert rel_err(full[:, off:off + cin], ref) < 1e-2
    assert before >= 0


def test_conv_fp32_strided_stem_and_its_gradient(dev):
    """The 7x7x7 stride-2 stem in fp32 mode (true strided gather, transposed data-gradient)."""
    from interpreting_video_features_b200 import _lib, engine, ops
    from interpreting_video_features_b200.ops import Act, same_pad
    g = torch.Generator().manual_seed(3)
    x = torch.randn((1, 3, 8, 20, 18), generator=g, requires_grad=True)
    w = torch.randn((16, 3, 7, 7, 7), generator=g) * 0.05
    geo = [same_pad(sz, 7, 2) for sz in x.shape[2:]]
    pf, od = tuple(q[0] for q in geo), tuple(q[2] for q in geo)
    y = ref_conv(x, w, (2, 2, 2), pf, od)
    dz = torch.randn(y.shape, generator=g)
    (gx,) = torch.autograd.grad(y, x, dz)
    xa = to_act(x.detach().to(dev), torch.float32)
    out = Act.empty(1, *od, 16, torch.float32, dev)
    ops.conv3d(xa, engine.pack_fwd(w.to(dev), "fp32"), out, (7, 7, 7), (2, 2, 2), pf)
    assert rel_err(out.ncdhw().cpu(), y.detach()) < 1e-4
    gxa = Act.empty(1, 8, 20, 18, 3, torch.float32, dev)
    ops.conv3d(to_act(dz.to(dev), torch.float32), engine.pack_dgrad(w.to(dev), "fp32"), gxa, (7, 7, 7), (2, 2, 2),
               pf, transposed=1)
    assert rel_err(gxa.ncdhw().cpu(), gx) < 1e-4


@pytest.mark.parametrize("hw", [(24, 20), (20, 64)])
def test_conv_bf16_space_to_depth_stem(dev, hw):
    """bf16 stem: stride-2 7x7x7 presented as a stride-1 4x4x4 conv over the space-to-depth clip,
    forward and data gradient, against the true strided convolution (narrow map: im2col kernel; 64-wide
    clip -> 32-wide operand: halo-slab kernel, 24-of-32-channel operand)."""
    from interpreting_video_features_b200 import _lib, engine, ops
    from interpreting_video_features_b200.ops import Act
    g = torch.Generator().manual_seed(5)
    b, t = 2, 8
    h, w_ = hw
    x = (torch.rand((b, 3, t, h, w_), generator=g) * 255).bfloat16().float()
    w = (torch.randn((64, 3, 7, 7, 7), generator=g) * 0.01).bfloat16().float()
    xr = x.clone().requires_grad_()
    y = F.conv3d(F.pad(xr, (2, 3, 2, 3, 2, 3)), w, stride=2)
    dz = torch.randn(y.shape, generator=g).bfloat16().float()
    (gx,) = torch.autograd.grad(y, xr, dz)
    # clip -> s2d operand via the perturb kernel with a zero mask (P == x)
    xin = Act.empty(b, t // 2, h // 2, w_ // 2, 32, torch.bfloat16, dev, zero=True)
    ops.perturb_fwd(x.to(dev), torch.zeros(t, device=dev), "freeze", _lib.PFMT_S2D_BF16, xin.buf)
    sd = {"u.conv3d.weight": w}
    unit = engine.Unit(sd, "u", (2, 2, 2), "bf16", dev, s2d=True)
    out = Act.empty(b, t // 2, h // 2, w_ // 2, 64, torch.bfloat16, dev)
    ops.conv3d(xin.slice(0, 24), unit.w_fwd, out, (4, 4, 4), (1, 1, 1), (1, 1, 1))
    assert rel_err(out.ncdhw().cpu(), y.detach()) < 1e-2
    gin = Act.empty(b, t // 2, h // 2, w_ // 2, 32, torch.bfloat16, dev, zero=True)
    ops.conv3d(to_act(dz.to(dev), torch.bfloat16), unit.w_dgrad, gin.slice(0, 24), (4, 4, 4), (1, 1, 1), (2, 2, 2))
    gs = gin.tensor()[..., :24].float().cpu().view(b, t // 2, h // 2, w_ // 2, 2, 2, 2, 3)
    got = gs.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(b, 3, t, h, w_)
    assert rel_err(got, gx) < 1e-2


@pytest.mark.parametrize("case", [(192, (64, 96 + 16), (2, 6, 7)), (480, (192, 96 + 16), (3, 5, 5)),
                                  (528, (112, 144 + 32), (1, 14, 14)), (832, (384, 192 + 48), (2, 7, 7)),
                                  (528, (160, 112 + 24), (2, 9, 9))])
def test_conv1x1_two_destinations_and_two_sources(dev, case):
    """ivf_conv3d_split, as the Inception modules use it: forward b0|b1a|b2a in one GEMM with b0's channels going
    to a slice of the concat buffer and the bottlenecks to their own buffer; backward one data-gradient GEMM
    over [dz of b0 (a slice of the concat gradient) | dz of the bottlenecks] with the consumer sum and the
    ReLU'/BN' mask.  160 and 112 are not multiples of the 64-channel K stage (zero-filled tail)."""
    from interpreting_video_features_b200 import engine, ops
    from interpreting_video_features_b200.ops import Act
    cin, (c0, c12), dhw = case
    n = 2
    g = torch.Generator().manual_seed(3)
    x = (torch.randn((n, cin) + dhw, generator=g)).bfloat16().float()
    w0 = torch.randn((c0, cin, 1, 1, 1), generator=g) * 0.05
    w12 = torch.randn((c12, cin, 1, 1, 1), generator=g) * 0.05
    wall = torch.cat([w0, w12])
    scale = torch.rand(c0 + c12, generator=g) + 0.5
    shift = torch.randn(c0 + c12, generator=g) * 0.1
    wb = wall.bfloat16().float()
    y = F.relu(F.conv3d(x, wb) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    xa = to_act(x.to(dev), torch.bfloat16)
    ctot = c0 + 40  # concat buffer: b0's channels first, others behind
    concat = Act.empty(n, *dhw, ctot, torch.bfloat16, dev, zero=True)
    t12 = Act.empty(n, *dhw, c12, torch.bfloat16, dev)
    ops.conv1x1_split(xa, engine.pack_fwd(wall.to(dev), "bf16"), concat.slice(0, c0), out2=t12,
                      flags=ops.EP_RELU, scale=scale.to(dev), shift=shift.to(dev))
    assert rel_err(concat.ncdhw().cpu()[:, :c0], y[:, :c0]) < 1e-2
    assert rel_err(t12.ncdhw().cpu(), y[:, c0:]) < 1e-2
    assert float(concat.ncdhw()[:, c0:].abs().max()) == 0.0, "channels behind b0's slice must stay untouched"
    # backward: g_x = mask(acc + W0^T dz0 + W12^T dz12)
    dz0 = torch.randn((n, c0) + dhw, generator=g).bfloat16().float()
    dz12 = torch.randn((n, c12) + dhw, generator=g).bfloat16().float()
    acc = torch.randn((n, cin) + dhw, generator=g)
    mscale = torch.rand(cin, generator=g) + 0.5
    want = (F.conv_transpose3d(dz0, w0.bfloat16().float()) + F.conv_transpose3d(dz12, w12.bfloat16().float()) + acc)
    want = torch.where(x > 0, want * mscale.view(1, -1, 1, 1, 1), torch.zeros_like(want))
    gconcat = Act.empty(n, *dhw, ctot, torch.bfloat16, dev, zero=True)
    gconcat.tensor()[..., :c0] = dz0.permute(0, 2, 3, 4, 1).to(dev).bfloat16()
    gconcat.tensor()[..., c0:] = 7.0  # the other branches' gradients: must not leak into the reduction
    wd = engine.pack_dgrad_two_sources(w0.to(dev), w12.to(dev))
    gx = xa.like()
    ops.conv1x1_split(gconcat.slice(0, c0), wd, gx, x2=to_act(dz12.to(dev), torch.bfloat16),
                      acc_in=to_act(acc.to(dev), torch.float32), mask=xa, mask_scale=mscale.to(dev))
    assert rel_err(gx.ncdhw().cpu(), want) < 1e-2


# ----------------------------------------------------------------------------- max-pool
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", [((1, 3, 3), (1, 2, 2), (4, 13, 12), 64), ((3, 3, 3), (2, 2, 2), (5, 9, 10), 40),
                                  ((2, 2, 2), (2, 2, 2), (4, 7, 6), 24), ((3, 3, 3), (1, 1, 1), (2, 7, 7), 48),
                                  ((3, 3, 3), (1, 1, 1), (3, 5, 4), 6)])
@pytest.mark.parametrize("nonneg", [False, True])
def test_maxpool_same_zero_padding(dev, dt, case, nonneg):
    """nonneg: ReLU-output-like input with the IVF_POOL_NONNEG promise (what the engine passes) - many exact zeros
    and repeated values, so that ties between real elements and with the zero padding decide the routing."""
    from interpreting_video_features_b200 import ops
    from interpreting_video_features_b200.ops import Act, same_pad
    from oracle import i3d_oracle
    k, s, dhw, c = case
    g = torch.Generator().manual_seed(11)
    # mix of negatives and zeros so that padded zeros and ties take part (post-ReLU maps are like this)
    x = torch.randn((2, c) + dhw, generator=g)
    x = torch.where(torch.rand(x.shape, generator=g) < 0.4, torch.zeros_like(x), x)
    if nonneg:
        x = torch.round(torch.relu(x) * 4) / 4  # few distinct values: ties between real elements
    x = x.to(dt).float().requires_grad_()
    y = i3d_oracle.maxpool_same(x, k, s)
    gy = torch.randn(y.shape, generator=g).to(dt).float()
    (gx,) = torch.autograd.grad(y, x, gy)
    geo = [same_pad(sz, kk, ss) for sz, kk, ss in zip(dhw, k, s)]
    pf, od = tuple(q[0] for q in geo), tuple(q[2] for q in geo)
    xa = to_act(x.detach().to(dev), dt)
    out = Act.empty(2, *od, c, dt, dev)
    am = torch.empty((out.pixels, c), dtype=torch.uint8, device=dev)
    ops.maxpool3d_fwd(xa, out, am, k, s, pf, nonneg=nonneg)
    assert torch.equal(out.ncdhw().cpu(), y.detach()), "max-pool forward must be exact"
    gxa = xa.like(zero=True)
    ops.maxpool3d_bwd(to_act(gy.to(dev), dt), am, gxa, k, s, pf)
    torch.testing.assert_close(gxa.ncdhw().cpu(), gx.to(dt).float(), rtol=2e-2 if dt == torch.bfloat16 else 1e-6,
                               atol=2e-2 if dt == torch.bfloat16 else 1e-6)


@pytest.mark.parametrize("flags", ["plain", "accum_mask_f32", "mask_bf16"])
@pytest.mark.parametrize("case", [((3, 3, 3), (1, 1, 1), (8, 28, 28), 32),    # branch pool, several row tiles
                                  ((3, 3, 3), (1, 1, 1), (4, 14, 14), 24),    # 8-channel blocks (c % 16 != 0)
                                  ((1, 3, 3), (1, 2, 2), (3, 38, 34), 64),    # stem pool: one depth per tile
                                  ((3, 3, 3), (2, 2, 2), (8, 28, 28), 16),    # Mixed_3c -> 4b pool
                                  ((3, 3, 3), (2, 2, 2), (5, 9, 11), 16),     # odd sizes, stride 2
                                  ((2, 2, 2), (2, 2, 2), (4, 14, 14), 32)])   # Mixed_4f -> 5b pool
def test_maxpool_backward_epilogues_bf16(dev, case, flags):
    """The shared-memory scatter backward (and its gather fallback) with the fused epilogues the engine uses:
    consumer-sum accumulation into an fp32 gradient and the ReLU'/BN' mask of the producing unit."""
    from interpreting_video_features_b200 import ops
    from interpreting_video_features_b200.ops import Act, same_pad
    from oracle import i3d_oracle
    k, s, dhw, c = case
    g = torch.Generator().manual_seed(5)
    x = torch.randn((2, c) + dhw, generator=g)
    x = torch.where(torch.rand(x.shape, generator=g) < 0.4, torch.zeros_like(x), x)
    x = x.bfloat16().float().requires_grad_()
    y = i3d_oracle.maxpool_same(x, k, s)
    gy = torch.randn(y.shape, generator=g).bfloat16().float()
    (gx,) = torch.autograd.grad(y, x, gy)
    geo = [same_pad(sz, kk, ss) for sz, kk, ss in zip(dhw, k, s)]
    pf, od = tuple(q[0] for q in geo), tuple(q[2] for q in geo)
    xa = to_act(x.detach().to(dev), torch.bfloat16)
    out = Act.empty(2, *od, c, torch.bfloat16, dev)
    am = torch.empty((out.pixels, c), dtype=torch.uint8, device=dev)
    ops.maxpool3d_fwd(xa, out, am, k, s, pf)
    assert torch.equal(out.ncdhw().cpu(), y.detach()), "max-pool forward must be exact"
    dy = to_act(gy.to(dev), torch.bfloat16)
    scale = torch.rand(c, generator=g) + 0.5
    sc5 = scale.view(1, c, 1, 1, 1)
    if flags == "plain":
        dx = xa.like(zero=True)
        ops.maxpool3d_bwd(dy, am, dx, k, s, pf)
        want = gx.bfloat16().float()
        tol = 1e-2
    elif flags == "accum_mask_f32":
        acc = torch.randn(x.shape, generator=g)
        dx = to_act(acc.to(dev), torch.float32)          # in place: acc_in == dx, as the engine does
        ops.maxpool3d_bwd(dy, am, dx, k, s, pf, acc_in=dx, mask=xa, mask_scale=scale.to(dev))
        want = torch.where(x.detach() > 0, (gx + acc) * sc5, torch.zeros_like(gx))
        tol = 1e-6
    else:
        dx = xa.like(zero=True)
        ops.maxpool3d_bwd(dy, am, dx, k, s, pf, mask=xa, mask_scale=scale.to(dev))
        want = torch.where(x.detach() > 0, gx * sc5, torch.zeros_like(gx)).bfloat16().float()
        tol = 1e-2
    torch.testing.assert_close(dx.ncdhw().float().cpu(), want, rtol=tol, atol=tol)
    if flags != "plain" and s[1] == 2:
        # the same through the ReLU' bit mask the forward kernel writes (ivf_maxpool3d_fwd_bits / _bwd_bits)
        bits = torch.zeros((xa.pixels, c // 8), dtype=torch.uint8, device=dev)
        out2 = Act.empty(2, *od, c, torch.bfloat16, dev)
        am2 = torch.empty_like(am)
        ops.maxpool3d_fwd(xa, out2, am2, k, s, pf, relu_bits=bits)
        assert torch.equal(am2, am) and torch.equal(out2.buf, out.buf)
        xcl = x.detach().permute(0, 2, 3, 4, 1).reshape(-1, c // 8, 8)
        want_bits = ((xcl > 0).to(torch.int32) << torch.arange(8)).sum(-1).to(torch.uint8)
        assert torch.equal(bits.cpu(), want_bits), "ReLU bit mask of the pool input"
        poison = Act(torch.full_like(xa.buf, float("nan")), xa.n, xa.d, xa.h, xa.w, xa.ld, xa.coff, xa.c)
        if flags == "accum_mask_f32":
            dx2 = to_act(acc.to(dev), torch.float32)
            ops.maxpool3d_bwd(dy, am, dx2, k, s, pf, acc_in=dx2, mask=poison, mask_scale=scale.to(dev), relu_bits=bits)
        else:
            dx2 = xa.like(zero=True)
            ops.maxpool3d_bwd(dy, am, dx2, k, s, pf, mask=poison, mask_scale=scale.to(dev), relu_bits=bits)
        assert torch.equal(dx2.buf, dx.buf), "bit-mask backward == bf16-mask backward (mask tensor not read)"


# ----------------------------------------------------------------------------- head
@pytest.mark.parametrize("softmax", [True, False])
def test_head_forward_backward(dev, softmax):
    from interpreting_video_features_b200 import ops
    from interpreting_video_features_b200.ops import Act
    g = torch.Generator().manual_seed(2)
    feat = torch.rand((3, 1024, 2, 7, 7), generator=g, requires_grad=True)
    w = torch.randn((174, 1024), generator=g) * 0.2
    b = torch.randn(174, generator=g) * 0.1
    logits = F.linear(feat.mean(dim=(2, 3, 4)), w, b)
    out = F.softmax(logits, dim=1) if softmax else logits
    dout = torch.randn(out.shape, generator=g)
    (gf,) = torch.autograd.grad(out, feat, dout)
    fa = to_act(feat.detach().to(dev), torch.float32)
    o = torch.empty((3, 174), device=dev)
    ops.head_fwd(fa, w.to(dev), b.to(dev), softmax, o)
    assert rel_err(o.cpu(), out.detach()) < 1e-5
    dfa = fa.like()
    ops.head_bwd(dfa, w.to(dev), softmax, o, dout.to(dev))
    assert rel_err(dfa.ncdhw().cpu(), gf) < 1e-4


@pytest.mark.parametrize("case", [(1024, (2, 7, 7), 174, torch.bfloat16), (832, (1, 3, 5), 6, torch.bfloat16),
                                  (99, (1, 2, 3), 11, torch.float32), (70, (1, 1, 1), 101, torch.bfloat16),
                                  (1568, (1, 1, 1), 6, torch.float32)])
def test_head_shapes_dtypes_and_mask(dev, case):
    """Channel counts that are not a multiple of the 64-channel block (and odd ones: scalar loads), bf16
    features, p == 1 (the ConvLSTM Linear) and the fused ReLU'/BN' mask with an fp32 gradient output."""
    from interpreting_video_features_b200 import ops
    c, dhw, ncls, dt = case
    g = torch.Generator().manual_seed(7)
    feat = (torch.rand((3, c) + dhw, generator=g) - 0.3).to(dt).float().requires_grad_()
    w = torch.randn((ncls, c), generator=g) * 0.2
    b = torch.randn(ncls, generator=g) * 0.1
    out = F.softmax(F.linear(feat.mean(dim=(2, 3, 4)), w, b), dim=1)
    dout = torch.randn(out.shape, generator=g)
    (gf,) = torch.autograd.grad(out, feat, dout)
    fa = to_act(feat.detach().to(dev), dt)
    o = torch.empty((3, ncls), device=dev)
    ops.head_fwd(fa, w.to(dev), b.to(dev), True, o)
    assert rel_err(o.cpu(), out.detach()) < 1e-5
    scale = torch.rand(c, generator=g) + 0.5
    dfa = fa.like(dtype=torch.float32)
    ops.head_bwd(dfa, w.to(dev), True, o, dout.to(dev), mask=fa, mask_scale=scale.to(dev))
    want = torch.where(feat.detach() > 0, gf * scale.view(1, c, 1, 1, 1), torch.zeros_like(gf))
    assert rel_err(dfa.ncdhw().cpu(), want) < 1e-4
    if dt == torch.bfloat16:
        dfb = fa.like()
        ops.head_bwd(dfb, w.to(dev), True, o, dout.to(dev))
        assert rel_err(dfb.ncdhw().float().cpu(), gf) < 1e-2


# ----------------------------------------------------------------------------- perturb
@pytest.mark.parametrize("mode", ["freeze", "reverse"])
def test_perturb_known_answers_and_random(dev, mode):
    from interpreting_video_features_b200 import _lib, ops
    from oracle import mask_oracle
    k = np.load(os.path.join(GOLD, "mask_kats.npz"))
    xs = torch.from_numpy(k["rand_x"])
    m = torch.from_numpy(k["rand_%s_mask" % mode])
    out = torch.empty_like(xs, device=dev)
    ops.perturb_fwd(xs.to(dev), m.to(dev), mode, _lib.PFMT_NCDHW_F32, out)
    np.testing.assert_allclose(out.cpu().numpy(), k["rand_%s_out" % mode], rtol=1e-6, atol=1e-4)
    dm = torch.empty((2, 12), device=dev)
    ops.perturb_bwd(xs.to(dev), m.to(dev), mode, _lib.PFMT_NCDHW_F32, torch.from_numpy(k["rand_%s_gout" % mode]).to(dev), dm)
    np.testing.assert_allclose(dm.sum(0).cpu().numpy(), k["rand_%s_dmask" % mode], rtol=1e-4, atol=1e-2)
    # RNG-free known answers of SURVEY §4.3
    if mode == "freeze":
        x = torch.arange(16.).reshape(2, 1, 4, 1, 2)
        o = torch.empty_like(x, device=dev)
        ops.perturb_fwd(x.to(dev), torch.tensor([.9, .5, 1, .25], device=dev), mode, _lib.PFMT_NCDHW_F32, o)
        assert o.flatten().tolist() == [0, 1, 1, 2, 1, 2, 4.75, 5.75, 8, 9, 9, 10, 9, 10, 12.75, 13.75]
    else:
        xr = torch.tensor([0., 10, 20, 30, 40, 50]).reshape(1, 1, 6, 1, 1)
        for mr, want in (([0, .5, 1, .2, .05, .8], [0, 20, 20, 20, 40, 50]),
                         ([.6, .5, 1, .2, .3, .05], [24, 20, 20, 20, 16, 50])):
            o = torch.empty_like(xr, device=dev)
            ops.perturb_fwd(xr.to(dev), torch.tensor(mr, device=dev), mode, _lib.PFMT_NCDHW_F32, o)
            np.testing.assert_allclose(o.flatten().cpu().numpy(), want, rtol=1e-6)


@pytest.mark.parametrize("mode", ["freeze", "reverse"])
@pytest.mark.parametrize("fmt", ["ndhwc", "s2d_bf16", "s2d_f32grad"])
def test_perturb_formats_per_clip_masks(dev, mode, fmt):
    """Per-clip masks [B,T], the NDHWC fp32 and space-to-depth bf16 layouts, forward and dmask."""
    from interpreting_video_features_b200 import _lib, ops
    from oracle import mask_oracle
    g = torch.Generator().manual_seed(9)
    b, t, h, w = 3, 16, 12, 20
    x = torch.rand((b, 3, t, h, w), generator=g) * 255
    masks = torch.rand((b, t), generator=g)
    gout = torch.randn((b, 3, t, h, w), generator=g)
    ref_out, ref_dm = [], []
    for i in range(b):
        mi = masks[i].clone().requires_grad_()
        o = mask_oracle.perturb_sequence(x[i:i + 1], mi, mode)
        (gm,) = torch.autograd.grad((o * gout[i:i + 1]).sum(), mi)
        ref_out.append(o.detach())
        ref_dm.append(gm)
    ref_out, ref_dm = torch.cat(ref_out), torch.stack(ref_dm)
    xd, md = x.to(dev), masks.to(dev)
    dm = torch.empty((b, t), device=dev)
    if fmt == "ndhwc":
        out = torch.empty((b, t, h, w, 3), device=dev)
        ops.perturb_fwd(xd, md, mode, _lib.PFMT_NDHWC_F32, out)
        torch.testing.assert_close(out.permute(0, 4, 1, 2, 3).cpu(), ref_out, rtol=1e-6, atol=1e-4)
        ops.perturb_bwd(xd, md, mode, _lib.PFMT_NDHWC_F32, gout.permute(0, 2, 3, 4, 1).contiguous().to(dev), dm)
        assert rel_err(dm.cpu(), ref_dm) < 1e-4
    else:
        out = torch.zeros((b, t // 2, h // 2, w // 2, 32), dtype=torch.bfloat16, device=dev)
        ops.perturb_fwd(xd, md, mode, _lib.PFMT_S2D_BF16, out)
        o = out[..., :24].float().cpu().view(b, t // 2, h // 2, w // 2, 2, 2, 2, 3)
        o = o.permute(0, 7, 1, 4, 2, 5, 3, 6).reshape(b, 3, t, h, w)
        assert rel_err(o, ref_out) < 4e-3  # bf16 rounding of the stored operand
        assert float(out[..., 24:].float().abs().max()) == 0.0
        gs = gout.view(b, 3, t // 2, 2, h // 2, 2, w // 2, 2).permute(0, 2, 4, 6, 3, 5, 7, 1).reshape(
            b, t // 2, h // 2, w // 2, 24)
        gs = torch.cat([gs, torch.zeros(b, t // 2, h // 2, w // 2, 8)], dim=-1).contiguous()
        if fmt == "s2d_bf16":
            gsd = gs.bfloat16()
            tol = 1e-2
        else:
            gsd, tol = gs, 1e-4
        ops.perturb_bwd(xd, md, mode, _lib.PFMT_S2D_BF16, gsd.to(dev), dm)
        assert rel_err(dm.cpu(), ref_dm) < tol


# ----------------------------------------------------------------------------- loss + Adam
def test_mask_loss_adam_matches_torch(dev):
    from interpreting_video_features_b200 import ops
    from oracle import mask_oracle
    g = torch.Generator().manual_seed(4)
    nclip, t = 5, 16
    m0 = torch.randn((nclip, t), generator=g) * 3
    dclass = torch.randn((20, nclip, t), generator=g) * 1e-3
    lam1, lam2 = 0.01, 0.02
    ref_m = []
    for c in range(nclip):
        p = m0[c].clone().requires_grad_()
        opt = torch.optim.Adam([p], lr=0.2)
        for it in range(20):
            s = torch.sigmoid(p)
            loss = lam1 * s.abs().sum() + lam2 * mask_oracle.calc_tv_norm(s, 3, 3) + (s * dclass[it, c]).sum()
            opt.zero_grad()
            loss.backward()
            opt.step()
        ref_m.append(p.detach())
    ref_m = torch.stack(ref_m)
    m = m0.clone().to(dev)
    ea, es = torch.zeros_like(m), torch.zeros_like(m)
    step = torch.zeros(nclip, dtype=torch.int32, device=dev)
    sig = torch.empty_like(m)
    losses = torch.empty((nclip, 3), device=dev)
    for it in range(20):
        ops.mask_loss_adam(m, ea, es, dclass[it].to(dev), 0, lam1, lam2, losses=losses, sig_out=sig, step_dev=step)
    assert step.tolist() == [20] * nclip
    torch.testing.assert_close(m.cpu(), ref_m, rtol=1e-3, atol=2e-3)
    torch.testing.assert_close(sig.cpu(), torch.sigmoid(ref_m), rtol=1e-3, atol=1e-3)


def test_tv_norm_general_pq_and_nan_on_constant_mask(dev):
    from interpreting_video_features_b200 import ops
    from oracle import mask_oracle
    g = torch.Generator().manual_seed(6)
    for p, q in ((3, 3), (2, 3), (3, 1)):
        m = torch.rand(16, generator=g, requires_grad=True)
        v = mask_oracle.calc_tv_norm(m, p, q)
        (gm,) = torch.autograd.grad(v, m)
        val = torch.empty(1, device=dev)
        dm = torch.empty(16, device=dev)
        ops.tv_norm(m.detach().to(dev), p, q, val, dm)
        assert abs(float(val) - float(v)) < 1e-4 * max(1.0, abs(float(v)))
        assert rel_err(dm.cpu(), gm) < 1e-3
    m = torch.full((16,), 0.3, requires_grad=True)
    (gm,) = torch.autograd.grad(mask_oracle.calc_tv_norm(m, 3, 3), m)
    val = torch.empty(1, device=dev)
    dm = torch.empty(16, device=dev)
    ops.tv_norm(m.detach().to(dev), 3, 3, val, dm)
    assert float(val) == 0.0 and bool(torch.isnan(gm).all()) and bool(torch.isnan(dm).all())


# ----------------------------------------------------------------------------- Grad-CAM tail
@pytest.mark.parametrize("per_frame", [True, False])
@pytest.mark.parametrize("geom", [((2, 7, 7), 16, (224, 224)), ((4, 4, 5), 32, (160, 120))])
def test_gradcam_fused_kernel(dev, per_frame, geom):
    from interpreting_video_features_b200 import ops
    from oracle import gradcam_oracle
    (tp, hp, wp), clip, size = geom
    g = torch.Generator().manual_seed(8)
    c = 1024
    act = torch.rand((2, c, tp, hp, wp), generator=g)
    grad = torch.randn((2, c, tp, hp, wp), generator=g) * 1e-3
    cam = torch.empty((2, clip, size[1], size[0]), device=dev)
    low = torch.empty((2, tp, hp, wp), device=dev)
    ops.gradcam(to_act(act.to(dev), torch.float32), to_act(grad.to(dev), torch.float32), clip // tp, size[1], size[0],
                per_frame, cam, low)
    for i in range(2):
        want, want_low = gradcam_oracle.cam_from_features(act[i].numpy(), grad[i:i + 1].numpy(), clip, size, per_frame)
        np.testing.assert_allclose(low[i].cpu().numpy(), want_low, rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(cam[i].cpu().numpy(), want, rtol=1e-3, atol=2e-5)


# ----------------------------------------------------------------------------- model-load kernels (pack.cu)
def test_pack_weights_kernel_equals_torch_packers(dev):
    """ivf_pack_weights (engine.pack) against the torch packers of tests/packing_ref.py, bit for bit: forward and
    data-gradient operands in both dtypes, the space-to-depth stem, the padded ConvLSTM gate stack, two sources."""
    import packing_ref as pr
    from interpreting_video_features_b200 import _lib, engine
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)

    def pads(n, k):
        return lib.ivf_conv_bf16_cout_pad(n), lib.ivf_conv_bf16_cin_pad(k)

    for shape in [(64, 24, 3, 3, 3), (208, 96, 3, 3, 3), (16, 480, 1, 1, 1), (128, 32, 1, 5, 5)]:
        w = torch.randn(shape, generator=g)
        co, ci = shape[:2]
        assert torch.equal(engine.pack_fwd(w.to(dev), "bf16").cpu(), pr.pack_fwd(w, "bf16", *pads(co, ci)))
        assert torch.equal(engine.pack_dgrad(w.to(dev), "bf16").cpu(), pr.pack_dgrad(w, "bf16", *pads(ci, co)))
        assert torch.equal(engine.pack_fwd(w.to(dev), "fp32").cpu(), pr.pack_fwd(w, "fp32"))
        assert torch.equal(engine.pack_dgrad(w.to(dev), "fp32").cpu(), pr.pack_dgrad(w, "fp32"))
    ws = torch.randn((64, 3, 7, 7, 7), generator=g)
    w2 = pr.s2d_weight(ws, 4)
    assert torch.equal(engine.pack([ws.to(dev)], "bf16", s2d=(2, 2, 2)).cpu(), pr.pack_fwd(w2, "bf16", *pads(64, 24)))
    assert torch.equal(engine.pack([ws.to(dev)], "bf16", dgrad=True, s2d=(2, 2, 2)).cpu(),
                       pr.pack_dgrad(w2, "bf16", *pads(24, 64)))
    # ConvLSTM gates: hidden 4 padded to 8; layer 0 (3 channels, record of 16) and layer 1 (4 -> 8 channels)
    for cin, cin_eff, total in ((3, 3, 16), (4, 8, None)):
        gates = [torch.randn((4, cin, 5, 5), generator=g) for _ in range(4)]
        wx = pr.s2d_weight_2d(pr.pad_gates(gates, 8, cin_eff)[:, :, 0], 3)
        if total:
            wx = torch.cat([wx, wx.new_zeros(32, total - wx.shape[1], 3, 3)], dim=1)
        wx = wx.unsqueeze(2)
        kw = dict(s2d=(1, 2, 2), ci_stride=cin_eff, ceff_total=total, co_offs=[0, 8, 16, 24], co_total=32)
        gd = [x.to(dev) for x in gates]
        assert torch.equal(engine.pack(gd, "bf16", **kw).cpu(), pr.pack_fwd(wx, "bf16", *pads(32, wx.shape[1])))
        assert torch.equal(engine.pack(gd, "bf16", dgrad=True, **kw).cpu(),
                           pr.pack_dgrad(wx, "bf16", *pads(wx.shape[1], 32)))
        wf = pr.pad_gates(gates, 8, cin_eff)
        kw = dict(ci_stride=cin_eff, co_offs=[0, 8, 16, 24], co_total=32)
        assert torch.equal(engine.pack(gd, "fp32", **kw).cpu(), pr.pack_fwd(wf, "fp32"))
        assert torch.equal(engine.pack(gd, "fp32", dgrad=True, **kw).cpu(), pr.pack_dgrad(wf, "fp32"))
    w0, w12 = torch.randn((96, 192, 1, 1, 1), generator=g), torch.randn((112, 192, 1, 1, 1), generator=g)
    assert torch.equal(engine.pack_dgrad_two_sources(w0.to(dev), w12.to(dev)).cpu(),
                       pr.pack_dgrad_two_sources(w0, w12, *pads(192, 128 + 112)))


def test_bn_fold_fill_one_hot_argmax(dev):
    from interpreting_video_features_b200 import engine, ops
    g = torch.Generator().manual_seed(4)
    c = 70
    sd = {"u.bn.weight": torch.rand(c, generator=g) + 0.5, "u.bn.bias": torch.randn(c, generator=g),
          "u.bn.running_mean": torch.randn(c, generator=g), "u.bn.running_var": torch.rand(c, generator=g) + 0.1,
          "u.conv3d.bias": torch.randn(c, generator=g)}
    scale, shift = torch.empty(c, device=dev), torch.empty(c, device=dev)
    keep = engine.bn_fold(sd, "u", dev, 1e-3, scale, shift)
    want_s = sd["u.bn.weight"] / torch.sqrt(sd["u.bn.running_var"] + 1e-3)
    want_b = sd["u.bn.bias"] - sd["u.bn.running_mean"] * want_s
    torch.testing.assert_close(scale.cpu(), want_s, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(shift.cpu(), want_b, rtol=1e-6, atol=1e-6)
    keep = engine.bn_fold(sd, "u", dev, 1e-3, scale, shift, with_conv_bias=True)
    torch.testing.assert_close(shift.cpu(), want_b + want_s * sd["u.conv3d.bias"], rtol=1e-6, atol=1e-6)
    keep = engine.bn_fold({"u.conv3d.bias": sd["u.conv3d.bias"]}, "u", dev, 1e-3, scale, shift, with_conv_bias=True)
    assert torch.equal(scale.cpu(), torch.ones(c)) and torch.equal(shift.cpu(), sd["u.conv3d.bias"])
    del keep
    z = ops.zeros((3, 5, 7), torch.float32, dev)
    assert torch.equal(z.cpu(), torch.zeros(3, 5, 7))
    zb = ops.zeros((1, 2, 3, 4, 8), torch.bfloat16, dev)
    assert float(zb.float().abs().sum()) == 0.0
    tg = torch.tensor([3, 0, 173])
    oh = torch.empty((3, 174), device=dev)
    ops.one_hot(ops.as_int32_targets(tg, dev), oh)
    assert torch.equal(oh.cpu(), F.one_hot(tg, 174).float())
    x = torch.randn((5, 174), generator=g)
    x[1, 7] = x[1, 90] = 9.0            # tie: first maximum
    x[2, 40] = float("nan")             # NaN wins, like np.argmax
    am = torch.empty(5, dtype=torch.int32, device=dev)
    ops.argmax_rows(x.to(dev), am)
    assert am.cpu().tolist() == np.argmax(x.numpy(), axis=1).tolist()
