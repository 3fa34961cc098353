"""GPU (-m gpu): the native ConvLSTM classifier (config C3: KTH, 32 frames of 120x160, 2 layers, 5x5
kernels, conv stride 2, reverse perturbation) against the oracle and the reference's golden vectors:
forward logits, d logit / d mask, Grad-CAM, the drop-in module surface and the mask search."""
import os

import numpy as np
import pytest
import torch

from common import GOLD, quiet, rel_err

pytestmark = pytest.mark.gpu

KW = dict(num_layers=2, kernel=5, conv_stride=2, effective_step=(7, 15, 23, 31))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from interpreting_video_features_b200 import _lib
    _lib.handle()
    return torch.device("cuda")


def build(hid, softmax=False):
    """The model of oracle/pin_against_reference.py §6 (same seed, same BN perturbation)."""
    from interpreting_video_features_b200.pt.models import CLSTM_4
    torch.manual_seed(0)
    m = quiet(CLSTM_4.Model, num_classes=6, nb_lstm_units=hid, channels=3, conv_kernel_size=(5, 5), lstm_layers=2,
              step=32, conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31],
              batch_normalization=True, dropout=0.5, add_softmax=softmax).eval()
    with torch.no_grad():
        m.clstm.bn.running_mean.uniform_(-0.05, 0.05)
        m.clstm.bn.running_var.uniform_(0.5, 1.5)
        m.clstm.bn.weight.uniform_(0.5, 1.5)
        m.clstm.bn.bias.uniform_(-0.1, 0.1)
    return m, {k: v.detach().clone() for k, v in m.state_dict().items()}


def engine(sd, hid, batch, mode, dev, softmax=False):
    from interpreting_video_features_b200.engine_clstm import CLSTMEngine
    return CLSTMEngine(sd, batch, (32, 120, 160), hid, 2, 6, mode=mode, softmax=softmax, device=dev)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("hid", [4, 32])
def test_forward_and_mask_gradient(dev, hid, mode):
    from oracle import clstm_oracle, mask_oracle, synthetic
    g = np.load(os.path.join(GOLD, "clstm_hid%d.npz" % hid))
    _, sd = build(hid)
    x1 = synthetic.clips(1, t=32, h=120, w=160) / 255.0
    x = torch.cat([x1, synthetic.clips(2, t=32, h=120, w=160)[1:] / 255.0])
    masks = torch.stack([torch.from_numpy(g["mask"]), torch.rand(32, generator=torch.Generator().manual_seed(8))])
    eng = engine(sd, hid, 2, mode, dev)
    eng.set_input(x.to(dev))
    eng.set_targets(torch.tensor([2, 4]))
    logits = eng.forward(masks.to(dev), "reverse").clone().cpu()
    dm = eng.backward().clone().cpu()
    tol = 1e-4 if mode == "fp32" else 1e-2
    # clip 0 is the reference's golden case (d logit[2] / d mask under the reverse perturbation)
    assert rel_err(logits[0], g["logits"][0]) < tol, rel_err(logits[0], g["logits"][0])
    if mode == "fp32":
        assert rel_err(dm[0], g["dmask"]) < 2e-3, rel_err(dm[0], g["dmask"])
    for i, tgt in enumerate((2, 4)):
        mi = masks[i].clone().requires_grad_()
        out = clstm_oracle.forward(sd, mask_oracle.perturb_sequence(x[i:i + 1], mi, "reverse"), hidden=hid,
                                   quant=(mode == "bf16"), **KW)
        (gm,) = torch.autograd.grad(out[0, tgt], mi)
        assert rel_err(logits[i], out.detach()[0]) < tol
        # bf16, hid 32: the 2x2 max-pools route the gradient by argmax, and bf16 rounding of the hidden
        # state flips 0.5-0.9 % of those decisions between two evaluations whose activations differ in the
        # last bit (measured with tools/debug_clstm.py: forward activations agree to 0.3 %, the gradient
        # fields to 11-13 % in L2 norm, and on these i.i.d.-noise clips d logit/d mask — a cancelling sum
        # over pixels — to ~10 % against the matched-rounding oracle).  The fp32 mode carries the tight bound.
        gtol = 2e-3 if mode == "fp32" else (3e-2 if hid == 4 else 3.5e-1)  # measured 0.10 / 0.21 on the two clips
        assert rel_err(dm[i], gm) < gtol, (i, rel_err(dm[i], gm))
        assert float(torch.nn.functional.cosine_similarity(dm[i], gm, dim=0)) > 0.95


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_freeze_perturbation_and_softmax(dev, mode):
    from oracle import clstm_oracle, mask_oracle, synthetic
    _, sd = build(4, softmax=True)
    x = synthetic.clips(2, t=32, h=120, w=160) / 255.0
    m = torch.rand((2, 32), generator=torch.Generator().manual_seed(2))
    eng = engine(sd, 4, 2, mode, dev, softmax=True)
    eng.set_input(x.to(dev))
    eng.set_targets(torch.tensor([1, 5]))
    p = eng.forward(m.to(dev), "freeze").clone().cpu()
    dm = eng.backward().clone().cpu()
    for i, tgt in enumerate((1, 5)):
        mi = m[i].clone().requires_grad_()
        out = clstm_oracle.forward(sd, mask_oracle.perturb_sequence(x[i:i + 1], mi, "freeze"), hidden=4, softmax=True,
                                   quant=(mode == "bf16"), **KW)
        (gm,) = torch.autograd.grad(out[0, tgt], mi)
        assert rel_err(p[i], out.detach()[0]) < (1e-4 if mode == "fp32" else 1e-2)
        assert rel_err(dm[i], gm) < (2e-3 if mode == "fp32" else 3e-2), (i, rel_err(dm[i], gm))


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("hid", [4, 32])
def test_gradcam_clstm_dropin(dev, hid, mode):
    from interpreting_video_features_b200.pt.grad_cam_videos import GradCamVideo
    from oracle import gradcam_oracle, synthetic
    g = np.load(os.path.join(GOLD, "clstm_hid%d.npz" % hid))
    model, sd = build(hid)
    model = model.to(dev).set_mode(mode)
    xc = synthetic.clips(1, t=32, h=120, w=160) / 255.0
    gc = GradCamVideo(model=model, target_layer_names=['clstm'], class_dict=None, use_cuda=True,
                      input_spatial_size=(160, 120), normalizePerFrame=True, archType="CLSTM")
    cam, out = gc(xc.to(dev), 2)
    assert cam.shape == (32, 120, 160)
    assert rel_err(out.cpu(), g["cam_output"]) < (1e-4 if mode == "fp32" else 1e-2)
    want, _, low = gradcam_oracle.gradcam_clstm(sd, xc, 2, (160, 120), True, hidden=hid, **KW)
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(cam), np.isnan(want))
    assert np.abs(cam[ok] - want[ok]).max() < (1e-3 if mode == "fp32" else 5e-2)
    if mode == "fp32":
        samp = cam[::8, ::12, ::16]
        gk = ~np.isnan(g["cam_sample"])
        assert np.abs(samp[gk] - g["cam_sample"][gk]).max() < 1e-3


def test_dropin_model_autograd_loop(dev):
    """model(perturb_sequence(x, sigmoid(m), 'reverse')) + stock Adam, as the KTH driver does
    (pt/FindMasksComparison_I3D_KTH.py:250-270), against the oracle loop."""
    from interpreting_video_features_b200.pt import mask
    from oracle import clstm_oracle, mask_oracle, synthetic
    model, sd = build(4, softmax=True)
    model = model.to(dev).set_mode("fp32")
    x = synthetic.clips(1, t=32, h=120, w=160) / 255.0
    raw = torch.tensor([-5.] * 8 + [5.] * 16 + [-5.] * 8)
    tm = raw.clone().to(dev).requires_grad_()
    opt = torch.optim.Adam([tm], lr=0.2)
    losses = []
    for _ in range(3):
        mc = torch.sigmoid(tm)
        loss = 0.02 * torch.sum(torch.abs(mc)) + 0.04 * mask.calc_tv_norm(mc, 3, 3) + \
            model(mask.perturb_sequence(x.to(dev), mc, 'reverse'))[0, 2]
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    o = clstm_oracle.Model(sd, hidden=4, softmax=True, **KW)
    tmo = raw.clone().requires_grad_()
    rec = {}
    final, _ = mask_oracle.mask_search(x, o, 0, [2], tmo, 0.02, 0.04, 3, mask_type="reverse", record=rec)
    np.testing.assert_allclose(losses, rec["loss"], rtol=1e-4)
    np.testing.assert_allclose(torch.sigmoid(tm).detach().cpu().numpy(), final.numpy(), rtol=1e-3, atol=1e-4)


def test_clstm_mask_search_reverse(dev):
    """Config C3: batched mask search on the ConvLSTM with the reverse perturbation, 20 iterations, vs the
    oracle loop (bf16 path vs fp32 reference path: final-mask IoU)."""
    from interpreting_video_features_b200.search import MaskSearch
    from oracle import clstm_oracle, mask_oracle, synthetic
    _, sd = build(32, softmax=True)
    x = synthetic.clips(2, t=32, h=120, w=160) / 255.0
    targets = torch.tensor([2, 4])
    raw0 = torch.tensor([-5.] * 8 + [5.] * 16 + [-5.] * 8).repeat(2, 1)
    eng = engine(sd, 32, 2, "bf16", dev, softmax=True)
    res = MaskSearch(eng, lam1=0.02, lam2=0.04, n_iter=20, perturb="reverse").run(x.to(dev), targets,
                                                                                raw_masks=raw0.to(dev))
    o = clstm_oracle.Model(sd, hidden=32, softmax=True, **KW)
    for i in range(2):
        tm = raw0[i].clone().requires_grad_()
        final, cls = mask_oracle.mask_search(x[i:i + 1], o, 0, [int(targets[i])], tm, 0.02, 0.04, 20,
                                             mask_type="reverse")
        a, b = res["time_mask"][i].cpu() > 0.5, final > 0.5
        union = float((a | b).sum())
        assert union == 0 or float((a & b).sum()) / union >= 0.95
        assert float((res["time_mask"][i].cpu() - final).abs().max()) < 2e-2
        assert abs(float(res["freeze_score"][i]) - cls) < 1e-2 * abs(cls) + 1e-5


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_convlstm_module_forward(dev, mode):
    """models.convolution_lstm.ConvLSTM.forward on its own (pt/models/convolution_lstm.py:96-132): the outputs at the
    effective steps and (x, new_c), as the reference's FeatureExtractor consumes them
    (pt/pytorch-grad-cam/grad-cam.py:33-49), against the oracle."""
    import torch.nn.functional as F
    from oracle import clstm_oracle, synthetic
    model, sd = build(4)
    model = model.to(dev).eval()
    model.clstm.ivf_mode = mode
    x = synthetic.clips(2, kind="square", t=32, h=120, w=160) / 255.0
    with torch.no_grad():
        outs, (last, new_c) = model.clstm(x.to(dev))
        _, want = clstm_oracle.forward(sd, x, hidden=4, return_outputs=True, **KW)
    tol = 1e-4 if mode == "fp32" else 1e-2
    assert len(outs) == 4 and tuple(outs[0].shape) == (2, 4, 7, 10)
    for a, b in zip(outs, want):
        assert rel_err(a.cpu(), b) < tol, rel_err(a.cpu(), b)
    assert torch.equal(last, outs[-1]) and tuple(new_c.shape) == (2, 4, 15, 20)
    with pytest.raises(Exception):
        model.clstm(x.to(dev).requires_grad_())


def test_convlstm_cell_forward_and_init_hidden(dev):
    """ConvLSTMCell.forward / init_hidden on their own (pt/models/convolution_lstm.py:38-60)."""
    from oracle import clstm_oracle
    model, sd = build(4)
    model = model.to(dev).eval()
    cell = model.clstm.cell0
    g = torch.Generator().manual_seed(2)
    x = torch.rand((2, 3, 24, 32), generator=g)
    h0, c0 = cell.init_hidden(batch_size=2, hidden=4, shape=(24, 32))
    assert tuple(h0.shape) == (2, 4, 12, 16) and float(h0.abs().sum()) == 0 and h0.is_cuda
    h = torch.randn((2, 4, 12, 16), generator=g) * 0.5
    c = torch.randn((2, 4, 12, 16), generator=g) * 0.5
    with torch.no_grad():
        h1, c1 = cell(x.to(dev), h.to(dev), c.to(dev))
        wh, wc = clstm_oracle._cell(sd, "clstm.cell0", x, h, c, 5, 2)
    assert rel_err(h1.cpu(), wh) < 1e-4 and rel_err(c1.cpu(), wc) < 1e-4


def test_fused_recurrent_step_equals_unfused(dev, monkeypatch):
    """ivf_conv3d_lstm (h-convolution with the gates, c/h update and the gate activations for the backward pass in
    its epilogue, unit-major channels) against the convolution + gate-kernel pair of round 1 (gate-major): same
    arithmetic on the same fp32 accumulators, so logits and mask gradients agree to rounding; and the unit-major
    weight packing is the gate-major one permuted."""
    from interpreting_video_features_b200 import engine as eng_mod
    from oracle import synthetic
    _, sd = build(32)
    x = synthetic.clips(2, kind="square", t=32, h=120, w=160)
    masks = torch.rand((2, 32), generator=torch.Generator().manual_seed(3))
    res = []
    for fused in ("1", "0"):
        monkeypatch.setenv("IVF_CLSTM_FUSED", fused)
        e = engine(sd, 32, 2, "bf16", dev)
        assert e.unit_major == (fused == "1")
        e.set_input(x.to(dev))
        e.set_targets(torch.tensor([2, 4]))
        logits = e.forward(masks.to(dev), "reverse").clone()
        dm = e.backward().clone()
        res.append((logits, dm, e.layers[0]["wh_f"].clone(), e.layers[0]["c"].clone()))
    # the data-gradient GEMMs reduce over the gate channels in a different order (unit-major K): measured 5e-4
    assert rel_err(res[0][0].cpu(), res[1][0].cpu()) < 1e-5 and rel_err(res[0][1].cpu(), res[1][1].cpu()) < 2e-3
    assert rel_err(res[0][3].cpu(), res[1][3].cpu()) < 1e-6
    um, gm = res[0][2], res[1][2]                      # [4*he, taps, k]
    he = um.shape[0] // 4
    assert torch.equal(um.view(he, 4, *um.shape[1:]).permute(1, 0, 2, 3).reshape(gm.shape), gm)


def test_wavefront_schedule_equals_layer_by_layer(dev, monkeypatch):
    """IVF_CLSTM_WAVE=1 (layers as a wavefront on one stream each, per-step x-convolution / BN+pool / their
    gradients, cross-stream events) computes what the layer-by-layer schedule computes: logits, cell states and
    gradients agree at the bf16 storage level."""
    from oracle import synthetic
    _, sd = build(32)
    x = synthetic.clips(2, kind="square", t=32, h=120, w=160)
    masks = torch.rand((2, 32), generator=torch.Generator().manual_seed(4))
    res = []
    for wave in ("1", "0"):
        monkeypatch.setenv("IVF_CLSTM_WAVE", wave)
        e = engine(sd, 32, 2, "bf16", dev)
        assert e.wave == (wave == "1")
        e.set_input(x.to(dev))
        e.set_targets(torch.tensor([1, 5]))
        for _ in range(2):  # twice: the events and lanes are reused
            logits = e.forward(masks.to(dev), "reverse").clone()
            dm = e.backward().clone()
        torch.cuda.synchronize()
        res.append((logits.cpu(), dm.cpu(), e.layers[0]["c"].clone().cpu(), e.layers[1]["dH"].clone().cpu()))
    # the same launches on the same operands, but the upper layer runs the tile plan with the fewest CTAs: another
    # fp32 summation order, hidden states that round to the other bf16 neighbour now and then (measured 5e-4 on the
    # logits) - bf16 storage level, not 1e-6
    assert rel_err(res[0][0], res[1][0]) < 3e-3 and rel_err(res[0][2], res[1][2]) < 3e-3
    assert rel_err(res[0][3], res[1][3]) < 3e-2 and rel_err(res[0][1], res[1][1]) < 3e-2
