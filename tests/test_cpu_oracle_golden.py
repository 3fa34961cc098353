"""CPU: the oracle restatement against the committed golden vectors (generated from the unmodified
reference by oracle/pin_against_reference.py) and, where /root/reference is present, against the
reference itself."""
import os

import numpy as np
import pytest
import torch

from common import GOLD, i3d_state_dict, quiet
from oracle import clstm_oracle, gradcam_oracle, i3d_oracle, mask_oracle, synthetic


def test_mask_known_answers():
    k = np.load(os.path.join(GOLD, "mask_kats.npz"))
    assert mask_oracle.find_submasks_from_mask(torch.tensor([0, .2, .3, 0, .5, .6, .7, .05, .11])) == [[1, 2], [4, 5, 6], [8]]
    assert mask_oracle.find_submasks_from_mask(torch.tensor([0.1, 0.1000001, 0.1, 0.9])) == [[1], [3]]
    assert mask_oracle.find_submasks_from_mask(torch.zeros(5)) == []
    assert mask_oracle.find_submasks_from_mask(torch.ones(4)) == [[0, 1, 2, 3]]
    s = torch.sigmoid(torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4))
    assert abs(float(mask_oracle.calc_tv_norm(s)) - float(k["tv_sig16"])) < 1e-6
    assert abs(float(k["tv_sig16"]) - 3.841513156890869) < 1e-6
    m8 = torch.tensor([0, .25, .5, 1, 1, .5, .25, 0])
    assert abs(float(mask_oracle.calc_tv_norm(m8)) - 0.59375) < 1e-6
    # closed form 2*sum(d) - d0 - d_{T-2}
    d = (m8[1:] - m8[:-1]).abs() ** 3
    assert abs(float(2 * d.sum() - d[0] - d[-1]) - 0.59375) < 1e-6
    x = torch.arange(16.).reshape(2, 1, 4, 1, 2)
    out = mask_oracle.perturb_sequence(x, torch.tensor([.9, .5, 1, .25]), 'freeze')
    np.testing.assert_allclose(out.numpy(), k["freeze_out"], rtol=0, atol=0)
    np.testing.assert_allclose(out.flatten().numpy(),
                               [0, 1, 1, 2, 1, 2, 4.75, 5.75, 8, 9, 9, 10, 9, 10, 12.75, 13.75])
    xr = torch.tensor([0., 10, 20, 30, 40, 50]).reshape(1, 1, 6, 1, 1)
    r0 = mask_oracle.perturb_sequence(xr, torch.tensor([0, .5, 1, .2, .05, .8]), 'reverse')
    np.testing.assert_allclose(r0.flatten().numpy(), [0, 20, 20, 20, 40, 50])
    r1 = mask_oracle.perturb_sequence(xr, torch.tensor([.6, .5, 1, .2, .3, .05]), 'reverse')
    np.testing.assert_allclose(r1.flatten().numpy(), [24, 20, 20, 20, 16, 50])
    sn = mask_oracle.perturb_sequence(xr, torch.tensor([0, .5, 1, .2, .05, .8]), 'freeze', snap_values=True)
    np.testing.assert_allclose(sn.flatten().numpy(), [0, 10, 10, 30, 40, 40])
    xs = torch.from_numpy(k["rand_x"])
    for mode in ("freeze", "reverse"):
        m = torch.from_numpy(k["rand_%s_mask" % mode]).requires_grad_()
        out = mask_oracle.perturb_sequence(xs, m, mode)
        (g,) = torch.autograd.grad((out * torch.from_numpy(k["rand_%s_gout" % mode])).sum(), m)
        np.testing.assert_allclose(out.detach().numpy(), k["rand_%s_out" % mode], rtol=1e-6, atol=1e-4)
        np.testing.assert_allclose(g.numpy(), k["rand_%s_dmask" % mode], rtol=1e-5, atol=1e-2)


def test_i3d_probs_and_classgrad_against_golden():
    g = np.load(os.path.join(GOLD, "i3d_smth.npz"))
    sd, _ = quiet(i3d_state_dict, 174)
    x2 = synthetic.clips(2)
    with torch.no_grad():
        p = i3d_oracle.forward(sd, x2)
    np.testing.assert_allclose(p.numpy(), g["probs_default"], rtol=2e-4, atol=1e-7)
    sds = i3d_oracle.calibrate_and_sharpen(sd, x2)
    with torch.no_grad():
        ps = i3d_oracle.forward(sds, x2)
    np.testing.assert_allclose(ps.numpy(), g["probs_sharp"], rtol=5e-3, atol=1e-5)
    assert 0.3 < float(ps.max()) < 0.7  # the sharpened head is not degenerate
    tm = torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4, requires_grad=True)
    tgt = int(g["targets"][1])
    out = i3d_oracle.forward(sds, mask_oracle.perturb_sequence(x2, torch.sigmoid(tm), 'freeze'))[1, tgt]
    (gr,) = torch.autograd.grad(out, tm)
    ref = g["classgrad_1"]
    # the fp32 evaluation of this gradient is reproducible to a few % only across thread counts / BLAS
    # blocking (softmax-Jacobian cancellation on the sharpened net; DESIGN.md "Parity")
    assert np.linalg.norm(gr.numpy() - ref) / np.linalg.norm(ref) < 6e-2
    # the class gradient is now comparable with the regulariser's
    assert np.abs(ref).max() > 1e-3


def test_gradcam_lowres_against_golden():
    g = np.load(os.path.join(GOLD, "gradcam_i3d.npz"))
    sd, _ = quiet(i3d_state_dict, 174)
    x2 = synthetic.clips(2)
    sds = i3d_oracle.calibrate_and_sharpen(sd, x2)
    cam, out, low = gradcam_oracle.gradcam_i3d(sds, x2[1:2], None, (224, 224), True)
    np.testing.assert_allclose(out.numpy(), g["output_argmax"], rtol=5e-3, atol=1e-5)
    ref_low = g["cam_lowres_argmax"]
    assert np.linalg.norm(low - ref_low) / np.linalg.norm(ref_low) < 2e-2
    samp = cam[::8, ::16, ::16]
    ok = ~np.isnan(g["cam_sample_argmax"])
    assert np.abs(samp[ok] - g["cam_sample_argmax"][ok]).max() < 2e-2
    assert cam.shape == (16, 224, 224) and np.nanmax(cam) <= 1.0 + 1e-6 and np.nanmin(cam) >= 0.0


def test_resize_bilinear_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(0)
    for (h, w), dsize in [((7, 7), (224, 224)), ((4, 5), (160, 120)), ((7, 10), (160, 120)), ((3, 3), (5, 4))]:
        src = rs.rand(h, w).astype(np.float32)
        np.testing.assert_allclose(gradcam_oracle.resize_bilinear(src, dsize), cv2.resize(src, dsize), rtol=1e-5,
                                   atol=1e-6)


@pytest.mark.parametrize("hid", [4, 32])
def test_clstm_against_golden(hid):
    g = np.load(os.path.join(GOLD, "clstm_hid%d.npz" % hid))
    from interpreting_video_features_b200.pt.models import CLSTM_4
    torch.manual_seed(0)
    m = quiet(CLSTM_4.Model, num_classes=6, nb_lstm_units=hid, channels=3, conv_kernel_size=(5, 5), lstm_layers=2,
              step=32, conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31],
              batch_normalization=True, dropout=0.5).eval()
    with torch.no_grad():
        m.clstm.bn.running_mean.uniform_(-0.05, 0.05)
        m.clstm.bn.running_var.uniform_(0.5, 1.5)
        m.clstm.bn.weight.uniform_(0.5, 1.5)
        m.clstm.bn.bias.uniform_(-0.1, 0.1)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    xc = synthetic.clips(1, t=32, h=120, w=160) / 255.0
    mk = torch.from_numpy(g["mask"]).requires_grad_()
    out = clstm_oracle.forward(sd, mask_oracle.perturb_sequence(xc, mk, 'reverse'), 2, hid)
    (gk,) = torch.autograd.grad(out[0, 2], mk)
    np.testing.assert_allclose(out.detach().numpy(), g["logits"], rtol=1e-4, atol=1e-6)
    assert np.linalg.norm(gk.numpy() - g["dmask"]) / np.linalg.norm(g["dmask"]) < 1e-3


@pytest.mark.skipif(not os.path.isdir("/root/reference/video_features_pytorch"), reason="reference not mounted")
def test_dropin_constructor_matches_reference_weights():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_i3d", "/root/reference/video_features_pytorch/models/I3D_doubled.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    r = ref.Model(174, last_stride=1, stride_mod_layers="", softMax=1)
    sd, m = quiet(i3d_state_dict, 174)
    rsd = r.state_dict()
    assert list(rsd.keys()) == list(sd.keys())
    assert list(r._modules.keys()) == list(m._modules.keys())
    for k in sd:
        assert torch.equal(rsd[k], sd[k]), k


def test_train_step_oracle_against_reference_golden():
    """oracle/train_oracle.py (one training step: BatchNorm with batch statistics, CrossEntropyLoss, every parameter
    gradient, SGD update) reproduces what the UNMODIFIED reference produced for the first step of
    tests/golden/i3d_train.npz (pt/train_i3d_smth.py:208-226; written by oracle/pin_train_step.py)."""
    from oracle import train_oracle
    gold = np.load(os.path.join(GOLD, "i3d_train.npz"))
    sd, _ = quiet(i3d_state_dict, 174)
    x = synthetic.clips(2)
    target = torch.as_tensor(gold["target"])
    loss, logits, grads, buf = train_oracle.loss_and_grads(sd, x, target)
    assert abs(loss - float(gold["loss_1"])) < 1e-5 * abs(float(gold["loss_1"]))
    np.testing.assert_allclose(logits.numpy(), gold["logits_1"], rtol=1e-4, atol=1e-5)
    new, _ = train_oracle.sgd_step(sd, grads, float(gold["lr"]), float(gold["momentum"]), float(gold["weight_decay"]))
    for k, g in grads.items():
        idx = train_oracle.sample_index(k, g.numel())
        # same torch operators in the same order as the reference: equal up to the thread count's summation order
        ref = gold["gsamp_1/" + k]
        assert np.linalg.norm(g.flatten()[idx].numpy() - ref) <= 2e-3 * np.linalg.norm(ref) + 1e-9, k
        np.testing.assert_allclose(new[k].flatten()[idx].numpy(), gold["psamp_1/" + k], rtol=1e-4, atol=1e-6)
    for k, v in buf.items():
        idx = train_oracle.sample_index(k, v.numel())
        np.testing.assert_allclose(v.flatten()[idx].numpy(), gold["psamp_1/" + k], rtol=1e-4, atol=1e-6)
