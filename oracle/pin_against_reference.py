"""Pin the oracle against the UNMODIFIED reference and (re)generate tests/golden/*.npz.

Runs only where /root/reference exists (the authoring container; CPU, fp32).  For every piece of
the hot path it asserts oracle == reference on seeded inputs, then stores the reference's outputs
as golden vectors for the tests that run without the reference (GPU box).  Usage:
    python oracle/pin_against_reference.py            # check + write tests/golden
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/video_features_pytorch"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "pytorch-grad-cam"))

from oracle import clstm_oracle, gradcam_oracle, i3d_oracle, mask_oracle, synthetic  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def close(a, b, tol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), what + ": NaN pattern differs"
    a, b = a[~np.isnan(a)], b[~np.isnan(b)]
    err = np.max(np.abs(a - b) / (np.abs(b) + 1e-12)) if a.size else 0.0
    abs_err = np.max(np.abs(a - b)) if a.size else 0.0
    ok = (err <= tol) or (abs_err <= tol * 1e-3)
    print("%-58s rel %.2e abs %.2e %s" % (what, err, abs_err, "ok" if ok else "MISMATCH"))
    assert ok, what


def main():
    import mask as ref_mask  # pt/mask.py
    from models import CLSTM_4, I3D_doubled, I3D_doubled_kth
    from grad_cam_videos import GradCamVideo
    import cv2

    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())

    # ---------------------------------------------------------------- 1. mask KATs (RNG free)
    kat = {}
    m1 = torch.tensor([0, .2, .3, 0, .5, .6, .7, .05, .11])
    assert ref_mask.find_submasks_from_mask(m1) == mask_oracle.find_submasks_from_mask(m1) == [[1, 2], [4, 5, 6], [8]]
    m2 = torch.tensor([0.1, 0.1000001, 0.1, 0.9])
    assert ref_mask.find_submasks_from_mask(m2) == mask_oracle.find_submasks_from_mask(m2) == [[1], [3]]
    s = torch.sigmoid(torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4))
    kat["tv_sig16"] = float(ref_mask.calc_tv_norm(s))
    close(mask_oracle.calc_tv_norm(s), kat["tv_sig16"], 1e-6, "calc_tv_norm sigmoid mask")
    m8 = torch.tensor([0, .25, .5, 1, 1, .5, .25, 0])
    kat["tv_m8"] = float(ref_mask.calc_tv_norm(m8))
    close(mask_oracle.calc_tv_norm(m8), kat["tv_m8"], 1e-6, "calc_tv_norm ramp mask")
    x = torch.arange(16.).reshape(2, 1, 4, 1, 2)
    mf = torch.tensor([.9, .5, 1, .25])
    kat["freeze_out"] = ref_mask.perturb_sequence(x, mf, 'freeze').numpy()
    close(mask_oracle.perturb_sequence(x, mf, 'freeze'), kat["freeze_out"], 1e-7, "freeze KAT")
    xr = torch.tensor([0., 10, 20, 30, 40, 50]).reshape(1, 1, 6, 1, 1)
    for i, mr in enumerate([[0, .5, 1, .2, .05, .8], [.6, .5, 1, .2, .3, .05]]):
        mr = torch.tensor(mr)
        kat["reverse_out%d" % i] = ref_mask.perturb_sequence(xr, mr, 'reverse').numpy()
        close(mask_oracle.perturb_sequence(xr, mr, 'reverse'), kat["reverse_out%d" % i], 1e-7, "reverse KAT %d" % i)
    ms = torch.tensor([0, .5, 1, .2, .05, .8])
    kat["snap_out"] = ref_mask.perturb_sequence(xr, ms.clone(), 'freeze', snap_values=True).numpy()
    close(mask_oracle.perturb_sequence(xr, ms.clone(), 'freeze', snap_values=True), kat["snap_out"], 1e-7, "snap KAT")
    # seeded random masks, both modes, with gradients
    g = torch.Generator().manual_seed(7)
    xs = torch.rand((2, 3, 12, 5, 6), generator=g) * 255
    for mode in ("freeze", "reverse"):
        mm = torch.rand(12, generator=g).requires_grad_()
        gout = torch.rand(xs.shape, generator=g)
        ref_out = ref_mask.perturb_sequence(xs, mm, mode)
        (ref_g,) = torch.autograd.grad((ref_out * gout).sum(), mm)
        mo = mm.detach().clone().requires_grad_()
        o_out = mask_oracle.perturb_sequence(xs, mo, mode)
        (o_g,) = torch.autograd.grad((o_out * gout).sum(), mo)
        close(o_out.detach(), ref_out.detach(), 1e-6, "perturb %s random: value" % mode)
        close(o_g, ref_g, 1e-5, "perturb %s random: dmask" % mode)
        kat["rand_%s_mask" % mode] = mm.detach().numpy()
        kat["rand_%s_gout" % mode] = gout.numpy()
        kat["rand_%s_out" % mode] = ref_out.detach().numpy()
        kat["rand_%s_dmask" % mode] = ref_g.numpy()
    kat["rand_x"] = xs.numpy()
    np.savez_compressed(os.path.join(GOLD, "mask_kats.npz"), **kat)

    # ---------------------------------------------------------------- 2. cv2 bilinear
    src = np.random.RandomState(0).rand(7, 7).astype(np.float32)
    for dsize in [(224, 224), (160, 120), (13, 9)]:
        close(gradcam_oracle.resize_bilinear(src, dsize), cv2.resize(src, dsize), 1e-5, "resize_bilinear %s" % (dsize,))
    src2 = np.random.RandomState(1).rand(4, 5).astype(np.float32)
    close(gradcam_oracle.resize_bilinear(src2, (160, 120)), cv2.resize(src2, (160, 120)), 1e-5, "resize_bilinear 4x5")

    # ---------------------------------------------------------------- 3. I3D smth: forward, class gradient, 3 iterations
    torch.manual_seed(0)
    ref = quiet(I3D_doubled.Model, 174, last_stride=1, stride_mod_layers="", softMax=1).eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    x2 = synthetic.clips(2)  # [2,3,16,224,224]
    with torch.no_grad():
        p_ref = ref(x2)
        p_or = i3d_oracle.forward(sd, x2)
    close(p_or, p_ref, 1e-5, "I3D smth forward probs (default init)")
    gold = {"probs_default": p_ref.numpy()}

    # sharpened init, applied to both through the shared state dict
    sds = i3d_oracle.calibrate_and_sharpen(sd, x2)
    ref.load_state_dict(sds)
    tm = torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4)
    with torch.no_grad():
        ps_ref = ref(x2)
    close(i3d_oracle.forward(sds, x2).detach(), ps_ref, 1e-5, "I3D smth forward probs (sharpened)")
    gold["probs_sharp"] = ps_ref.numpy()
    # targets = the predicted class of each clip: a class with ~1e-30 probability has no gradient
    target = ps_ref.argmax(dim=1).tolist()
    gold["targets"] = np.array(target, dtype=np.int64)
    for bi in (0, 1):
        tmr = tm.clone().requires_grad_()
        out = ref(ref_mask.perturb_sequence(x2, torch.sigmoid(tmr), 'freeze'))[bi, target[bi]]
        (g_ref,) = torch.autograd.grad(out, tmr)
        tmo = tm.clone().requires_grad_()
        out_o = i3d_oracle.forward(sds, mask_oracle.perturb_sequence(x2, torch.sigmoid(tmo), 'freeze'))[bi, target[bi]]
        (g_or,) = torch.autograd.grad(out_o, tmo)
        close(g_or, g_ref, 1e-4, "I3D smth class-gradient d p/d raw-mask, clip %d (|g|max %.1e)" % (bi, g_ref.abs().max()))
        gold["classgrad_%d" % bi] = g_ref.numpy()
        gold["classprob_%d" % bi] = np.float32(out.item())
    # three reference iterations (SURVEY §4.3 last bullet) on the sharpened model, clip 0
    rec_or = {}
    tmr = tm.clone().requires_grad_()
    opt = torch.optim.Adam([tmr], lr=0.2)
    losses = []
    for _ in range(3):
        mc = torch.sigmoid(tmr)
        loss = 0.01 * mc.abs().sum() + 0.02 * ref_mask.calc_tv_norm(mc) + \
            ref(ref_mask.perturb_sequence(x2, mc, 'freeze'))[0, target[0]]
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    tmo = tm.clone().requires_grad_()
    fm, _ = mask_oracle.mask_search(x2, i3d_oracle.Model(sds), 0, target, tmo, 0.01, 0.02, 3, record=rec_or)
    close(rec_or["loss"], losses, 1e-5, "3 mask-search iterations: losses")
    close(fm, torch.sigmoid(tmr).detach(), 1e-5, "3 mask-search iterations: sigmoid(mask)")
    gold["iter3_losses"] = np.array(losses, dtype=np.float32)
    gold["iter3_mask"] = torch.sigmoid(tmr).detach().numpy()
    np.savez_compressed(os.path.join(GOLD, "i3d_smth.npz"), **gold)

    # ---------------------------------------------------------------- 4. Grad-CAM I3D (reference GradCamVideo on CPU)
    gc = GradCamVideo(model=ref, target_layer_names=['Mixed_5c'], class_dict=None, use_cuda=False,
                      input_spatial_size=(224, 224), normalizePerFrame=True, archType="I3D")
    cam_ref, out_ref = quiet(gc, x2[:1], 3)
    cam_or, out_or, low_or = gradcam_oracle.gradcam_i3d(sds, x2[:1], 3, (224, 224), True)
    close(out_or, out_ref.detach(), 1e-5, "Grad-CAM I3D: output")
    close(cam_or, cam_ref, 1e-4, "Grad-CAM I3D: cam [16,224,224] (class 3; may hold NaN slices)")
    cam_ref2, out_ref2 = quiet(gc, x2[1:2], None)
    cam_or2, _, low_or2 = gradcam_oracle.gradcam_i3d(sds, x2[1:2], None, (224, 224), True)
    close(cam_or2, cam_ref2, 1e-4, "Grad-CAM I3D: cam [16,224,224] (argmax class)")
    np.savez_compressed(os.path.join(GOLD, "gradcam_i3d.npz"), cam_lowres=low_or, cam_sample=cam_ref[::8, ::16, ::16],
                        output=out_ref.detach().numpy(), cam_lowres_argmax=low_or2,
                        cam_sample_argmax=cam_ref2[::8, ::16, ::16], output_argmax=out_ref2.detach().numpy())

    # ---------------------------------------------------------------- 5. I3D KTH geometry
    torch.manual_seed(0)
    refk = quiet(I3D_doubled_kth.Model, 6, last_stride=1, stride_mod_layers="", softMax=1, finalTimeLength=4,
                 dropout_keep_prob=0.5).eval()
    sdk = {k: v.detach().clone() for k, v in refk.state_dict().items()}
    xk = synthetic.clips(1, t=32, h=120, w=160)
    with torch.no_grad():
        pk = refk(xk)
    close(i3d_oracle.forward(sdk, xk, avg_pool=(4, 4, 5)).detach(), pk, 1e-5, "I3D KTH forward probs")
    gck = GradCamVideo(model=refk, target_layer_names=['Mixed_5c'], class_dict=None, use_cuda=False,
                       input_spatial_size=(160, 120), normalizePerFrame=True, archType="I3D")
    camk_ref, _ = quiet(gck, xk, None)
    camk_or, _, lowk = gradcam_oracle.gradcam_i3d(sdk, xk, None, (160, 120), True, avg_pool=(4, 4, 5))
    close(camk_or, camk_ref, 1e-4, "Grad-CAM I3D-KTH: cam [32,120,160]")
    np.savez_compressed(os.path.join(GOLD, "i3d_kth.npz"), probs=pk.numpy(), cam_lowres=lowk,
                        cam_sample=camk_ref[::8, ::12, ::16])

    # ---------------------------------------------------------------- 6. ConvLSTM (hid 4 shipped config and hid 32)
    for hid in (4, 32):
        torch.manual_seed(0)
        refc = quiet(CLSTM_4.Model, num_classes=6, nb_lstm_units=hid, channels=3, conv_kernel_size=(5, 5),
                     lstm_layers=2, step=32, conv_stride=2, image_size=(160, 120),
                     effective_step=[7, 15, 23, 31], batch_normalization=True, dropout=0.5).eval()
        # non-trivial BN statistics so the shared BatchNorm2d is exercised
        with torch.no_grad():
            refc.clstm.bn.running_mean.uniform_(-0.05, 0.05)
            refc.clstm.bn.running_var.uniform_(0.5, 1.5)
            refc.clstm.bn.weight.uniform_(0.5, 1.5)
            refc.clstm.bn.bias.uniform_(-0.1, 0.1)
        sdc = {k: v.detach().clone() for k, v in refc.state_dict().items()}
        xc = synthetic.clips(1, t=32, h=120, w=160) / 255.0
        mk = torch.rand(32, generator=torch.Generator().manual_seed(3)).requires_grad_()
        out = refc(ref_mask.perturb_sequence(xc, mk, 'reverse'))
        (gk,) = torch.autograd.grad(out[0, 2], mk)
        mo = mk.detach().clone().requires_grad_()
        out_o = clstm_oracle.forward(sdc, mask_oracle.perturb_sequence(xc, mo, 'reverse'), 2, hid)
        (go,) = torch.autograd.grad(out_o[0, 2], mo)
        close(out_o.detach(), out.detach(), 1e-5, "ConvLSTM hid %d forward logits" % hid)
        close(go, gk, 1e-4, "ConvLSTM hid %d d logit/d mask (reverse)" % hid)
        # Grad-CAM through the reference's own GradCamVideo with the child names it expects (SURVEY bug 8)
        class Shim(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.firstCLSTMLayer, self.endFC, self.sm = m.clstm, m.endFC, m.sm
                self.UseEntireSeq = m.use_entire_seq
                for a in ("nb_lstm_units", "im_size", "conv_stride", "pool_kernel_size", "lstm_layers", "effective_step"):
                    setattr(self, a, getattr(m, a))
        gcc = GradCamVideo(model=Shim(refc), target_layer_names=['firstCLSTMLayer'], class_dict=None, use_cuda=False,
                           input_spatial_size=(160, 120), normalizePerFrame=True, archType="CLSTM")
        camc_ref, outc_ref = quiet(gcc, xc, 2)
        camc_or, outc_or, lowc = gradcam_oracle.gradcam_clstm(sdc, xc, 2, (160, 120), True, num_layers=2, hidden=hid)
        close(outc_or, outc_ref.detach(), 1e-5, "ConvLSTM hid %d Grad-CAM: softmax output" % hid)
        close(camc_or, camc_ref, 2e-3, "ConvLSTM hid %d Grad-CAM: cam [32,120,160]" % hid)
        np.savez_compressed(os.path.join(GOLD, "clstm_hid%d.npz" % hid), logits=out.detach().numpy(),
                            dmask=gk.numpy(), mask=mk.detach().numpy(), cam_lowres=lowc,
                            cam_sample=camc_ref[::8, ::12, ::16], cam_output=outc_ref.detach().numpy())
    print("oracle pinned against the reference; golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
