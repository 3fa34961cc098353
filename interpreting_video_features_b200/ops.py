"""Thin tensor-level wrappers over the C ABI (one function per entry point of include/ivf.h).

Activations are channels-last views described by `Act`: a torch buffer plus the per-pixel channel
stride (`ld`) and the channel offset/extent of the slice — the Inception concat of
pt/models/I3D_doubled.py:146 is never materialised, each branch writes its slice.
torch is used for allocation and streams only; every computation below is a libivf kernel.
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import (EP_ACCUM, EP_AFFINE, EP_MASK, EP_OUT_F32, EP_RELU, IVF_BF16, IVF_F32, POOL_NONNEG,
                   PFMT_NCDHW_F32, PFMT_NDHWC_F32, PFMT_S2D2_BF16, PFMT_S2D_BF16, PFMT_TBHWC_F32, ConvDesc,
                   PoolDesc, check, ptr)


def fill_zero(t):
    """Zero a contiguous device tensor with a libivf launch on the current stream of its device."""
    assert t.is_contiguous() and (t.numel() * t.element_size()) % 4 == 0, "fill_zero: 4-byte granularity"
    if t.numel():
        check(_lib.load().ivf_fill_u32(_lib.handle(t.device), ptr(t), t.numel() * t.element_size(), 0,
                                       _lib.stream_ptr(t.device)), "ivf_fill_u32")
    return t


def zeros(shape, dtype, device):
    """torch.zeros without a torch kernel: allocation by the caching allocator, fill by libivf."""
    t = torch.empty(shape, dtype=dtype, device=device)
    if (t.numel() * t.element_size()) % 4:
        return t.zero_()  # odd byte counts (tiny uint8 buffers): not worth a kernel variant
    return fill_zero(t)


def u8_to_f32(src, dst):
    """dst (fp32, device) = float(src) for a uint8 device tensor of the same element count."""
    assert src.dtype == torch.uint8 and dst.dtype == torch.float32 and src.numel() == dst.numel()
    assert src.is_contiguous() and dst.is_contiguous()
    check(_lib.load().ivf_u8_to_f32(_lib.handle(dst.device), ptr(src), ptr(dst), src.numel(),
                                    _lib.stream_ptr(dst.device)), "ivf_u8_to_f32")
    return dst


def one_hot(targets, out):
    """out[n, ncls] (fp32) = one_hot(targets[n]); targets int32 on the device."""
    n, ncls = out.shape
    assert targets.dtype == torch.int32 and targets.numel() == n
    check(_lib.load().ivf_one_hot(_lib.handle(out.device), ptr(targets), n, ncls, ptr(out),
                                  _lib.stream_ptr(out.device)), "ivf_one_hot")
    return out


def argmax_rows(x, out):
    """out[n] (int32) = first arg-maximum of each row of x[n, ncls] (fp32), on the device."""
    n, ncls = x.shape
    check(_lib.load().ivf_argmax_rows(_lib.handle(x.device), ptr(x), n, ncls, ptr(out), _lib.stream_ptr(x.device)),
          "ivf_argmax_rows")
    return out


def as_int32_targets(targets, device):
    """Class indices as an int32 device tensor (converted on the host when they arrive from the host)."""
    t = torch.as_tensor(targets)
    if not t.is_cuda:
        return t.to(torch.int32).to(device, non_blocking=True)
    return t.to(device=device, dtype=torch.int32)


class Act:
    """A channel slice [coff, coff+c) of a channels-last buffer of logical shape (n,d,h,w,ld)."""

    __slots__ = ("buf", "n", "d", "h", "w", "ld", "coff", "c")

    def __init__(self, buf, n, d, h, w, ld, coff=0, c=None):
        self.buf, self.n, self.d, self.h, self.w, self.ld, self.coff = buf, n, d, h, w, ld, coff
        self.c = ld - coff if c is None else c

    @staticmethod
    def empty(n, d, h, w, c, dtype, device, zero=False):
        buf = zeros((n, d, h, w, c), dtype, device) if zero else torch.empty((n, d, h, w, c), dtype=dtype, device=device)
        return Act(buf, n, d, h, w, c, 0, c)

    def slice(self, coff, c):
        assert coff + c <= self.c
        return Act(self.buf, self.n, self.d, self.h, self.w, self.ld, self.coff + coff, c)

    def like(self, dtype=None, zero=False):
        return Act.empty(self.n, self.d, self.h, self.w, self.c, dtype or self.buf.dtype,
                         self.buf.device, zero)

    @property
    def pixels(self):
        return self.n * self.d * self.h * self.w

    def tensor(self):
        """(n,d,h,w,c) torch view of the slice."""
        return self.buf.view(self.n, self.d, self.h, self.w, self.ld)[..., self.coff:self.coff + self.c]

    def ncdhw(self):
        """fp32 (n,c,d,h,w) copy — test/interop helper, not on the hot path."""
        return self.tensor().permute(0, 4, 1, 2, 3).float().contiguous()


def same_pad(size, k, s):
    """TF 'same' padding of pt/models/I3D_doubled.py:77-101: (front, back, out)."""
    total = max(k - s, 0) if size % s == 0 else max(k - (size % s), 0)
    front = total // 2
    return front, total - front, int(math.ceil(size / s))


def conv_desc(x, out, kernel, stride, pad_front, flags=0, scale=None, acc_in=None, mask=None, transposed=0,
              cin=None, cout=None, plan=None):
    """The ivf_conv_desc of a launch (also used without launching: tile-plan queries of the tuner).
    plan = (kwm, mt, acc, ncta, ntiles) requests a slab-kernel tile plan, None/0 = cost model."""
    d = ConvDesc()
    d.n, d.id, d.ih, d.iw = x.n, x.d, x.h, x.w
    d.od, d.oh, d.ow = out.d, out.h, out.w
    d.cin = x.c if cin is None else cin
    d.cout = out.c if cout is None else cout
    d.kd, d.kh, d.kw = kernel
    d.sd, d.sh, d.sw = stride
    d.pd, d.ph, d.pw = pad_front
    d.transposed = transposed
    d.in_ld, d.in_coff = x.ld, x.coff
    d.out_ld, d.out_coff = out.ld, out.coff
    if mask is not None:
        d.mask_ld, d.mask_coff = mask.ld, mask.coff
        flags |= EP_MASK
    if acc_in is not None:
        flags |= EP_ACCUM
    if scale is not None:
        flags |= EP_AFFINE
    d.dtype = _lib.dtype_code(x.buf)
    if out.buf.dtype == torch.float32 and x.buf.dtype == torch.bfloat16:
        flags |= EP_OUT_F32
    d.flags = flags
    if plan is not None:  # (kwm, mt, acc, ncta, ntiles[, ds]); five-entry requests leave depth stacking to the library's default
        d.plan_kwm, d.plan_mt, d.plan_acc, d.plan_ncta, d.plan_ntiles = plan[:5]
        d.plan_ds = plan[5] if len(plan) > 5 else 0
    return d


def conv3d(x, w, out, kernel, stride, pad_front, flags=0, scale=None, shift=None, acc_in=None,
           mask=None, mask_scale=None, transposed=0, cin=None, cout=None, plan=None):
    """out = epilogue(conv(x, w)); see ivf_conv3d. x/out/mask are Act; w is the packed weight."""
    d = conv_desc(x, out, kernel, stride, pad_front, flags, scale, acc_in, mask, transposed, cin, cout, plan)
    check(_lib.load().ivf_conv3d(_lib.handle(x.buf.device), C.byref(d), ptr(x.buf), ptr(w), ptr(scale),
                                 ptr(shift), ptr(acc_in.buf if isinstance(acc_in, Act) else acc_in),
                                 ptr(mask.buf if mask is not None else None), ptr(mask_scale),
                                 ptr(out.buf), _lib.stream_ptr(x.buf.device)), "ivf_conv3d")
    return out


def conv3d_pair(a, b):
    """Two independent convolutions as one launch where the library can group them (ivf_conv3d_pair: the two 3x3x3
    branches of an Inception module), else one after the other.  a, b: dicts of conv3d's arguments."""
    def parts(k):
        d = conv_desc(k["x"], k["out"], k["kernel"], k["stride"], k["pad_front"], k.get("flags", 0), k.get("scale"),
                      k.get("acc_in"), k.get("mask"), k.get("transposed", 0), None, None, k.get("plan"))
        acc, mask = k.get("acc_in"), k.get("mask")
        return d, [ptr(k["x"].buf), ptr(k["w"]), ptr(k.get("scale")), ptr(k.get("shift")),
                   ptr(acc.buf if isinstance(acc, Act) else acc), ptr(mask.buf if mask is not None else None),
                   ptr(k.get("mask_scale")), ptr(k["out"].buf)]
    da, pa = parts(a)
    db, pb = parts(b)
    dev = a["x"].buf.device
    check(_lib.load().ivf_conv3d_pair(_lib.handle(dev), C.byref(da), *pa, C.byref(db), *pb, _lib.stream_ptr(dev)),
          "ivf_conv3d_pair")


def conv1x1_split(x, w, out, x2=None, out2=None, flags=0, scale=None, shift=None, acc_in=None, mask=None,
                  mask_scale=None):
    """1x1x1 bf16 convolution with two sources (channels of x then of x2) and/or two destinations (produced
    channels fill out, then out2); see ivf_conv3d_split.  Totals are taken from the Acts."""
    cin = x.c + (x2.c if x2 is not None else 0)
    cout = out.c + (out2.c if out2 is not None else 0)
    d = conv_desc(x, out, (1, 1, 1), (1, 1, 1), (0, 0, 0), flags, scale, acc_in, mask, 0, cin, cout)
    sp = _lib.ConvSplit()
    if out2 is not None:
        sp.split_cout, sp.out2_ld, sp.out2_coff = out.c, out2.ld, out2.coff
    if x2 is not None:
        sp.split_cin, sp.in2_ld, sp.in2_coff = x.c, x2.ld, x2.coff
    check(_lib.load().ivf_conv3d_split(_lib.handle(x.buf.device), C.byref(d), C.byref(sp), ptr(x.buf),
                                       ptr(x2.buf if x2 is not None else None), ptr(w), ptr(scale), ptr(shift),
                                       ptr(acc_in.buf if isinstance(acc_in, Act) else acc_in),
                                       ptr(mask.buf if mask is not None else None), ptr(mask_scale), ptr(out.buf),
                                       ptr(out2.buf if out2 is not None else None), _lib.stream_ptr(x.buf.device)),
          "ivf_conv3d_split")
    return out


def _pool_desc(x, out, kernel, stride, pad_front, mask=None, flags=0):
    d = PoolDesc()
    d.n, d.id, d.ih, d.iw, d.c = x.n, x.d, x.h, x.w, x.c
    d.od, d.oh, d.ow = out.d, out.h, out.w
    d.kd, d.kh, d.kw = kernel
    d.sd, d.sh, d.sw = stride
    d.pd, d.ph, d.pw = pad_front
    d.in_ld, d.in_coff = x.ld, x.coff
    d.out_ld, d.out_coff = out.ld, out.coff
    if mask is not None:
        d.mask_ld, d.mask_coff = mask.ld, mask.coff
    d.flags = flags
    return d


def maxpool3d_fwd(x, out, argmax, kernel, stride, pad_front, relu_bits=None, nonneg=False):
    """relu_bits: optional uint8 [x.pixels, c/8] that receives one bit per input element (element > 0) for the
    backward pass (ivf_maxpool3d_fwd_bits).  nonneg: the caller vouches that x holds no negative value (a ReLU
    output): IVF_POOL_NONNEG, the packed bf16 kernels then order bit patterns directly."""
    d = _pool_desc(x, out, kernel, stride, pad_front, flags=POOL_NONNEG if nonneg else 0)
    d.dtype = _lib.dtype_code(x.buf)
    check(_lib.load().ivf_maxpool3d_fwd_bits(_lib.handle(x.buf.device), C.byref(d), ptr(x.buf), ptr(out.buf),
                                             ptr(argmax), ptr(relu_bits), _lib.stream_ptr(x.buf.device)), "ivf_maxpool3d_fwd")
    return out


def maxpool3d_bwd(dy, argmax, dx, kernel, stride, pad_front, acc_in=None, mask=None, mask_scale=None, relu_bits=None):
    """dx (Act shaped like the pool input) from dy (Act shaped like the pool output)."""
    flags = 0
    if acc_in is not None:
        flags |= EP_ACCUM
    if mask is not None:
        flags |= EP_MASK
    if dx.buf.dtype == torch.float32 and dy.buf.dtype == torch.bfloat16:
        flags |= EP_OUT_F32
    d = _pool_desc(dx, dy, kernel, stride, pad_front, mask, flags)
    d.dtype = _lib.dtype_code(dy.buf)
    check(_lib.load().ivf_maxpool3d_bwd_bits(_lib.handle(dy.buf.device), C.byref(d), ptr(dy.buf), ptr(argmax),
                                             ptr(acc_in.buf if isinstance(acc_in, Act) else acc_in),
                                             ptr(mask.buf if mask is not None else None), ptr(relu_bits),
                                             ptr(mask_scale), ptr(dx.buf), _lib.stream_ptr(dy.buf.device)), "ivf_maxpool3d_bwd")
    return dx


def head_workspace(n, c, ncls, device):
    """The caller-owned partial-logit buffer of ivf_i3d_head_fwd: one per engine (engines on parallel streams
    must not share it)."""
    nbytes = int(_lib.load().ivf_i3d_head_workspace_bytes(n, c, ncls))
    return torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=device)


def head_fwd(feat, w, b, softmax, out, logits=None, workspace=None):
    """feat: Act (whole map is pooled); w fp32 [ncls, c]; out fp32 [n, ncls]."""
    assert feat.coff == 0
    p = feat.d * feat.h * feat.w
    if workspace is None:  # one-off calls (tests, the per-op surface); engines pass their own buffer
        workspace = head_workspace(feat.n, feat.c, w.shape[0], feat.buf.device)
    check(_lib.load().ivf_i3d_head_fwd(_lib.handle(feat.buf.device), _lib.dtype_code(feat.buf),
                                       ptr(feat.buf), feat.n, p, feat.c, feat.ld, ptr(w), ptr(b),
                                       w.shape[0], int(bool(softmax)), ptr(logits), ptr(out), ptr(workspace),
                                       workspace.numel() * 4, _lib.stream_ptr(feat.buf.device)),
          "ivf_i3d_head_fwd")
    return out


def head_bwd(dfeat, w, softmax, out, dout, mask=None, mask_scale=None):
    """dfeat: Act to fill (dtype of the activations, or fp32)."""
    assert dfeat.coff == 0
    p = dfeat.d * dfeat.h * dfeat.w
    flags = 0
    dtype = _lib.dtype_code(mask.buf) if mask is not None else _lib.dtype_code(dfeat.buf)
    if mask is not None:
        flags |= EP_MASK
    if dfeat.buf.dtype == torch.float32 and dtype == IVF_BF16:
        flags |= EP_OUT_F32
    check(_lib.load().ivf_i3d_head_bwd(_lib.handle(dfeat.buf.device), dtype, dfeat.n, p, dfeat.c, dfeat.ld,
                                       ptr(w), w.shape[0], int(bool(softmax)), ptr(out), ptr(dout), flags,
                                       ptr(mask.buf if mask is not None else None),
                                       mask.ld if mask is not None else 0,
                                       mask.coff if mask is not None else 0, ptr(mask_scale),
                                       ptr(dfeat.buf), _lib.stream_ptr(dfeat.buf.device)), "ivf_i3d_head_bwd")
    return dfeat


_MODES = {"freeze": 0, "reverse": 1}


def perturb_fwd(x, mask, mode, out_fmt, out):
    """x fp32 [b,c,t,h,w]; mask fp32 [t] (shared) or [b,t]; out preallocated per out_fmt."""
    b, c, t, hh, ww = x.shape
    bstride = 0 if mask.dim() == 1 else t
    check(_lib.load().ivf_perturb_fwd(_lib.handle(x.device), _MODES[mode], ptr(x), ptr(mask), bstride, b, c,
                                      t, hh, ww, out_fmt, ptr(out), _lib.stream_ptr(x.device)), "ivf_perturb_fwd")
    return out


def perturb_bwd(x, mask, mode, out_fmt, gout, dmask):
    """dmask fp32 [b,t] (one row per clip, also for a shared mask)."""
    b, c, t, hh, ww = x.shape
    bstride = 0 if mask.dim() == 1 else t
    check(_lib.load().ivf_perturb_bwd(_lib.handle(x.device), _MODES[mode], ptr(x), ptr(mask), bstride, b, c,
                                      t, hh, ww, out_fmt, _lib.dtype_code(gout), ptr(gout), ptr(dmask),
                                      _lib.stream_ptr(x.device)), "ivf_perturb_bwd")
    return dmask


def mask_loss_adam(m, exp_avg, exp_avg_sq, dclass, step, lam1, lam2, lr=0.2, beta1=0.9, beta2=0.999,
                   eps=1e-8, losses=None, sig_out=None, step_dev=None):
    nclip, t = m.shape
    check(_lib.load().ivf_mask_loss_adam(_lib.handle(m.device), ptr(m), ptr(exp_avg), ptr(exp_avg_sq),
                                         ptr(dclass), nclip, t, int(step), ptr(step_dev), lam1, lam2, lr,
                                         beta1, beta2, eps, ptr(losses), ptr(sig_out), _lib.stream_ptr(m.device)),
          "ivf_mask_loss_adam")


def sigmoid(m, out):
    check(_lib.load().ivf_sigmoid(_lib.handle(m.device), ptr(m), ptr(out), m.numel(), _lib.stream_ptr(m.device)),
          "ivf_sigmoid")
    return out


def select_scores(probs, targets, out):
    """out[n] = probs[n, targets[n]] (targets int32 on the device)."""
    n, ncls = probs.shape
    check(_lib.load().ivf_select_scores(_lib.handle(probs.device), ptr(probs), ptr(targets), n, ncls, ptr(out),
                                        _lib.stream_ptr(probs.device)), "ivf_select_scores")
    return out


def init_mask_select(scores, t, threshold, raw, chosen=None):
    """scores fp32 [1 + t/2, n] (rows: original, fully frozen, centred windows) -> raw masks [n, t] on the device."""
    rows, n = scores.shape
    assert rows == 1 + max(t // 2, 1) and tuple(raw.shape) == (n, t)
    check(_lib.load().ivf_init_mask_select(_lib.handle(scores.device), ptr(scores), n, t, float(threshold), ptr(raw),
                                           ptr(chosen), _lib.stream_ptr(scores.device)), "ivf_init_mask_select")
    return raw


def tv_norm(mask, p, q, val, dmask=None):
    check(_lib.load().ivf_tv_norm(_lib.handle(mask.device), ptr(mask), mask.numel(), float(p), float(q),
                                  ptr(val), ptr(dmask), _lib.stream_ptr(mask.device)), "ivf_tv_norm")
    return val


def gradcam(act, grad, step, hout, wout, per_frame, cam, cam_lowres=None):
    """act/grad: Act of the target layer (same geometry); cam fp32 [n, tp*step, hout, wout] (None: only the
    low-resolution map cam_lowres [n, tp, hp, wp] is written)."""
    assert act.coff == 0 and grad.coff == 0 and act.ld == grad.ld
    check(_lib.load().ivf_gradcam(_lib.handle(act.buf.device), _lib.dtype_code(act.buf),
                                  _lib.dtype_code(grad.buf), ptr(act.buf), ptr(grad.buf), act.n, act.d,
                                  act.h, act.w, act.c, act.ld, step, hout, wout, int(bool(per_frame)),
                                  ptr(cam), ptr(cam_lowres), _lib.stream_ptr(act.buf.device)), "ivf_gradcam")
    return cam


def viz_triptych(clip, cam, pert, mask, out, draw_dots=True):
    """clip [3,T,H,W] fp32/uint8, cam fp32 [T,H,W], pert fp32 [3,T,H,W], mask fp32 [T] -> out uint8 [T,H,3W,3] (BGR)."""
    c, t, hh, ww = clip.shape
    assert c == 3 and tuple(cam.shape) == (t, hh, ww) and tuple(pert.shape) == (3, t, hh, ww)
    assert tuple(out.shape) == (t, hh, 3 * ww, 3) and out.dtype == torch.uint8
    code = _lib.IVF_U8 if clip.dtype == torch.uint8 else IVF_F32
    check(_lib.load().ivf_viz_triptych(_lib.handle(out.device), code, ptr(clip), ptr(cam), ptr(pert), ptr(mask), t, hh,
                                       ww, int(bool(draw_dots)), ptr(out), _lib.stream_ptr(out.device)),
          "ivf_viz_triptych")
    return out


def clstm_gates_fwd(pre, c_prev, c_next, h_next, gate_act, unit_major=False):
    m, four_hid = pre.shape
    check(_lib.load().ivf_clstm_gates_fwd(_lib.handle(pre.device), _lib.dtype_code(h_next), ptr(pre),
                                          ptr(c_prev), m, four_hid // 4, ptr(c_next), ptr(h_next),
                                          ptr(gate_act), int(bool(unit_major)), _lib.stream_ptr(pre.device)),
          "ivf_clstm_gates_fwd")


def clstm_gates_bwd(gate_act, c_prev, c_next, dh, dc_io, dgates, unit_major=False):
    m, four_hid = gate_act.shape
    check(_lib.load().ivf_clstm_gates_bwd(_lib.handle(dh.device), _lib.dtype_code(dgates), ptr(gate_act),
                                          ptr(c_prev), ptr(c_next), ptr(dh), ptr(dc_io), m, four_hid // 4,
                                          ptr(dgates), int(bool(unit_major)), _lib.stream_ptr(dh.device)),
          "ivf_clstm_gates_bwd")


def conv_lstm_step(h_prev, w, pre_x, c_prev, c_next, h_next, gate_act, kernel, pad_front, plan=None):
    """The recurrent ConvLSTM step as one kernel (ivf_conv3d_lstm): h_prev Act [b,1,h,w,hid] bf16, w the packed
    UNIT-MAJOR h-convolution weights, pre_x Act [b,1,h,w,4hid] fp32 (x-convolution + bias of this step); writes
    c_next (fp32 [m,hid]), h_next (bf16 Act) and gate_act (fp32 [m,4hid])."""
    d = conv_desc(h_prev, pre_x, kernel, (1, 1, 1), pad_front, 0, None, None, None, 0, plan=plan)
    d.flags = 0
    check(_lib.load().ivf_conv3d_lstm(_lib.handle(h_prev.buf.device), C.byref(d), ptr(h_prev.buf), ptr(w),
                                      ptr(pre_x.buf), ptr(c_prev), ptr(c_next), ptr(h_next.buf), ptr(gate_act),
                                      _lib.stream_ptr(h_prev.buf.device)), "ivf_conv3d_lstm")


def bn_pool2d_fwd(x, scale, shift, y, argmax, s2d=False):
    n, hh, ww, c = x.shape
    check(_lib.load().ivf_bn_pool2d_fwd(_lib.handle(x.device), _lib.dtype_code(x), ptr(x), n, hh, ww, c,
                                        ptr(scale), ptr(shift), ptr(y), ptr(argmax), int(s2d),
                                        _lib.stream_ptr(x.device)), "ivf_bn_pool2d_fwd")


def bn_pool2d_bwd(dy, argmax, scale, dx, acc_in=None, s2d=False):
    n, hh, ww, c = dx.shape
    check(_lib.load().ivf_bn_pool2d_bwd(_lib.handle(dy.device), _lib.dtype_code(dy), ptr(dy), ptr(argmax), n,
                                        hh, ww, c, ptr(scale), ptr(acc_in), ptr(dx), int(s2d),
                                        _lib.stream_ptr(dy.device)), "ivf_bn_pool2d_bwd")


def probe_im2col(x, kernel, stride, pad_front, out_dhw, m0, tap, c0):
    """Bring-up probe: returns the [128, kchunk] bf16 tile the conv kernel's TMA would stage."""
    d = ConvDesc()
    d.n, d.id, d.ih, d.iw = x.n, x.d, x.h, x.w
    d.od, d.oh, d.ow = out_dhw
    d.cin, d.cout = x.c, 16
    d.kd, d.kh, d.kw = kernel
    d.sd, d.sh, d.sw = stride
    d.pd, d.ph, d.pw = pad_front
    d.in_ld, d.in_coff = x.ld, x.coff
    d.out_ld, d.out_coff = 16, 0
    d.dtype = IVF_BF16
    kch = _lib.load().ivf_conv_bf16_kchunk(x.c)
    tile = torch.empty((128, kch), dtype=torch.bfloat16, device=x.buf.device)
    check(_lib.load().ivf_probe_im2col(_lib.handle(x.buf.device), C.byref(d), ptr(x.buf), m0, tap, c0,
                                       ptr(tile), _lib.stream_ptr(x.buf.device)), "ivf_probe_im2col")
    return tile


# ---------------------------------------------------------------------------------------- training step (train.cu)
def bn_train_fwd(z, gamma, beta, eps, momentum, running_mean, running_var, save_mean, save_rstd, ws, y, relu=True):
    """y = relu(BatchNorm(z)) with BATCH statistics over all pixels of the Act z (ivf_bn_train_fwd); the running
    statistics (fp32 [c], may be None) are updated in place; ws: float64 [2c] scratch."""
    assert z.pixels == y.pixels and z.c == y.c and ws.dtype == torch.float64 and ws.numel() >= 2 * z.c
    check(_lib.load().ivf_bn_train_fwd(_lib.handle(z.buf.device), _lib.dtype_code(z.buf), ptr(z.buf), z.ld, z.coff,
                                       z.pixels, z.c, ptr(gamma), ptr(beta), eps, momentum, ptr(running_mean),
                                       ptr(running_var), ptr(save_mean), ptr(save_rstd), ptr(ws), ptr(y.buf), y.ld,
                                       y.coff, int(bool(relu)), _lib.stream_ptr(z.buf.device)), "ivf_bn_train_fwd")
    return y


def bn_train_bwd(dy, y, z, gamma, save_mean, save_rstd, ws, dz, dgamma, dbeta):
    """dz, dgamma, dbeta of relu(BatchNorm_train(z)) given dy (Act) and the forward's y (Act, None: no ReLU)."""
    check(_lib.load().ivf_bn_train_bwd(_lib.handle(z.buf.device), _lib.dtype_code(z.buf), _lib.dtype_code(dy.buf),
                                       ptr(dy.buf), dy.ld, dy.coff,
                                       ptr(y.buf if y is not None else None), y.ld if y is not None else 0,
                                       y.coff if y is not None else 0, ptr(z.buf), z.ld, z.coff, z.pixels, z.c,
                                       ptr(gamma), ptr(save_mean), ptr(save_rstd), ptr(ws), ptr(dz.buf), dz.ld, dz.coff,
                                       ptr(dgamma), ptr(dbeta), _lib.stream_ptr(z.buf.device)), "ivf_bn_train_bwd")
    return dz


def conv3d_wgrad(x, dz, dw, kernel, stride, pad_front):
    """dw (fp32, the nn.Conv3d parameter's [cout, cin, kd, kh, kw] layout) = weight gradient of the convolution that
    maps the Act x to an Act shaped like dz (ivf_conv3d_wgrad)."""
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == dz.c * x.c * kernel[0] * kernel[1] * kernel[2]
    d = conv_desc(x, dz, kernel, stride, pad_front)
    d.flags = 0
    d.dtype = _lib.dtype_code(dz.buf)
    check(_lib.load().ivf_conv3d_wgrad(_lib.handle(x.buf.device), C.byref(d), _lib.dtype_code(x.buf), ptr(x.buf),
                                       ptr(dz.buf), ptr(dw),
                                       _lib.stream_ptr(x.buf.device)), "ivf_conv3d_wgrad")
    return dw


def conv3d_wgrad_s2d(x_s2d, dz, dw, kernel_eff, pad_front_eff):
    """Weight gradient of a stride-2 layer from the 2x2x2 space-to-depth record of its input (Act of 8 * ci channels,
    bf16): dw in the original [cout, ci, kd, kh, kw] layout (ivf_conv3d_wgrad_s2d)."""
    co, ci, kd, kh, kw = dw.shape
    assert dw.dtype == torch.float32 and dw.is_contiguous() and x_s2d.c == 8 * ci and dz.c == co
    d = conv_desc(x_s2d, dz, kernel_eff, (1, 1, 1), pad_front_eff)
    d.flags = 0
    d.dtype = _lib.dtype_code(dz.buf)
    check(_lib.load().ivf_conv3d_wgrad_s2d(_lib.handle(dz.buf.device), C.byref(d), ptr(x_s2d.buf), ptr(dz.buf), ptr(dw),
                                           ci, kd, kh, kw, _lib.stream_ptr(dz.buf.device)), "ivf_conv3d_wgrad_s2d")
    return dw


def head_train_fwd(feat, drop, w, bias, target, pooled, logits, dlogits, loss):
    """Training-mode classifier head + cross-entropy (ivf_head_train_fwd); feat: Act whose whole map is pooled."""
    p = feat.d * feat.h * feat.w
    check(_lib.load().ivf_head_train_fwd(_lib.handle(feat.buf.device), _lib.dtype_code(feat.buf), ptr(feat.buf), feat.ld,
                                         feat.coff, feat.n, p, feat.c, ptr(drop), ptr(w), ptr(bias), ptr(target),
                                         w.shape[0], ptr(pooled), ptr(logits), ptr(dlogits), ptr(loss),
                                         _lib.stream_ptr(feat.buf.device)), "ivf_head_train_fwd")
    return loss


def head_train_bwd(dlogits, pooled, drop, w, dw, db, dfeat):
    p = dfeat.d * dfeat.h * dfeat.w
    check(_lib.load().ivf_head_train_bwd(_lib.handle(dfeat.buf.device), _lib.dtype_code(dfeat.buf), ptr(dlogits),
                                         ptr(pooled), ptr(drop), ptr(w), dfeat.n, p, dfeat.c, w.shape[0], ptr(dw),
                                         ptr(db), ptr(dfeat.buf), dfeat.ld, dfeat.coff,
                                         _lib.stream_ptr(dfeat.buf.device)), "ivf_head_train_bwd")
    return dfeat


def optim_step(kind, p, g, s1, s2, lr, beta1, beta2, eps, weight_decay, step):
    """In-place torch.optim.SGD (kind 'sgd': beta1 = momentum) / Adam ('adam') update of the fp32 tensor p."""
    assert p.dtype == torch.float32 and g.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
    check(_lib.load().ivf_optim_step(_lib.handle(p.device), {"sgd": 0, "adam": 1}[kind], ptr(p), ptr(g), ptr(s1), ptr(s2),
                                     p.numel(), lr, beta1, beta2, eps, weight_decay, step, _lib.stream_ptr(p.device)),
          "ivf_optim_step")
    return p


def dropout_mask(out, p, seed):
    """out (fp32, device) = the scaled mask of nn.Dropout(p): 0 with probability p, else 1/(1-p)."""
    check(_lib.load().ivf_dropout_mask(_lib.handle(out.device), ptr(out), out.numel(), p, seed,
                                       _lib.stream_ptr(out.device)), "ivf_dropout_mask")
    return out


OPTIM_CHUNK = 4096


def optim_table(params, grads, state1, state2, device):
    """Device table for optim_step_multi: one row {p, g, s1, s2, n} per chunk of at most OPTIM_CHUNK elements of
    every tensor (lists of equally shaped fp32 tensors; state lists may hold None)."""
    rows = []
    for p, g, a, b in zip(params, grads, state1, state2):
        assert p.dtype == torch.float32 and g.dtype == torch.float32 and p.is_contiguous() and g.is_contiguous()
        for off in range(0, p.numel(), OPTIM_CHUNK):
            n = min(OPTIM_CHUNK, p.numel() - off)
            rows.append([p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, (a.data_ptr() + 4 * off) if a is not None else 0,
                         (b.data_ptr() + 4 * off) if b is not None else 0, n])
    return torch.tensor(rows, dtype=torch.int64).to(device)


def optim_step_multi(kind, table, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(_lib.load().ivf_optim_step_multi(_lib.handle(table.device), {"sgd": 0, "adam": 1}[kind], ptr(table),
                                           table.shape[0], lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                                           _lib.stream_ptr(table.device)), "ivf_optim_step_multi")
