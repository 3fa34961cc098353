"""Test infrastructure: the weight packings of the convolution kernels written with plain torch ops on any
device (the round-1 host packers).  The product packs with the ivf_pack_weights kernel (engine.pack); the GPU
tests require the two to agree bit for bit, and the CPU tests prove these layouts equivalent to the reference
convolutions (tests/test_cpu_host_logic.py)."""
import numpy as np
import torch


def cin_pad(c):
    """ivf_conv_bf16_cin_pad (checked against the library in tests/test_cpu_abi.py)."""
    if c <= 16:
        return 16
    if c <= 32:
        return 32
    if c <= 64:
        return 64
    return (c + 15) // 16 * 16


def s2d_weight(w, win):
    """(co,ci,k,k,k) stride-2 kernel -> (co, 8*ci, win,win,win) stride-1 kernel over the
    space-to-depth input; channel = ((a*2+b)*2+c)*ci + ch, source tap = 2*delta + parity."""
    co, ci, kd, kh, kw = w.shape
    out = w.new_zeros(co, 8 * ci, win, win, win)
    for a in range(2):
        for b in range(2):
            for c in range(2):
                blk = ((a * 2 + b) * 2 + c) * ci
                for dt in range(win):
                    kt = 2 * dt + a
                    if kt >= kd:
                        continue
                    for dh in range(win):
                        kh_ = 2 * dh + b
                        if kh_ >= kh:
                            continue
                        for dw in range(win):
                            kw_ = 2 * dw + c
                            if kw_ >= kw:
                                continue
                            out[:, blk:blk + ci, dt, dh, dw] = w[:, :, kt, kh_, kw_]
    return out


def s2d_weight_2d(w, win):
    """(co,ci,k,k) stride-2 kernel -> (co,4ci,win,win) stride-1 kernel over the 2-D space-to-depth
    input; channel = (a*2+b)*ci + ch, source tap = 2*delta + parity."""
    co, ci, kh, kw = w.shape
    out = w.new_zeros(co, 4 * ci, win, win)
    for a in range(2):
        for b in range(2):
            blk = (a * 2 + b) * ci
            for dh in range(win):
                if 2 * dh + a >= kh:
                    continue
                for dw in range(win):
                    if 2 * dw + b >= kw:
                        continue
                    out[:, blk:blk + ci, dh, dw] = w[:, :, 2 * dh + a, 2 * dw + b]
    return out


def _pack_bf16(wk, n_pad, k_pad):
    n, taps, k = wk.shape
    out = torch.zeros((n_pad, taps, k_pad), dtype=torch.bfloat16, device=wk.device)
    out[:n, :, :k] = wk.to(torch.bfloat16)
    return out.contiguous()


def pack_fwd(w, mode, n_pad=None, k_pad=None):
    co, ci = w.shape[:2]
    if mode == "fp32":
        return w.permute(2, 3, 4, 1, 0).reshape(-1, co).contiguous().float()
    return _pack_bf16(w.permute(0, 2, 3, 4, 1).reshape(co, -1, ci), n_pad, k_pad)


def pack_dgrad(w, mode, n_pad=None, k_pad=None):
    co, ci = w.shape[:2]
    if mode == "fp32":
        return w.permute(2, 3, 4, 0, 1).reshape(-1, ci).contiguous().float()
    return _pack_bf16(w.flip(2, 3, 4).permute(1, 2, 3, 4, 0).reshape(ci, -1, co), n_pad, k_pad)


def pack_dgrad_two_sources(w_first, w_second, n_pad, k_pad):
    ci, c_first = w_first.shape[1], w_first.shape[0]
    k1 = (c_first + 63) // 64 * 64
    wk = w_first.new_zeros(ci, 1, k1 + w_second.shape[0])
    wk[:, 0, :c_first] = w_first.reshape(c_first, ci).t()
    wk[:, 0, k1:] = w_second.reshape(w_second.shape[0], ci).t()
    return _pack_bf16(wk, n_pad, k_pad)


def pad_gates(w4, he, cin_eff):
    """four per-gate [hid, cin, k, k] -> [4*he, cin_eff, 1, k, k], zero padded (ConvLSTM gate stacking)."""
    hid, cin, k, _ = w4[0].shape
    out = torch.zeros((4 * he, cin_eff, 1, k, k))
    for gi, w in enumerate(w4):
        out[gi * he:gi * he + hid, :cin, 0] = w
    return out


def emulate_pack_kernel(w, dgrad, layout, s2d, ci_stride, n_pad, k_pad, n_off, k_off, dst):
    """Line-by-line numpy restatement of pack_weights_kernel's index arithmetic (csrc/pack.cu) — lets the CPU
    suite check the kernel's mapping against the torch packers without a GPU."""
    co_n, ci, kd, kh, kw = w.shape
    fd, fh, fw = s2d
    wd, wh, ww = (kd + fd - 1) // fd, (kh + fh - 1) // fh, (kw + fw - 1) // fw
    cis = max(ci_stride, ci)
    ceff = fd * fh * fw * cis
    taps = wd * wh * ww
    swap, flip = dgrad != 0, dgrad == 1
    nsrc, ksrc = (ceff, co_n) if swap else (co_n, ceff)
    wn = w.numpy()
    for n in range(nsrc):
        for tap in range(taps):
            for k in range(ksrc):
                co, ce = (k, n) if swap else (n, k)
                dw_, dh_, dt_ = tap % ww, (tap // ww) % wh, tap // (ww * wh)
                if flip:
                    dw_, dh_, dt_ = ww - 1 - dw_, wh - 1 - dh_, wd - 1 - dt_
                ch, par = ce % cis, ce // cis
                c_, b_, a_ = par % fw, (par // fw) % fh, par // (fw * fh)
                kt, khh, kww = fd * dt_ + a_, fh * dh_ + b_, fw * dw_ + c_
                v = wn[co, ch, kt, khh, kww] if (ch < ci and kt < kd and khh < kh and kww < kw) else 0.0
                nn, kk = n_off + n, k_off + k
                if layout == 0:
                    dst[nn, tap, kk] = v
                else:
                    dst[tap, kk, nn] = v
    return dst
