"""Oracle (test infrastructure): ConvLSTM classifier restated functionally from
pt/models/convolution_lstm.py:38-132 and pt/models/CLSTM_4.py:69-85 over a reference-keyed state
dict (clstm.cell{i}.W{x,h}{i,f,c,o}.{weight,bias}, clstm.bn.*, endFC.*).  Eval mode: dropout is
the identity, the ONE BatchNorm2d shared by all layers/steps uses running stats (eps 1e-5).
"""
import torch
import torch.nn.functional as F

from .i3d_oracle import _RoundGrad, _q, quant_input


def _cell(sd, p, x, h, c, k, stride, quant=False):
    pad = (k - 1) // 2
    wq = (lambda w: w.bfloat16().float()) if quant else (lambda w: w)

    def gate(g):
        pre = (F.conv2d(x, wq(sd[p + ".Wx%s.weight" % g]), sd[p + ".Wx%s.bias" % g], stride=stride, padding=pad)
               + F.conv2d(h, wq(sd[p + ".Wh%s.weight" % g]), None, stride=1, padding=pad))
        return _RoundGrad.apply(pre) if quant else pre  # the stored gate-pre-activation gradient is bf16

    ci = torch.sigmoid(gate("i"))  # + c * Wci with Wci == 0 (convolution_lstm.py:50-54)
    cf = torch.sigmoid(gate("f"))
    cc = cf * c + ci * torch.tanh(gate("c"))
    co = torch.sigmoid(gate("o"))
    return co * torch.tanh(cc), cc


def forward(sd, x, num_layers, hidden, kernel=5, conv_stride=2, step=None, effective_step=(7, 15, 23, 31),
            batch_norm=True, softmax=False, use_entire_seq=False, return_outputs=False, quant=False,
            force_argmax=None):
    """x [B,C,T,H,W] -> logits/probs [B,classes] (pt/models/CLSTM_4.py:69-85).
    quant=True: the same network with the bf16 rounding points of the tensor-core path (bf16 conv
    weights, bf16-stored clip / hidden states / pooled maps and their stored gradients; fp32 cell
    state, gate math and accumulation) — see oracle/i3d_oracle.py.
    force_argmax: optional per-layer list of IMPOSED 2x2 max-pool routings, int tensors [T,B,hid,h/2,w/2] holding
    the window element (row*2 + col) each pooled value is taken from (instead of its own arg-maximum): the
    max-pools are the only non-smooth operations of this network, so two evaluations that agree on the routing
    agree on the gradient up to rounding."""
    B = x.shape[0]
    if quant:
        x = quant_input(x)
    step = x.shape[2] if step is None else step
    state = [None] * num_layers
    outputs = []
    for t in range(step):
        cur = x[:, :, t]
        for i in range(num_layers):
            p = "clstm.cell%d" % i
            if state[i] is None:
                hh, ww = cur.shape[2] // conv_stride, cur.shape[3] // conv_stride
                z = torch.zeros(B, hidden, hh, ww, dtype=x.dtype, device=x.device)
                state[i] = (z, z)
            h, c = state[i]
            cur, new_c = _cell(sd, p, cur, h, c, kernel, conv_stride, quant)
            if quant:
                cur = _q(cur)  # h is stored in bf16 (next step's h-conv operand and the pool input)
            state[i] = (cur, new_c)
            if batch_norm:
                cur = F.batch_norm(cur, sd["clstm.bn.running_mean"], sd["clstm.bn.running_var"],
                                   sd["clstm.bn.weight"], sd["clstm.bn.bias"], training=False, eps=1e-5)
            if force_argmax is not None:
                bb, cc, hh2, ww2 = cur.shape
                ev = cur[:, :, :hh2 // 2 * 2, :ww2 // 2 * 2]  # MaxPool2d(2) floors odd maps (15x20 -> 7x10)
                win = ev.reshape(bb, cc, hh2 // 2, 2, ww2 // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(bb, cc, hh2 // 2, ww2 // 2, 4)
                cur = win.gather(-1, force_argmax[i][t].long().unsqueeze(-1)).squeeze(-1)
            else:
                cur = F.max_pool2d(cur, 2)
            if quant:
                cur = _q(_RoundGrad.apply(cur))
        if t in effective_step:
            outputs.append(cur)
    if use_entire_seq:
        flat = torch.stack(outputs).reshape(-1, len(effective_step) * outputs[-1][0].numel())
    else:
        flat = outputs[-1].reshape(B, -1)
    out = F.linear(flat, sd["endFC.weight"], sd["endFC.bias"])
    if softmax:
        out = F.softmax(out, dim=1)
    return (out, outputs) if return_outputs else out


class Model:
    def __init__(self, sd, **kw):
        self.sd, self.kw = sd, kw

    def __call__(self, x):
        return forward(self.sd, x, **self.kw)
