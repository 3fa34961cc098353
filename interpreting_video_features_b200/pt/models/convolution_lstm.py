"""Drop-in for the reference's models.convolution_lstm (pt/models/convolution_lstm.py): ConvLSTMCell and
ConvLSTM with the same constructor arguments, parameter names (cell{i}.W{x,h}{i,f,c,o}, bn) and creation
order, so checkpoints and seeded initialisation are interchangeable.  CLSTM_4.Model runs the whole classifier
on engine_clstm.CLSTMEngine (one autograd node); called on their own, `ConvLSTM.forward` runs the recurrent
stack on the same engine and `ConvLSTMCell.forward` one cell step with libivf kernels - both forward-only
(eval mode, no autograd), for code that walks the children (pt/pytorch-grad-cam/grad-cam.py:23-54)."""
import torch
import torch.nn as nn

from ... import _lib, ops
from ...ops import Act


class ConvLSTMCell(nn.Module):
    def __init__(self, input_channels, hidden_channels, kernel_size, conv_stride, device):
        super().__init__()
        assert hidden_channels % 2 == 0
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.num_features = 4
        self.conv_stride = conv_stride
        self.device = device
        self.padding = int((kernel_size - 1) / 2)
        k, s, p = kernel_size, conv_stride, self.padding
        for g in ("i", "f", "c", "o"):  # creation order Wxi, Whi, Wxf, Whf, ... as the reference (:25-32)
            setattr(self, "Wx" + g, nn.Conv2d(input_channels, hidden_channels, k, s, p, bias=True))
            setattr(self, "Wh" + g, nn.Conv2d(hidden_channels, hidden_channels, k, 1, p, bias=False))
        self.Wci = self.Wcf = self.Wco = None  # zero peepholes, created lazily by the reference (:50-54)

    def _packed(self, device):
        """fp32 operands of the two gate convolutions (four gates side by side), repacked when a weight changes."""
        from ...engine import _dev32, pack
        ver = tuple(p._version for p in self.parameters())
        hit = getattr(self, "_pack_cache", None)
        if hit is None or hit[0] != ver or hit[1] != str(device):
            hid = self.hidden_channels
            wx = [_dev32(getattr(self, "Wx" + g).weight, device) for g in "ifco"]
            wh = [_dev32(getattr(self, "Wh" + g).weight, device) for g in "ifco"]
            offs = [gi * hid for gi in range(4)]
            bias = torch.cat([getattr(self, "Wx" + g).bias.detach().float() for g in "ifco"]).to(device).contiguous()
            ones = torch.ones(4 * hid, device=device)
            hit = (ver, str(device), pack(wx, "fp32", co_offs=offs, co_total=4 * hid),
                   pack(wh, "fp32", co_offs=offs, co_total=4 * hid), bias, ones, (wx, wh))
            self._pack_cache = hit
        return hit[2:6]

    def forward(self, x, h, c):
        """One cell step (pt/models/convolution_lstm.py:38-48): returns (h', c') as fp32 NCHW tensors.  Native and
        forward-only: x-convolution (stride conv_stride, bias) and h-convolution accumulated by the conv epilogue,
        then the fused gate kernel (zero peepholes, :50-54)."""
        if not x.is_cuda:
            raise _lib.IvfError("ConvLSTMCell: input is on %s; the native path runs on a B200 only" % x.device)
        if torch.is_grad_enabled() and (x.requires_grad or h.requires_grad or c.requires_grad):
            raise _lib.IvfError("ConvLSTMCell.forward is forward-only; gradients flow through CLSTM_4.Model")
        dev = x.device
        wx, wh, bias, ones = self._packed(dev)
        n, cin, hh, ww = x.shape
        k, s, p, hid = self.kernel_size, self.conv_stride, self.padding, self.hidden_channels
        ho, wo = (hh + 2 * p - k) // s + 1, (ww + 2 * p - k) // s + 1

        def cl(t):  # NCHW -> channels-last Act with depth 1
            b = t.detach().float().permute(0, 2, 3, 1).contiguous()
            return Act(b, b.shape[0], 1, b.shape[1], b.shape[2], b.shape[3], 0, b.shape[3])

        pre = Act.empty(n, 1, ho, wo, 4 * hid, torch.float32, dev)
        ops.conv3d(cl(x), wx, pre, (1, k, k), (1, s, s), (0, p, p), scale=ones, shift=bias)
        ops.conv3d(cl(h), wh, pre, (1, k, k), (1, 1, 1), (0, p, p), acc_in=pre)
        m = n * ho * wo
        c_next = torch.empty((m, hid), dtype=torch.float32, device=dev)
        h_next = torch.empty((m, hid), dtype=torch.float32, device=dev)
        c_prev = c.detach().float().permute(0, 2, 3, 1).contiguous().view(m, hid)
        ops.clstm_gates_fwd(pre.buf.view(m, 4 * hid), c_prev, c_next, h_next, None)

        def nchw(t):
            return t.view(n, ho, wo, hid).permute(0, 3, 1, 2).contiguous()

        return nchw(h_next), nchw(c_next)

    def init_hidden(self, batch_size, hidden, shape):
        """Zero state at the cell's output resolution (pt/models/convolution_lstm.py:56-60)."""
        dev = self.Wxi.weight.device
        size = (batch_size, hidden, shape[0] // self.conv_stride, shape[1] // self.conv_stride)
        return torch.zeros(size, device=dev), torch.zeros(size, device=dev)


class ConvLSTM(nn.Module):
    def __init__(self, input_channels, hidden_channels, kernel_size, conv_stride, pool_kernel_size=(2, 2), step=1,
                 effective_step=[1], batch_normalization=True, dropout=0, device='cpu'):
        super().__init__()
        self.input_channels = [input_channels] + hidden_channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.num_layers = len(hidden_channels)
        self.step = step
        self.effective_step = effective_step
        self._all_layers = []
        self.pool_kernel_size = pool_kernel_size
        self.conv_stride = conv_stride
        self.mp = nn.MaxPool2d(kernel_size=self.pool_kernel_size)
        self.batch_norm = batch_normalization
        self.dropout_rate = dropout
        self.device = device
        self.bn = nn.BatchNorm2d(self.hidden_channels[0], eps=1e-05, momentum=0.1, affine=True)
        self.dropout = torch.nn.Dropout(p=self.dropout_rate)
        for i in range(self.num_layers):
            cell = ConvLSTMCell(self.input_channels[i], self.hidden_channels[i], self.kernel_size, self.conv_stride,
                                self.device)
            setattr(self, 'cell{}'.format(i), cell)
            self._all_layers.append(cell)

    def forward(self, input):
        """pt/models/convolution_lstm.py:96-132 on the native engine, forward-only (eval: dropout off, BatchNorm
        running statistics): returns (outputs at the effective steps, (x, new_c)) as the reference does."""
        from ...engine_clstm import CLSTMEngine
        from ._i3d_native import default_mode
        if not input.is_cuda:
            raise _lib.IvfError("ConvLSTM: input is on %s; the native path runs on a B200 only" % input.device)
        if self.training:
            raise _lib.IvfError("the native ConvLSTM runs eval mode only (dropout off, BatchNorm running stats)")
        if torch.is_grad_enabled() and input.requires_grad:
            raise _lib.IvfError("ConvLSTM.forward is forward-only; gradients flow through CLSTM_4.Model")
        if len(set(self.hidden_channels)) != 1:
            raise _lib.IvfError("the native ConvLSTM needs the same number of hidden channels in every layer")
        b, c, t, h, w = input.shape
        if t < self.step:
            raise _lib.IvfError("clip has %d frames, the stack runs %d steps" % (t, self.step))
        mode = getattr(self, "ivf_mode", None) or default_mode()
        ver = tuple(p._version for p in self.parameters()) + tuple(bf._version for bf in self.buffers())
        key = (b, c, self.step, h, w, mode, str(input.device))
        hit = getattr(self, "_engine_cache", None)
        if hit is None or hit[0] != key or hit[1] != ver:
            sd = {"clstm." + k: v for k, v in self.state_dict().items()}
            stride = self.conv_stride if isinstance(self.conv_stride, int) else self.conv_stride[0]
            eng = CLSTMEngine(sd, b, (self.step, h, w), self.hidden_channels[0], self.num_layers, None,
                              kernel=self.kernel_size, conv_stride=stride, mode=mode, batch_norm=self.batch_norm,
                              effective_step=self.effective_step, device=input.device, in_channels=c)
            hit = (key, ver, eng)
            self._engine_cache = hit
        eng = hit[2]
        eng.set_input(input.detach()[:, :, :self.step].contiguous())
        eng.forward(None, "freeze")
        return eng.step_outputs()
