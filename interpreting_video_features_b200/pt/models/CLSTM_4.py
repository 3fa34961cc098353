"""Drop-in for the reference's models.CLSTM_4 (pt/models/CLSTM_4.py): same Model signature, attributes and
state-dict keys (clstm.cell{i}.W{x,h}{i,f,c,o}.{weight,bias}, clstm.bn.*, endFC.*); forward/backward run
on the native ConvLSTM engine (one autograd node).  CUDA only, eval mode only."""
import torch

from ... import _lib, ops
from ...engine_clstm import CLSTMEngine
from ._i3d_native import default_mode
from .convolution_lstm import ConvLSTM


class _ClstmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model):
        eng = model._engine(x)
        eng.set_input(x.detach().contiguous())
        out = eng.forward(None, "freeze").clone()  # zero mask: the operand is the clip itself
        ctx.eng, ctx.gen, ctx.shape = eng, eng.generation, x.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        eng = ctx.eng
        if eng.generation != ctx.gen:
            raise _lib.IvfError("backward through a forward whose activations were overwritten by a later forward")
        eng.dprobs.copy_(gout)
        eng.generation += 1
        eng.backward(to_mask=False)
        b, c, t, h, w = ctx.shape
        g = eng.g_xin.buf
        if eng.mode == "bf16":  # [t*b][h/2][w/2][16] space-to-depth records -> NCDHW
            g = g.view(t, b, h // 2, w // 2, 16)[..., :4 * c].float().view(t, b, h // 2, w // 2, 2, 2, c)
            g = g.permute(1, 6, 0, 2, 4, 3, 5).reshape(b, c, t, h, w)
        else:
            g = g.view(t, b, h, w, c).permute(1, 4, 0, 2, 3).contiguous()
        return g, None


class Model(torch.nn.Module):
    def __init__(self, num_classes=174, nb_lstm_units=32, channels=3, conv_kernel_size=(5, 5), pool_kernel_size=(2, 2),
                 top_layer=True, avg_pool=False, batch_normalization=True, lstm_layers=4, step=16,
                 image_size=(224, 224), dropout=0, conv_stride=(1, 1), effective_step=[4, 8, 12, 15],
                 use_entire_seq=False, add_softmax=False):
        super().__init__()
        self.num_classes = num_classes
        self.nb_lstm_units = nb_lstm_units
        self.channels = channels
        self.top_layer = top_layer
        self.avg_pool = avg_pool
        self.c_kernel_size = conv_kernel_size
        self.lstm_layers = lstm_layers
        self.step = step
        self.im_size = image_size
        self.pool_kernel_size = pool_kernel_size
        self.batch_normalization = batch_normalization
        self.dropout = dropout
        self.conv_stride = conv_stride
        self.effective_step = effective_step
        self.add_softmax = add_softmax
        self.use_entire_seq = use_entire_seq
        self.clstm = None
        self.endFC = None
        self.sm = None
        self.ivf_mode = default_mode()
        self._engines = {}
        self.build()

    def _features(self):
        red = (self.conv_stride * self.pool_kernel_size[0]) ** self.lstm_layers
        return self.nb_lstm_units * int(self.im_size[0] / red) * int(self.im_size[1] / red)

    def build(self):
        self.clstm = ConvLSTM(input_channels=self.channels, hidden_channels=[self.nb_lstm_units] * self.lstm_layers,
                              kernel_size=self.c_kernel_size[0], conv_stride=self.conv_stride,
                              pool_kernel_size=self.pool_kernel_size, step=self.step,
                              effective_step=self.effective_step, batch_normalization=self.batch_normalization,
                              dropout=self.dropout)
        n_in = self._features() * (len(self.effective_step) if self.use_entire_seq else 1)
        self.endFC = torch.nn.Linear(in_features=n_in, out_features=self.num_classes)
        print("use entire sequence is: ", self.use_entire_seq)
        print("shape of FC is: ", self.endFC)
        self.sm = torch.nn.Softmax(dim=1)

    def set_mode(self, mode):
        assert mode in ("bf16", "fp32")
        self.ivf_mode = mode
        return self

    def _engine(self, x, batch=None):
        b, c, t, h, w = x.shape
        device = next(self.parameters()).device
        if device.type != "cuda":
            raise _lib.IvfError("the native ConvLSTM needs the model on a CUDA device (model.cuda())")
        if self.training:
            raise _lib.IvfError("the native ConvLSTM runs eval mode only (dropout off, BatchNorm running stats)")
        if self.use_entire_seq:
            raise _lib.IvfError("use_entire_seq=True is not supported natively (the reference's view() mixes batch "
                                "and steps there for batch > 1)")
        if t != self.step:
            raise _lib.IvfError("clip has %d frames, the model was built for step=%d" % (t, self.step))
        key = (batch or b, c, t, h, w, self.ivf_mode, str(device))
        ver = tuple(p._version for p in self.parameters()) + tuple(bf._version for bf in self.buffers())
        hit = self._engines.get(key)
        if hit is None or hit[0] != ver:
            stride = self.conv_stride if isinstance(self.conv_stride, int) else self.conv_stride[0]
            eng = CLSTMEngine(self.state_dict(), batch or b, (t, h, w), self.nb_lstm_units, self.lstm_layers,
                              self.num_classes, kernel=self.c_kernel_size[0], conv_stride=stride, mode=self.ivf_mode,
                              softmax=bool(self.add_softmax), batch_norm=self.batch_normalization,
                              effective_step=self.effective_step, device=device, in_channels=c)
            hit = (ver, eng)
            self._engines = {key: hit}
        return hit[1]

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.IvfError("CLSTM_4.Model: input is on %s; the native path runs on a B200 only" % x.device)
        return _ClstmFn.apply(x, self)

    def native_gradcam(self, x, indices, input_spatial_size, per_frame):
        """Grad-CAM on the stacked effective-step outputs (pt/grad_cam_videos.py:87-140 for archType 'CLSTM',
        the layer the TF tree and the paper use); returns (cam [B,T,H,W] device tensor, output [B,classes])."""
        eng = self._engine(x)
        eng.set_input(x)
        keep, eng.softmax = eng.softmax, True  # the walked children include `sm`: the score is the softmax prob
        try:
            out = eng.forward(None, "freeze").clone()
            tg = torch.argmax(out, dim=1) if indices is None else torch.as_tensor(indices, device=out.device)
            eng.set_targets(tg)
            act, grad = eng.gradcam_operands()
        finally:
            eng.softmax = keep
        step = x.shape[2] // act.d
        w_out, h_out = input_spatial_size
        cam = torch.empty((x.shape[0], act.d * step, h_out, w_out), dtype=torch.float32, device=x.device)
        ops.gradcam(act, grad, step, h_out, w_out, per_frame, cam)
        return cam, out
