"""Drop-in for the reference's models.convolution_lstm (pt/models/convolution_lstm.py): ConvLSTMCell and
ConvLSTM with the same constructor arguments, parameter names (cell{i}.W{x,h}{i,f,c,o}, bn) and creation
order, so checkpoints and seeded initialisation are interchangeable.  The computation is the native
engine's (engine_clstm.CLSTMEngine); these modules hold the parameters and expose a forward-only native
`ConvLSTM.forward` for code that walks the children (pt/pytorch-grad-cam/grad-cam.py:23-54)."""
import torch
import torch.nn as nn

from ... import _lib


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

    def forward(self, x, h, c):
        raise _lib.IvfError("ConvLSTMCell has no stand-alone native kernel: the gates are fused into the "
                            "ConvLSTM schedule (call ConvLSTM / CLSTM_4.Model)")


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
