"""One eager ConvLSTM mask-search iteration (config C3) under the profiler:
ncu --profile-from-start off ... python tools/profile_clstm.py"""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpreting_video_features_b200 import ops, search  # noqa: E402
from interpreting_video_features_b200.pt.models import CLSTM_4  # noqa: E402
from oracle import synthetic  # noqa: E402

dev = torch.device("cuda:0")
clips_n = 8
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = CLSTM_4.Model(num_classes=6, nb_lstm_units=32, channels=3, conv_kernel_size=(5, 5), lstm_layers=2, step=32,
                      conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31], batch_normalization=True,
                      dropout=0.5, add_softmax=True).to(dev).eval().set_mode("bf16")
x = torch.stack([synthetic.uniform_clip(2000 + i, t=32, h=120, w=160) for i in range(clips_n)]).to(dev) / 255.0
eng = m._engine(x, batch=clips_n)
ms = search.MaskSearch(eng, 0.02, 0.04, 0.2, 100, "reverse", 0.9, use_graph=False)
ms.set_input(x)
ms.set_targets((torch.arange(clips_n) % 6).to(dev))
ms.m.copy_(torch.tensor([-5.] * 8 + [5.] * 16 + [-5.] * 8, device=dev).repeat(clips_n, 1))
ops.sigmoid(ms.m, ms.sig)
for _ in range(2):
    ms._iteration()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ms._iteration()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one ConvLSTM iteration for %d clips" % clips_n)
