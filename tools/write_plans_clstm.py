"""Measure the slab-kernel tile plans of the ConvLSTM (config C3, 8 and 16 clips) recurrent convolutions on this GPU
and write them to gpurun_out/plans_clstm.json (merge the new keys into plans_sm100.json to adopt them).
GPU only:  python tools/write_plans_clstm.py"""
import contextlib
import io
import json
import os
import sys

os.environ["IVF_TUNE"] = "force"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpreting_video_features_b200 import tune  # noqa: E402
from interpreting_video_features_b200.pt.models import CLSTM_4  # noqa: E402
from oracle import synthetic  # noqa: E402

dev = torch.device("cuda:0")
ROUNDS = int(os.environ.get("IVF_PLAN_ROUNDS", "3"))
votes = {}
for r in range(ROUNDS):
    for clips_n in (8, 16, 32):
        tune._CACHE.clear()
        tune.MEASURED.clear()
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            m = CLSTM_4.Model(num_classes=6, nb_lstm_units=32, channels=3, conv_kernel_size=(5, 5), lstm_layers=2,
                              step=32, conv_stride=2, image_size=(160, 120), effective_step=[7, 15, 23, 31],
                              batch_normalization=True, dropout=0.5, add_softmax=True).to(dev).eval().set_mode("bf16")
        x = torch.stack([synthetic.uniform_clip_u8(2000 + i, t=32, h=120, w=160) for i in range(clips_n)]).to(dev).float()
        m._engine(x, batch=clips_n)
        torch.cuda.synchronize()
        for k, v in tune.MEASURED.items():
            votes.setdefault(k, []).append(None if v is None else tuple(v))
        del m
        torch.cuda.empty_cache()
plans = {}
for k, vs in votes.items():
    best = max(set(vs), key=vs.count)
    plans[k] = None if best is None else list(best)
    print(k, vs, "->", best)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"plans": plans}, open("gpurun_out/plans_clstm.json", "w"), indent=1, sort_keys=True)
print("wrote %d plans" % len(plans))
