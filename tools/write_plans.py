"""Measure the slab-kernel tile plans for the shapes of the shipped configurations on this GPU and write them
to interpreting_video_features_b200/plans_sm100.json format (gpurun_out/plans_sm100.json; copy it next to
tune.py to adopt it).  GPU only:  IVF_TUNE=force python tools/write_plans.py"""
import json
import os
import sys

os.environ["IVF_TUNE"] = "force"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CLIPS, state_dict  # noqa: E402
from interpreting_video_features_b200 import search, tune  # noqa: E402
from oracle import synthetic  # noqa: E402

dev = torch.device("cuda:0")
ROUNDS = int(os.environ.get("IVF_PLAN_ROUNDS", "3"))
votes = {}
for r in range(ROUNDS):
    tune._CACHE.clear()
    tune.MEASURED.clear()
    # C2 / C4: I3D smth, 8 clips of 16 x 224 x 224 per micro-batch
    model = state_dict().to(dev).eval().set_mode("bf16")
    clips = torch.stack([synthetic.uniform_clip(i) for i in range(CLIPS)])
    search.MaskSearch(search.make_engines(model, clips, CLIPS, 1), use_graph=False)
    # C1: I3D KTH Grad-CAM, 8 clips of 32 x 120 x 160
    from interpreting_video_features_b200.pt.grad_cam_videos import GradCamVideo
    from interpreting_video_features_b200.pt.models import I3D_doubled_kth
    m = I3D_doubled_kth.Model(6, last_stride=1, stride_mod_layers="", softMax=1, finalTimeLength=4)
    m = m.to(dev).eval().set_mode("bf16")
    gc = GradCamVideo(model=m, target_layer_names=['Mixed_5c'], class_dict=None, use_cuda=True,
                      input_spatial_size=(160, 120), normalizePerFrame=True, archType="I3D")
    x = torch.stack([synthetic.uniform_clip(1000 + i, t=32, h=120, w=160) for i in range(8)]).to(dev)
    gc._i3d(x, [i % 6 for i in range(8)])
    torch.cuda.synchronize()
    for k, v in tune.MEASURED.items():
        votes.setdefault(k, []).append(None if v is None else tuple(v))
    del model, m, gc
    torch.cuda.empty_cache()
plans = {}
for k, vs in votes.items():
    best = max(set(vs), key=vs.count)  # the plan most rounds agree on
    plans[k] = None if best is None else list(best)
    print(k, vs, "->", best)
os.makedirs("gpurun_out", exist_ok=True)
out = {"device": torch.cuda.get_device_name(0), "fields": list(tune._FIELDS),
       "request": ["kwm", "mt", "acc", "ncta", "ntiles", "ds"], "plans": plans}
json.dump(out, open("gpurun_out/plans_sm100.json", "w"), indent=1, sort_keys=True)
print("wrote %d plans" % len(plans))
