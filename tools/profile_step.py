"""One eager (no CUDA graph) mask-search iteration for 8 clips under the profiler:
ncu --profile-from-start off ... python tools/profile_step.py      (cudaProfilerStart/Stop bracket it)"""
import contextlib
import io
import os
import sys

import torch

os.environ.setdefault("IVF_STREAMS", "0")  # one stream: the launch list is serialised anyway

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CLIPS, NCLS, state_dict  # noqa: E402
from interpreting_video_features_b200 import ops, search  # noqa: E402
from oracle import synthetic  # noqa: E402

clips_n = int(os.environ.get("IVF_PROFILE_CLIPS", CLIPS))
dev = torch.device("cuda:0")
model = state_dict().to(dev).eval().set_mode("bf16")
clips = torch.stack([synthetic.uniform_clip(i) for i in range(clips_n)])
groups = int(os.environ.get("IVF_PROFILE_GROUPS", "1"))  # serialised launch list: one group by default
ms = search.MaskSearch(search.make_engines(model, clips, clips_n, groups), use_graph=False)
ms.set_input(clips.to(dev))
ms.set_targets((torch.arange(clips_n) % NCLS).to(dev))
ms.m.copy_(torch.tensor([-5.] * 4 + [5.] * 8 + [-5.] * 4, device=dev).repeat(clips_n, 1))
ops.sigmoid(ms.m, ms.sig)
for _ in range(2):
    ms._iteration()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ms._iteration()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one iteration for %d clips" % clips_n)

# ---- program structure (lanes, forks, joins, launches per op) for critical-path analysis of the launch list
import json  # noqa: E402
from interpreting_video_features_b200 import _lib  # noqa: E402
eng = ms.engs[0]
prog = []
for name, ops_list in (("fwd", eng.fwd_ops), ("bwd", eng.bwd_ops)):
    for item in ops_list:
        if isinstance(item[0], int):
            n0 = _lib.launch_count(dev)
            item[1]()
            prog.append([name, "op", item[0], _lib.launch_count(dev) - n0])
        else:
            prog.append([name] + [list(v) if isinstance(v, tuple) else v for v in item])
torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(prog, open("gpurun_out/program.json", "w"))
