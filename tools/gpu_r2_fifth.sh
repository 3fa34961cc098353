#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 -s -k "c1_kth or other_target" > gpurun_out/gputests5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests5.log
python bench.py --steps 20 --no-cpu --no-clstm --clips-per-gpu 32 > gpurun_out/bench5.log 2>&1
tail -8 gpurun_out/gputests5.log; grep -n "^E  \|C1 " gpurun_out/gputests5.log | head -20
python - <<'PY'
import json
for l in open("gpurun_out/bench5.log"):
    if l.startswith("{"):
        d=json.loads(l); g=d["gradcam"]; print("gradcam value",g["value"],"e2e",g["e2e"],"224:",g["geometry_224"]["value"])
PY
tail -3 gpurun_out/bench5.log | cut -c1-300
