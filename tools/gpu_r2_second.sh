#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 -s > gpurun_out/gputests2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests2.log
for mb in 8 16 32; do
  python bench.py --steps 30 --warmup 3 --clips-per-gpu 64 --e2e-micro-batch $mb --no-cpu --no-gradcam --no-clstm > gpurun_out/bench_mb$mb.log 2>&1
  echo "bench mb=$mb rc=$?" >> gpurun_out/bench_mb$mb.log
done
tail -12 gpurun_out/gputests2.log
for mb in 8 16 32; do python - <<PY
import json
for l in open("gpurun_out/bench_mb$mb.log"):
    if l.startswith("{"):
        d=json.loads(l); print("mb", $mb, "value", d["value"], "e2e", d["e2e"]["value"], "sec", d["e2e"]["seconds_per_job"])
PY
tail -2 gpurun_out/bench_mb$mb.log | cut -c1-300; done
