#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/gputests8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests8.log
python bench.py > gpurun_out/bench_full2.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_full2.log
tail -4 gpurun_out/gputests8.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_full2.log"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.1f ms %.4f e2e %.1f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), "frac", r["frac"], "mb32", r.get("at_e2e_micro_batch"))
        print("gradcam", d["gradcam"]["value"], d["gradcam"]["e2e"], "clstm", d["clstm"]["value"], "cpu", d["cpu_baseline"]["value"])
PY
tail -2 gpurun_out/bench_full2.log | cut -c1-200
