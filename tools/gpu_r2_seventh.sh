#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_clstm.py tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q --maxfail=20 -k "clstm or pack or conv" > gpurun_out/gputests7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests7.log
for f in 1 0; do IVF_CLSTM_FUSED=$f python bench.py --steps 20 --no-cpu --no-gradcam --clips-per-gpu 8 --e2e-micro-batch 8 > gpurun_out/bench_clstm_f$f.log 2>&1; done
for p in "0,0,0" "-1,0,0" "-1,-1,0" "0,0,-1"; do IVF_LANE_PRIO=$p python bench.py --steps 100 --no-cpu --no-gradcam --no-clstm --clips-per-gpu 8 --e2e-micro-batch 8 > gpurun_out/bench_prio_$p.log 2>&1; done
tail -8 gpurun_out/gputests7.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_clstm_f*.log"))+sorted(glob.glob("gpurun_out/bench_prio_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); print(f, "value %.1f ms %.4f"%(d["value"], d["ms_per_step"]), (d.get("clstm") or {}).get("ms_per_step"), (d.get("clstm") or {}).get("launches_per_step"))
PY
