#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_i3d.py -m gpu -q --maxfail=5 -x > gpurun_out/gputests_pdl.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests_pdl.log
for v in 1 0; do IVF_PDL=$v python bench.py --steps 200 --no-cpu --no-gradcam --no-clstm --clips-per-gpu 32 > gpurun_out/bench_pdl$v.log 2>&1; echo "rc=$?" >> gpurun_out/bench_pdl$v.log; done
tail -4 gpurun_out/gputests_pdl.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_pdl*.log")):
    ok=False
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); ok=True; print(f, "value %.1f ms %.4f e2e %.1f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]))
    if not ok: print(f, open(f).read()[-800:])
PY
