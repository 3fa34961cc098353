#!/bin/bash
# round 2, first GPU call: full GPU suite (with durations), smoke, short bench, launch list of smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -q --maxfail=10 --durations=30 -s > gpurun_out/gputests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 50 --warmup 3 --clips-per-gpu 16 > gpurun_out/bench_short.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_short.log
tail -5 gpurun_out/gputests.log; tail -3 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench_short.log
