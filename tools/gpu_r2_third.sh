#!/bin/bash
mkdir -p gpurun_out
python tools/debug_traj.py > gpurun_out/debug_traj.log 2>&1
python -m pytest tests -m gpu -q --maxfail=20 -s --durations=15 > gpurun_out/gputests3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests3.log
python bench.py --steps 30 --warmup 3 --clips-per-gpu 128 --e2e-micro-batch 64 --no-cpu --no-clstm > gpurun_out/bench_mb64.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_mb64.log
cat gpurun_out/debug_traj.log; tail -15 gpurun_out/gputests3.log; tail -2 gpurun_out/bench_mb64.log | cut -c1-2500
