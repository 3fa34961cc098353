#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 --durations=12 > gpurun_out/gputests6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests6.log
python tools/sweep_c5.py --clips 2 --iters 10 > gpurun_out/c5_sweep.json 2> gpurun_out/c5_sweep.err
echo "sweep rc=$?" >> gpurun_out/c5_sweep.err
tail -25 gpurun_out/gputests6.log; tail -3 gpurun_out/c5_sweep.err; grep -n "violations" -A5 gpurun_out/c5_sweep.json
