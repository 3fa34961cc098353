#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_drivers.py tests/test_gpu_clstm.py -m gpu -q --maxfail=20 -s > gpurun_out/gputests4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gputests4.log
python bench.py > gpurun_out/bench_full.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_full.log
tail -12 gpurun_out/gputests4.log; grep -n "clip [0-9]: ours" gpurun_out/gputests4.log; tail -3 gpurun_out/bench_full.log | cut -c1-6000
