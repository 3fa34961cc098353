#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, optional extras.  Usage: tools/gpu_round.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-700; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu --deselect tests/test_gpu_clstm.py::test_forward_and_mask_gradient
TMO=900 run bench python bench.py --steps 20 --warmup 3
TMO=300 run dbg_clstm python tools/debug_clstm.py
