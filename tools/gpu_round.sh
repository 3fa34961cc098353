#!/bin/bash
# One gpurun call: kernel probes, kernel parity, end-to-end parity, a first bench; logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
TMO=1800 run t_i3d_fp32 python -m pytest tests/test_gpu_i3d.py -m gpu -q -k "fp32 or dropin_reference"
TMO=1800 run t_i3d_bf16 python -m pytest tests/test_gpu_i3d.py -m gpu -q -k "not fp32 and not dropin_reference"
TMO=300 run t_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run bench python bench.py --steps 20 --warmup 3
