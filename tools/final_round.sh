#!/bin/bash
# End-of-round evidence: GPU tests, smoke, default bench (with the CPU reference leg), launch list, ConvLSTM launch list.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-600; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu
TMO=600 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=1500 run bench python bench.py
TMO=300 run prof_plain python tools/profile_step.py
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
