#!/bin/bash
# One gpurun call: GPU tests, bench, launch list (ncu) of one iteration.  Extra diagnostics via $EXTRA.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-700; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
TMO=600 run t_conv python -m pytest tests/test_gpu_kernels.py -k conv -q -m gpu -x || exit 1
IVF_SLAB_2CTA=0 TMO=300 run t_conv_1cta python -m pytest tests/test_gpu_kernels.py -k conv -q -m gpu -x
TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu
TMO=900 run bench python bench.py --steps 20 --warmup 3
[ -n "$EXTRA" ] && TMO=600 run extra bash -c "$EXTRA"
TMO=300 run prof_plain python tools/profile_step.py
if grep -q "exit 0" gpurun_out/prof_plain.log; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
  echo "ncu exit $?"
fi
