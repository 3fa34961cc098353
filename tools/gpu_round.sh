#!/bin/bash
# One gpurun call: slab bring-up (which descriptor mode works), GPU tests, bench, diagnostics.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-700; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
SLABTESTS="tests/test_gpu_kernels.py -k conv"
if IVF_SLAB_BASEOFF=0 TMO=600 run t_slab_b0 python -m pytest $SLABTESTS -q -m gpu; then export IVF_SLAB_BASEOFF=0
elif IVF_SLAB_BASEOFF=1 TMO=600 run t_slab_b1 python -m pytest $SLABTESTS -q -m gpu; then export IVF_SLAB_BASEOFF=1
else export IVF_SLAB=0; fi
echo "chosen: IVF_SLAB=$IVF_SLAB IVF_SLAB_BASEOFF=$IVF_SLAB_BASEOFF" | tee gpurun_out/slab_mode.txt
TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu --deselect tests/test_gpu_clstm.py::test_forward_and_mask_gradient
TMO=900 run bench python bench.py --steps 20 --warmup 3
TMO=300 run dbg_clstm python tools/debug_clstm.py
TMO=300 run prof_plain python tools/profile_step.py
if grep -q "exit 0" gpurun_out/prof_plain.log; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
  echo "ncu exit $?"
fi
