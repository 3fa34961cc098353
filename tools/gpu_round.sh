#!/bin/bash
# One gpurun call: the driver's GPU test command, smoke, bench, then the ncu launch list of one iteration.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log | cut -c1-600; }
TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu
TMO=300 run t_smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=900 run bench python bench.py --steps 20 --warmup 3
TMO=300 run prof_plain python tools/profile_step.py
if grep -q "exit 0" gpurun_out/prof_plain.log; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_r01.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
  echo "ncu exit $?"
fi
