#!/bin/bash
# One gpurun call for kernel work: $TESTS (pytest -k expression), bench, then ONE ncu pass:
#   NCU=list  -> launch list with a few metrics for every kernel of one iteration
#   NCU=full  -> --set full with source for the kernels matching $KREGEX (keep the count small: 64 MiB cap)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n 15 gpurun_out/$name.log | cut -c1-900; return $rc; }
if [ -n "$TESTS" ]; then TMO=900 run t_kernels python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "$TESTS" || exit 1; fi
if [ -z "$SKIP_ALL" ]; then TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu; fi
TMO=900 run bench python bench.py --steps 50 --warmup 5 ${BENCH_ARGS}
TMO=300 run prof_plain python tools/profile_step.py
grep -q "exit 0" gpurun_out/prof_plain.log || exit 1
if [ "${NCU:-list}" = "list" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
else
  timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:$KREGEX" ${NCU_COUNT:+-c $NCU_COUNT} \
      -o gpurun_out/full_${NCU_TAG:-k} -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
fi
echo "ncu exit $?"; ls -la gpurun_out | tail -n 12
