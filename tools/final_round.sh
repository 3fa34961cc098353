#!/bin/bash
# End-of-round evidence: GPU tests, smoke (plain and as a launch list), default bench (with the CPU reference leg),
# launch lists of one iteration at 8 and 64 clips with a few metrics, ncu --set full of the first convolution launches,
# in-situ per-launch timings, ConvLSTM launch list, launch list + ncu --set full (weight gradient) of a training step.
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name" ; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log | cut -c1-400; return $rc; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
TMO=2400 run t_gpu_all python -m pytest tests -q -m gpu
TMO=600 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=1500 run bench python bench.py
TMO=300 run prof_plain python tools/profile_step.py
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "ncu list exit $?"
IVF_PROFILE_CLIPS=64 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches64.csv python tools/profile_step.py > gpurun_out/ncu_launches64.log 2>&1
echo "ncu list64 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke.log 2>&1
echo "ncu smoke exit $?"
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:conv_slab|conv_tc" -c 10 -o gpurun_out/full_conv -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_clstm.csv python tools/profile_clstm.py > gpurun_out/ncu_clstm.log 2>&1
M2=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M2 --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py bf16 > gpurun_out/ncu_train.log 2>&1
echo "ncu train list exit $?"
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:wgrad_mma" -c 6 -o gpurun_out/full_wgrad -f python tools/profile_train.py bf16 > gpurun_out/ncu_full_wgrad.log 2>&1
echo "ncu full wgrad exit $?"
python tools/train_events.py 8 bf16 > gpurun_out/train_events8.txt 2>&1
python tools/step_events.py 8 5 > gpurun_out/events8.txt 2>&1
python tools/step_events.py 64 3 > gpurun_out/events64.txt 2>&1
ls -la gpurun_out | tail -n 14
