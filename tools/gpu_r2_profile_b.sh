#!/bin/bash
# round 2 evidence, call B: launch list at 32 clips per step (the e2e micro-batch) and ncu --set full of the top kernels
mkdir -p gpurun_out
IVF_PROFILE_CLIPS=32 python tools/profile_step.py > gpurun_out/prof32_plain.log 2>&1 &&
IVF_PROFILE_CLIPS=32 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step32clips.csv \
    python tools/profile_step.py > gpurun_out/ncu32.log 2>&1
echo "ncu32 rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_step32clips.csv > gpurun_out/r02_launches_step32clips.txt 2>&1
head -22 gpurun_out/r02_launches_step32clips.txt
python tools/profile_step.py > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --import-source on --clock-control none --profile-from-start off -k "regex:conv_slab_kernel|conv_tc_kernel" -c 12 \
    -o gpurun_out/r02_full_conv -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r02_full_conv.ncu-rep
