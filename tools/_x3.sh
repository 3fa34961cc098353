M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/x3_clstm.csv python tools/profile_clstm.py > gpurun_out/x3_ncu.log 2>&1
echo done $?
