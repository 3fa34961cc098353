set -x
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_i3d.py -m gpu -x -q > gpurun_out/r3_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3_tests.log
tail -3 gpurun_out/r3_tests.log
IVF_TC_TRACE=1 python tools/trace_tc.py > gpurun_out/r3_trace_tc.log 2>&1; echo "trace rc=$?"
grep -v "tile [23]:" gpurun_out/r3_trace_tc.log | cut -c1-250
python bench.py --steps 100 --warmup 3 --no-cpu --no-gradcam --no-clstm > gpurun_out/r3_bench.log 2>&1; echo "bench rc=$?"
tail -c 600 gpurun_out/r3_bench.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r3_launches.csv python tools/profile_step.py > gpurun_out/r3_ncu.log 2>&1; echo "ncu rc=$?"
