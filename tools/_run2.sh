set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r3_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r3_tests.log
tail -3 gpurun_out/r3_tests.log
IVF_TC_TRACE=1 python tools/trace_tc.py > gpurun_out/r3_trace_tc.log 2>&1; echo "trace rc=$?"
cat gpurun_out/r3_trace_tc.log
