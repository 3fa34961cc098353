mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "conv or two_dest" 2>&1 | tail -2
for cfg in "1" "force"; do
  echo "=== IVF_TUNE=$cfg"
  IVF_TUNE=$cfg timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-gradcam --no-clstm 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],4),'conv_ms',round(d['roofline']['conv_ms_per_step'],4),'frac',round(d['roofline']['frac'],4))
"
done
timeout 800 python tools/write_plans.py > gpurun_out/write_plans.log 2>&1; echo rc=$?; grep -v "None, None, None\] -> None" gpurun_out/write_plans.log | cut -c1-260
