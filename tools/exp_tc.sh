mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "conv or two_dest" 2>&1 | tail -2
timeout 800 python tools/tune_slab.py 5 2>&1 | tail -24
for cfg in "1" "force"; do
  echo "=== IVF_TUNE=$cfg"
  IVF_TUNE=$cfg timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-gradcam --no-clstm 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step',round(d['ms_per_step'],4),'conv_ms',round(d['roofline']['conv_ms_per_step'],4),'frac',round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1))
"
done
