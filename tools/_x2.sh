python -m pytest tests/test_gpu_kernels.py tests/test_gpu_clstm.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/x2_tests.log
for x in 0 1 2; do
  echo "== IVF_SLAB_XPOSE=$x" >> gpurun_out/x2_bench.txt
  IVF_SLAB_XPOSE=$x python bench.py --steps 200 --warmup 3 --no-cpu --no-gradcam --no-clstm --clips-per-gpu 8 --e2e-micro-batch 8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'conv_ms',d['roofline']['conv_ms_per_step'])" >> gpurun_out/x2_bench.txt
  IVF_SLAB_XPOSE=$x python - >> gpurun_out/x2_bench.txt 2>&1 <<'P'
import torch, bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
for n in (8, 32):
    r = bench.clstm_throughput(dev, 0, 1, "bf16", clips_n=n, steps=10)
    print("clstm", n, r["value"], r["ms_per_step"], flush=True)
P
done
