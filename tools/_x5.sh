IVF_CLSTM_FUSED=2 python -m pytest tests/test_gpu_clstm.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/x5_tests.log
rm -f gpurun_out/x5_bench.txt
for f in 0 2; do
  echo "== IVF_CLSTM_FUSED=$f" >> gpurun_out/x5_bench.txt
  IVF_CLSTM_FUSED=$f python - >> gpurun_out/x5_bench.txt 2>&1 <<'P'
import torch, bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
for n in (8, 16):
    r = bench.clstm_throughput(dev, 0, 1, "bf16", clips_n=n, steps=10)
    print("clstm", n, r["value"], r["ms_per_step"], r["launches_per_step"], flush=True)
P
done
