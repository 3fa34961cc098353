set -x
python bench.py --steps 100 --warmup 3 --no-cpu --no-gradcam --no-clstm --e2e-micro-batch 128 2>gpurun_out/x1_mb128.err | tail -1 > gpurun_out/x1_mb128.json
python - <<'P' > gpurun_out/x1_clstm.txt 2>&1
import torch, json, bench
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
for n in (8, 16, 32):
    r = bench.clstm_throughput(dev, 0, 1, "bf16", clips_n=n, steps=10)
    print(n, r["value"], r["ms_per_step"], r["roofline"]["frac"], flush=True)
P
