#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 50 --warmup 3 --clips-per-gpu 64 --no-clstm > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
echo "rc=$?" >> gpurun_out/bench_2gpu.log
grep -c "NCCL INFO" gpurun_out/bench_2gpu.err; grep -i "allgather\|AllGather" gpurun_out/bench_2gpu.err | head -3
tail -3 gpurun_out/bench_2gpu.log | cut -c1-4000; tail -5 gpurun_out/bench_2gpu.err | cut -c1-300
