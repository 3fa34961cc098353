python -m pytest tests/test_gpu_kernels.py tests/test_gpu_i3d.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python tools/step_events.py 8 5 > gpurun_out/r3_events8c.txt 2>&1; grep "k1x1x1" gpurun_out/r3_events8c.txt | grep dgrad; tail -10 gpurun_out/r3_events8c.txt
python bench.py --steps 100 --warmup 3 --no-cpu --no-gradcam --no-clstm 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'conv_ms',d['roofline']['conv_ms_per_step'])"
python bench.py --steps 100 --warmup 3 --no-cpu --no-gradcam --no-clstm --e2e-micro-batch 64 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('mb64 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
