python -m pytest tests -m gpu -x -q > gpurun_out/r3_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3_tests.log
for cfg in "IVF_PDL=1" "IVF_PDL=0"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 200 --warmup 3 --no-cpu --no-gradcam --no-clstm --clips-per-gpu 64 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
done
