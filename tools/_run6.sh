for cfg in "IVF_STREAMS=1" "IVF_STREAMS=0" "IVF_STREAMS=0 IVF_PDL=0" "IVF_STREAMS=1 IVF_PDL=0"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 100 --warmup 3 --no-cpu --no-gradcam --no-clstm --clips-per-gpu 32 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'conv_ms',d['roofline']['conv_ms_per_step'])"
done
