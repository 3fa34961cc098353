for cfg in "IVF_TC_SPREAD=100" "IVF_TC_SPREAD=67" "IVF_TC_SPREAD=34" "IVF_TC_SPREAD=0"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 100 --warmup 3 --no-cpu --no-gradcam --no-clstm --clips-per-gpu 32 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'conv_ms',d['roofline']['conv_ms_per_step'])"
done
