mkdir -p gpurun_out
for cfg in "" "IVF_POOL_S1COL=0" "IVF_FUSE_B0=0" "IVF_TC_TMA_STORE=0" "IVF_TUNE=0"; do
  echo "=== $cfg"
  env $cfg timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-clstm 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('i3d ms',round(d['ms_per_step'],4),'gradcam',round(d['gradcam']['value'],1), round(d['gradcam']['e2e'],1))
"
done
