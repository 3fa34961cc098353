#!/bin/bash
# round 2 evidence, call A: launch lists (smoke, one mask-search iteration, one ConvLSTM iteration)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1
echo "smoke ncu rc=$?"
python tools/profile_step.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step8clips.csv \
    python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "step ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches_step8clips.csv --all > gpurun_out/r02_launches_step8clips.txt 2>&1
python tools/step_metrics.py gpurun_out/r02_launches_step8clips.csv gpurun_out/r02_step_metrics.json
head -30 gpurun_out/r02_launches_step8clips.txt
python - <<'PY'
import csv, collections
rows=list(csv.reader(open("gpurun_out/r02_smoke_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
names=[r[4] for r in rows[h+1:] if len(r)>8]
c=collections.Counter(n.split("(")[0][-60:] for n in names)
print(len(names),"launches in smoke list; first 12:"); 
for n in names[:12]: print("   ",n[:100])
for k,v in c.most_common(25): print("%5d %s"%(v,k))
PY
