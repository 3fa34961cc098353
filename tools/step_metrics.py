"""Per-launch ncu metrics of one iteration (multi-metric launch list written by tools/gpu_round2.sh, NCU=list)
-> JSON summary: per launch and per kernel family (conv = the tcgen05 implicit GEMMs).  bench.py reads
families.conv.dram_bytes from the committed copy (profiles/rNN_step_metrics.json) for roofline.traffic.
usage: python tools/step_metrics.py gpurun_out/launches.csv out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
mi, vi = H.index("Metric Name"), H.index("Metric Value")
launches = {}
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    i = int(r[0])
    name = r[4]
    short = name.split("::")[1].split("(")[0] if "::" in name else name[:40]
    e = launches.setdefault(i, {"id": i, "kernel": short[:40], "grid": r[8]})
    e[r[mi]] = float(r[vi].replace(",", ""))


def family(k):
    if k.startswith("conv_"):
        return "conv"
    if k.startswith("maxpool"):
        return "pool"
    return "other"


fam = {}
for e in launches.values():
    f = fam.setdefault(family(e["kernel"]), {"launches": 0, "time_us": 0.0, "dram_bytes": 0.0})
    f["launches"] += 1
    f["time_us"] += e.get("gpu__time_duration.sum", 0.0) / 1e3
    f["dram_bytes"] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
out = {"what": "per-launch ncu metrics of one mask-search iteration for 8 clips (tools/profile_step.py); times are "
               "cold-cache and serialised",
       "families": fam, "launches": [launches[i] for i in sorted(launches)]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print({k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in fam.items()})
