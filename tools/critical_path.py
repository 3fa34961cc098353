"""Critical path of one iteration from the serial launch list (ncu) and the program structure
(gpurun_out/program.json written by tools/profile_step.py): usage critical_path.py launches.csv program.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
lau = []
H = rows[hdr]
MI = H.index("Metric Name") if "Metric Name" in H else None
for r in rows[hdr + 1:]:
    if len(r) >= 15 and (MI is None or r[MI] == "gpu__time_duration.sum"):  # multi-metric lists: time rows only
        name = r[4]
        lau.append((name.split("::")[1].split("(")[0] if "::" in name else name[:30], float(r[-1]) / 1e3))
prog = json.load(open(sys.argv[2]))
i = 1  # launch 0 = perturb forward
lanes = [lau[0][1], 0.0, 0.0, 0.0]
onpath = {}
serial = lau[0][1]
phase_t = {}
for item in prog:
    ph, kind = item[0], item[1]
    if kind == "op":
        lane, n = item[2], item[3]
        dur = sum(lau[i + k][1] for k in range(n))
        key = lau[i][0][:18]
        lanes[lane] += dur
        serial += dur
        onpath.setdefault((ph, lane, key), 0.0)
        onpath[(ph, lane, key)] += dur
        i += n
    elif kind == "fork":
        which = item[2] if len(item) > 2 else [1, 2, 3]
        for l in which:
            lanes[l] = max(lanes[l], lanes[0])
    elif kind == "sync":
        lanes[item[3]] = max(lanes[item[3]], lanes[item[2]])
    elif kind == "join":
        lanes[0] = max(lanes)
        lanes = [lanes[0]] * 4
    if ph == "fwd":
        phase_t["fwd_end"] = lanes[0]
tail = sum(l[1] for l in lau[i:])
print("launches used %d of %d; tail (perturb', adam) %.1f us" % (i, len(lau), tail))
print("serial sum %.1f us ; critical path %.1f us (forward %.1f) + tail %.1f = %.1f us" % (
    serial + tail, lanes[0], phase_t.get("fwd_end", 0), tail, lanes[0] + tail))
agg = {}
for (ph, lane, key), v in onpath.items():
    agg.setdefault((ph, "main" if lane == 0 else "side", key), 0.0)
    agg[(ph, "main" if lane == 0 else "side", key)] += v
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("  %-4s %-5s %-20s %8.1f us" % (k[0], k[1], k[2], v))
