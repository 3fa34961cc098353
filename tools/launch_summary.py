"""Summarise an ncu launch list (gpu__time_duration.sum CSV) per kernel and per launch.
usage: python tools/launch_summary.py gpurun_out/launches.csv [--all]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    tot, agg, per = 0.0, {}, []
    mi = H.index("Metric Name") if "Metric Name" in H else None
    for r in rows[hdr + 1:]:
        if len(r) < len(H) or (mi is not None and r[mi] != "gpu__time_duration.sum"):
            continue  # multi-metric lists: time rows only
        name = r[4]
        short = name.split("::")[1].split("(")[0] if "::" in name else name[:40]
        t = float(r[-1].replace(",", ""))
        tot += t
        a = agg.setdefault(short[:44], [0, 0.0])
        a[0] += 1
        a[1] += t
        per.append((int(r[0]), short[:44], r[8], t))
    print("total %.1f us over %d launches" % (tot / 1e3, len(per)))
    print("%-46s %5s %10s %6s" % ("kernel", "n", "us", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-46s %5d %10.1f %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    if "--all" in sys.argv:
        for i, k, g, t in per:
            print(i, k, g, "%.1f" % (t / 1e3))
    else:
        print("top launches:")
        for i, k, g, t in sorted(per, key=lambda x: -x[3])[:24]:
            print("  #%d %-40s grid %-16s %8.1f us" % (i, k, g, t / 1e3))


if __name__ == "__main__":
    main()
