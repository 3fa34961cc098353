#!/usr/bin/env python
"""Opcode histogram of the convolution kernels in libivf.so (cuobjdump -sass): the SASS mnemonics that prove
tcgen05 / TMEM / TMA (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG =
cp.async.bulk.tensor).  Needs no GPU.   python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "interpreting_video_features_b200", "libivf.so")
KEEP = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA",
        "UTCATOMSWS", "ELECT", "FFMA", "LDG", "STG", "LDS", "STS", "ATOMS", "RED", "BAR", "MUFU")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    want = sys.argv[1:] or ["conv_slab_kernel", "conv_tc_kernel"]
    print("# cuobjdump -sass interpreting_video_features_b200/libivf.so  (sm_100a)")
    for name, ops in kernels.items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if not any(w in dem for w in want):
            continue
        fam = collections.Counter()
        for op, n in ops.items():
            base = op.split(".")[0]
            if base in KEEP:
                fam[op if base.startswith(("UTC", "UTMA", "LDTM", "STTM")) else base] += n
        print("\n%s\n  %d instructions" % (dem[:150], sum(ops.values())))
        for op, n in sorted(fam.items(), key=lambda kv: (-kv[1], kv[0])):
            print("  %-28s %6d" % (op, n))
    tot = collections.Counter()
    for ops in kernels.values():
        for op, n in ops.items():
            tot[op.split(".")[0]] += n
    print("\n# whole library: " + ", ".join("%s %d" % (k, tot[k]) for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA")))


if __name__ == "__main__":
    main()
