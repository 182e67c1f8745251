#!/usr/bin/env python
"""Grep-able SASS evidence: per kernel of libfsb200.so the counts of the Blackwell-native instructions
(UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops), plus BRA.U.ANY = per-instruction "waterfall" loops around uniform-datapath instructions issued from divergent
code (0 in the MMA / TMA-load issue loops since round 2).  No GPU needed:  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fasterseg_b200", "libfsb200.so")
PAT = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "BRA.U.ANY", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r"\b%s\b" % re.escape(p), line) or (p == "BRA.U.ANY" and "BRA.U.ANY" in line):
                counts[cur][p] += 1
    print("# cuobjdump -sass fasterseg_b200/libfsb200.so (sm_100a), instruction counts per kernel; HMMA = legacy mma.sync path (must be 0)")
    print("%-62s" % "kernel" + "".join("%10s" % p for p in PAT))
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c[p] for p in PAT):
            continue
        print("%-62s" % k[:62] + "".join("%10d" % c[p] for p in PAT))
        tot.update(c)
    print("%-62s" % "TOTAL" + "".join("%10d" % tot[p] for p in PAT))


if __name__ == "__main__":
    main()
