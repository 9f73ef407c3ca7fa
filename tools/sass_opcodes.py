#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of nice_slam_b200/libnsb.so (cuobjdump -sass; runs without a GPU).  The Blackwell-native evidence:
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier,
REDG / RED = red.global.add, ELECT = elect.sync.   python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nice_slam_b200", "libnsb.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "ELECT", "R2UR", "REDG", "RED", "ATOMG", "LDG", "STG", "LDS", "STS", "FFMA", "DFMA", "MUFU", "SHFL", "BAR"]
print("SASS opcode counts per kernel (%s)" % os.path.relpath(lib, ROOT))
print("%-44s %8s  " % ("kernel", "instrs") + " ".join("%7s" % k for k in KEY))
for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    print("%-44s %8d  " % (k[:44], sum(h.values())) + " ".join("%7d" % h.get(x, 0) for x in KEY))
