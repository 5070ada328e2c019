#!/usr/bin/env python3
"""Schedule quality of a SASS address range: for every FP64-pipe instruction, the distance (in FP64-pipe
instructions) back to the producer of its nearest FP64 operand.  A schedule with many distance-1..3 pairs
stalls on the DFMA latency unless other warps fill in.
usage: python tools/sass_depdist.py file.sass 0xSTART 0xEND"""
import collections
import re
import sys

f, a, b = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
last, idx, dist = {}, 0, []
for l in open(f):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if not m or not (a <= int(m.group(1), 16) <= b):
        continue
    t = m.group(2)
    m2 = re.search(r"\b(DFMA|DMUL|DADD)\s+R(\d+),\s*(.*)", t)
    if not m2:
        continue
    d = int(m2.group(2))
    srcs = [int(x) for x in re.findall(r"R(\d+)", m2.group(3))]
    ds = [idx - last[s] for s in srcs if s in last]
    if ds:
        dist.append(min(ds))
    last[d] = idx
    idx += 1
c = collections.Counter(min(x, 8) for x in dist)
print("fp64 %d; producer distance 1..7, >=8: %s" % (idx, [c[k] for k in range(1, 9)]))
