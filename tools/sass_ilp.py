#!/usr/bin/env python3
"""Rough ILP check on a SASS dump: fraction of FP64-pipe instructions that read the destination of
the FP64-pipe instruction issued immediately before them (a serialised chain stalls on latency).
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_ilp.py"""
import re
import sys

prev_dst = None
n = dep = 0
for line in sys.stdin:
    m = re.search(r"\b(DFMA|DMUL|DADD)\s+(R\d+),\s*(.*?);", line)
    if not m:
        if re.search(r"\b(BRA|BSYNC|EXIT|CALL)\b", line):
            prev_dst = None
        continue
    dst, srcs = m.group(2), m.group(3)
    n += 1
    if prev_dst and re.search(r"\b%s\b" % prev_dst, srcs.replace("-", "")):
        dep += 1
    prev_dst = dst
print("fp64 instrs %d, back-to-back dependent %d (%.1f%%)" % (n, dep, 100.0 * dep / max(n, 1)))
