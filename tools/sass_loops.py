#!/usr/bin/env python3
"""Loops of one kernel in a SASS dump: for every backward branch, the body's instruction mix (FP64-pipe,
MUFU, shared/local loads and stores).  Shows at a glance whether a hot loop carries spill traffic.
usage: cuobjdump -sass -fun <mangled> file.o | python tools/sass_loops.py"""
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
print("instructions %d, loops %d" % (len(ins), len(loops)))
for s, e in sorted(loops):
    body = [t for _, t in ins[s:e + 1]]
    def cnt(p):
        return sum(1 for t in body if re.search(p, t))
    print("  [%5x..%5x] n=%4d fp64=%4d mufu=%3d lds=%3d sts=%2d ldl=%3d stl=%3d ldg=%2d bra=%2d" % (
        ins[s][0], ins[e][0], len(body), cnt(r"\b(DFMA|DMUL|DADD)\b"), cnt(r"\bMUFU"), cnt(r"\bLDS"), cnt(r"\bSTS"),
        cnt(r"\bLDL"), cnt(r"\bSTL"), cnt(r"\bLDG"), cnt(r"\bBRA\b")))
