#!/usr/bin/env python
"""Opcode histogram of a SASS address range: sasshist.py file.sass 0xLO 0xHI [...more ranges]."""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
pat = re.compile(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?')
ranges = [(int(sys.argv[i], 16), int(sys.argv[i + 1], 16)) for i in range(2, len(sys.argv) - 1, 2)]
for lo, hi in ranges:
    h = collections.Counter()
    for ln in lines:
        m = pat.match(ln)
        if not m:
            continue
        a = int(m.group(1), 16)
        if lo <= a < hi:
            op = m.group(2)
            if op in ("IMAD",) and m.group(3) and ".MOV" in m.group(3):
                op = "IMAD.MOV"
            h[op] += 1
    tot = sum(h.values())
    print("%#x-%#x: %d instr: %s" % (lo, hi, tot, " ".join("%s=%d" % kv for kv in h.most_common())))
