#!/usr/bin/env python
"""usage: sass_hist.py sass.txt lo hi  -- opcode histogram of the instructions whose address is in [lo, hi] (hex)"""
import re, sys, collections
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
h = collections.Counter(); n = 0
for ln in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
    if not m: continue
    a = int(m.group(1), 16)
    if lo <= a <= hi:
        h[m.group(2)] += 1; n += 1
print(n, "instructions")
for k, v in h.most_common(40): print(f"{v:5d} {k}")
