#!/usr/bin/env python
"""Dynamic SASS profile of one kernel: python scratch/ncu_dyn.py rep [mode]
 mode 'hist': executed warp-instructions by opcode; 'dump': every executed SASS line with its count, samples; 'src': by CUDA source line"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]; mode = sys.argv[2] if len(sys.argv) > 2 else "hist"
args = ["ncu", "-i", rep, "--page", "source", "--csv"]
if mode == "src": args += ["--print-source", "cuda,sass"] if False else []
out = subprocess.run(args, capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
r = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = r[0]
isrc = h.index("Source"); iex = h.index("Instructions Executed"); isamp = h.index("# Samples")
tot = sum(int(row[iex]) for row in r[1:] if row[iex].isdigit())
if mode == "hist":
    c = collections.Counter()
    for row in r[1:]:
        if not row[iex].isdigit(): continue
        m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", row[isrc])
        c[m.group(1) if m else "?"] += int(row[iex])
    print("total warp instructions", tot)
    for k, v in c.most_common(45): print(f"{v:10d} {100*v/tot:5.1f}%  {k}")
else:
    for i, row in enumerate(r[1:]):
        if not row[iex].isdigit(): continue
        print(f"{i:5d} {int(row[iex]):9d} {row[isamp]:>5s}  {row[isrc]}")
