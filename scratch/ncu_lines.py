"""Instructions executed per CUDA source line of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).
usage: python scratch/ncu_lines.py <rep> <kernel-substring> [top]"""
import csv, io, subprocess, sys, collections
rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
blocks = out.split('"File Path",')
agg = collections.Counter(); text = {}; total = 0; samples = collections.Counter()
for b in blocks[1:]:
    lines = b.splitlines()
    fpath = lines[0].strip('"')
    fn = lines[1]
    if sub not in fn: continue
    rows = list(csv.reader(io.StringIO("\n".join(lines[2:]))))
    if not rows: continue
    h = rows[0]
    iL, iS, iI, iN = h.index("Line No"), 1, h.index("Instructions Executed"), h.index("# Samples")
    for r in rows[1:]:
        if len(r) <= iI or not r[iL]: continue
        try: n = int(r[iI])
        except ValueError: continue
        key = (fpath.split("/")[-1], int(r[iL]))
        agg[key] += n; text[key] = r[iS].strip()[:110]
        try: samples[key] += int(r[iN])
        except ValueError: pass
    break_after = True
total = sum(agg.values())
print("total warp instructions attributed:", total)
for key, n in agg.most_common(top):
    print(f"{100.0*n/total:5.1f}%  {n:>10d}  smp {samples[key]:>5d}  {key[0]}:{key[1]:<5d} {text[key]}")
