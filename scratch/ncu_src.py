"""Top stalled SASS/source lines of an .ncu-rep: python scratch/ncu_src.py rep [n] [view]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
args = ["ncu", "-i", rep, "--page", "source", "--csv"]
if view == "cuda": args += ["--print-source", "cuda"]
out = subprocess.run(args, capture_output=True, text=True).stdout
lines = out.splitlines()
# skip the 'Kernel Name' preamble
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"Line') or l.startswith('"#"'))
r = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
h = r[0]
isrc = h.index("Source"); isamp = h.index("# Samples")
names = ["stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_membar", "stall_not_selected"]
idx = [h.index(x) for x in names]
rows = []
tot = 0
for row in r[1:]:
    try: s = int(row[isamp])
    except: continue
    tot += s
    rows.append((s, row))
rows.sort(key=lambda t: -t[0])
print("total samples", tot)
for s, row in rows[:n]:
    st = {nm: row[i] for nm, i in zip(names, idx) if row[i] not in ("0", "")}
    print(f"{100*s/tot:5.1f}%  {row[isrc][:110]:110s} {st}")
