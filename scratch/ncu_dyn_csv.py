"""usage: ncu_dyn_csv.py <csv of `ncu -i rep --page source --csv`> <kernel index>
Executed warp-instructions and warp-state samples of one captured kernel by opcode."""
import csv, sys, re, collections, io
f=sys.argv[1]; which=int(sys.argv[2])
txt=open(f).read().split('\n')
ks=[i for i,l in enumerate(txt) if l.startswith('"Kernel Name"')]
s=ks[which]; e=ks[which+1] if which+1<len(ks) else len(txt)
r=list(csv.reader(io.StringIO("\n".join(txt[s+1:e]))))
h=r[0]; isrc=h.index("Source"); iex=h.index("Instructions Executed"); ismp=h.index("# Samples")
ops=collections.Counter(); smp=collections.Counter(); tot=0; tots=0
rows=[]
for row in r[1:]:
    if len(row)<=iex: continue
    try: ex=int(row[iex]); sm=int(row[ismp])
    except: continue
    m=re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", row[isrc])
    op=m.group(1) if m else '?'
    ops[op]+=ex; smp[op]+=sm; tot+=ex; tots+=sm
    rows.append((ex,sm,row[isrc]))
print(txt[s][:90]); print("warp instr executed", tot, "samples", tots)
for k,v in ops.most_common(24): print(f"{k:14s} {100*v/tot:5.1f}% instr  {100*smp[k]/tots:5.1f}% samples")
