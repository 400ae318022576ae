"""usage: sass_hot.py <cuobjdump -sass output> <substring of the mangled kernel name>
Static SASS count of a kernel: whole function and the part before its last unpredicated EXIT (the cold out-of-line handlers of
divergent shuffles and IEEE slow paths sit after it), with the opcode histogram of the hot part."""
import re, sys, collections
f, pat = sys.argv[1], sys.argv[2]
lines = open(f).read().split('\n')
starts = [i for i,l in enumerate(lines) if 'Function :' in l]
for si, s in enumerate(starts):
    if pat not in lines[s]: continue
    e = starts[si+1] if si+1 < len(starts) else len(lines)
    body=[]
    for ln in lines[s:e]:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\w+\s+)?)([A-Z0-9_]+)(.*?);", ln)
        if m: body.append((m.group(3), m.group(2).strip(), ln))
    # hot part = up to the last unpredicated EXIT
    last = max(i for i,(op,pr,_) in enumerate(body) if op=='EXIT' and not pr)
    h = collections.Counter(op for op,_,_ in body[:last+1])
    print(len(body), "total;", last+1, "before the last EXIT")
    print(" ".join(f"{k}:{v}" for k, v in h.most_common(30)))
