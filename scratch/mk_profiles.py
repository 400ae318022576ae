"""Summarise ncu reports into profiles/ (tracked).  usage: mk_profiles.py <round-tag> <launch-csv> <rep>..."""
import csv, io, os, subprocess, sys, re, collections, json
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
def num(s):
    try: return float(s.replace(",", ""))
    except Exception: return None
def tobytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(u, 1)
out = {}
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    h, u, v = r[0], r[1], r[2]
    name = os.path.basename(rep)[:-8]
    kn = v[h.index("Kernel Name")]
    lines = [f"# {name}: {kn}", f"# source report: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on, one launch, 2048^2 fp32)"]
    d = {}
    for w in WANT:
        if w in h:
            i = h.index(w); lines.append(f"{w:85s} {v[i]:>16s} {u[i]}"); d[w] = (num(v[i]), u[i])
    rd = tobytes(*d['dram__bytes_read.sum']); wr = tobytes(*d['dram__bytes_write.sum'])
    t = d['gpu__time_duration.sum'][0] * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}.get(d['gpu__time_duration.sum'][1], 1e-6)
    lines.append(f"{'dram traffic (read+write) bytes':85s} {rd+wr:16.0f}")
    lines.append(f"{'dram GB/s under ncu (cold, serialised)':85s} {(rd+wr)/t/1e9:16.1f}")
    out[name] = {"kernel": kn, "dram_bytes": rd + wr, "time_us": t * 1e6}
    open(f"profiles/{name}.txt", "w").write("\n".join(lines) + "\n")
json.dump(out, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
# launch list
rows = [l for l in open(launches) if not l.startswith("==")]
agg = collections.OrderedDict(); tot = 0.0
for row in csv.DictReader(rows):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    n = re.sub(r"<unnamed>::", "", re.sub(r"\(.*", "", row["Kernel Name"]))
    x = float(row["Metric Value"].replace(",", "")); un = row["Metric Unit"]
    x = x / 1e3 if un in ("ns", "nsecond") else x * 1e3 if un in ("ms", "msecond") else x
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += x; tot += x
with open(f"profiles/{tag}_launches.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 1 --quick (2 passes over the six methods, 2048^2 fp32)\n")
    f.write(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:60]:60s} {n:8d} {t:12.1f} {t/n:10.2f} {100*t/tot:6.1f}%\n")
    f.write(f"total_us {tot:.1f}\n")
print(open(f"profiles/{tag}_launches.txt").read())
