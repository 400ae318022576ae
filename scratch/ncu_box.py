"""Run ON the GPU box after ncu captures: summarise every kernel of every .ncu-rep under gpurun_out/ into small text
files (metrics + top stalled SASS lines), and drop the reports themselves if they would blow gpurun's 64 MiB return limit.
usage: python scratch/ncu_box.py <tag> [keep_mb]"""
import csv, glob, io, os, re, subprocess, sys, json
tag = sys.argv[1]
keep_mb = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TU = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "s": 1, "second": 1}
def num(s):
    try: return float(s.replace(",", ""))
    except Exception: return None
def short(kn):
    kn = kn.replace("<unnamed>::", "")
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)(<[^>]*>)?", kn)
    s = m.group(1) + (m.group(2) or "") if m else kn[:40]
    return re.sub(r"[^A-Za-z0-9_]+", "_", s).strip("_")
traffic = {}
os.makedirs("gpurun_out/summ", exist_ok=True)
for rep in sorted(glob.glob("gpurun_out/*.ncu-rep")):
    base = os.path.basename(rep)[:-8]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    if len(r) < 3: continue
    h, u = r[0], r[1]
    seen = {}
    for v in r[2:]:
        kn = v[h.index("Kernel Name")]
        sn = short(kn)
        if sn in seen: continue
        seen[sn] = 1
        kid = v[h.index("ID")]
        name = f"{tag}_{base}__{sn}"
        lines = [f"# {name}: {kn}", f"# ncu --set full --clock-control none, one launch (ID {kid} of {base}.ncu-rep), 2048^2 fp32, inside bench.py --quick"]
        d = {}
        for w in WANT:
            if w in h:
                i = h.index(w); lines.append(f"{w:85s} {v[i]:>16s} {u[i]}"); d[w] = (num(v[i]), u[i])
        try:
            rd = d['dram__bytes_read.sum'][0] * UNIT.get(d['dram__bytes_read.sum'][1], 1)
            wr = d['dram__bytes_write.sum'][0] * UNIT.get(d['dram__bytes_write.sum'][1], 1)
            t = d['gpu__time_duration.sum'][0] * TU.get(d['gpu__time_duration.sum'][1], 1e-6)
            lines.append(f"{'dram traffic (read+write) bytes':85s} {rd+wr:16.0f}")
            lines.append(f"{'dram GB/s under ncu (cold, serialised)':85s} {(rd+wr)/t/1e9:16.1f}")
            traffic[name] = {"kernel": kn, "dram_bytes": rd + wr, "time_us": t * 1e6}
        except Exception as e:
            lines.append(f"# traffic unavailable: {e}")
        # top stalled SASS lines of this kernel
        try:
            so = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(int(kid)), "--launch-count", "1"],
                                capture_output=True, text=True).stdout
            sl = so.splitlines()
            start = next(i for i, l in enumerate(sl) if l.startswith('"Address"') or l.startswith('"#"'))
            rr = list(csv.reader(io.StringIO("\n".join(sl[start:]))))
            hh = rr[0]; isrc = hh.index("Source"); isamp = hh.index("# Samples")
            rows = []
            tot = 0
            for row in rr[1:]:
                try: s = int(row[isamp])
                except Exception: continue
                tot += s; rows.append((s, row))
            order = sorted(range(len(rows)), key=lambda k: -rows[k][0])[:30]
            lines.append(f"# top SASS lines by warp-state samples (total {tot}); index = position in the kernel's SASS")
            stallcols = [c for c in hh if c.startswith("stall_")]
            for k in order:
                s, row = rows[k]
                st = {c[6:]: row[hh.index(c)] for c in stallcols if row[hh.index(c)] not in ("0", "")}
                top = sorted(st.items(), key=lambda kv: -float(kv[1]))[:3]
                lines.append(f"{100*s/max(tot,1):5.1f}%  #{k:5d}  {row[isrc][:90]:90s} {dict(top)}")
        except Exception as e:
            lines.append(f"# source page unavailable: {e}")
        open(f"gpurun_out/summ/{name}.txt", "w").write("\n".join(lines) + "\n")
tj = f"gpurun_out/summ/{tag}_traffic.json"
try:
    prev = json.load(open(tj))
except Exception:
    prev = {}
prev.update(traffic)
json.dump(prev, open(tj, "w"), indent=1)
# return-size guard
reps = sorted(glob.glob("gpurun_out/*.ncu-rep"), key=os.path.getmtime, reverse=True)
tot = 0
for rep in reps:
    sz = os.path.getsize(rep) / 1e6
    if tot + sz > keep_mb:
        os.remove(rep); print(f"dropped {rep} ({sz:.1f} MB)")
    else:
        tot += sz; print(f"kept {rep} ({sz:.1f} MB)")
