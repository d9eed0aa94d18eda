"""Summarise an .ncu-rep: headline raw metrics of the first kernel + per-source-line hot spots.
usage: ncu_summary.py report.ncu-rep [top]"""
import csv, collections, subprocess, sys, json
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg']
out = {}
for i, h in enumerate(hdr):
    if h in want or ('issue_stalled' in h and 'per_issue_active' in h):
        try:
            v = float(vals[i].replace(',', ''))
        except ValueError:
            continue
        if 'issue_stalled' in h and v < 0.05:
            continue
        out[h] = [v, units[i]]
print(json.dumps(out, indent=1))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; hd = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0]); text = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hd = {h: i for i, h in enumerate(r)}; continue
    if hd is None or len(r) < 10: continue
    try: ln = int(r[0])
    except ValueError: continue
    if r[2] == '-':
        a = agg[(cur, ln)]
        a[0] += int(r[hd["Instructions Executed"]] or 0); a[1] += int(r[hd["Thread Instructions Executed"]] or 0)
        a[2] += int(r[hd["# Samples"]] or 0); a[3] += int(r[hd["stall_no_inst"]] or 0); a[4] += int(r[hd["stall_long_sb"]] or 0)
        text[(cur, ln)] = r[1]
te = sum(a[0] for a in agg.values()) or 1; ts = sum(a[2] for a in agg.values()) or 1
print("total warp-instructions", te, "samples", ts, "no_inst samples %.1f%%" % (100 * sum(a[3] for a in agg.values()) / ts),
      "long_sb samples %.1f%%" % (100 * sum(a[4] for a in agg.values()) / ts))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f}:{l} exec {100*a[0]/te:5.2f}% thr/inst {a[1]/max(a[0],1):5.1f} samples {100*a[2]/ts:5.2f}% (no_inst {100*a[3]/max(a[2],1):3.0f}% long_sb {100*a[4]/max(a[2],1):3.0f}%)  {text[(f,l)][:90]}")
