"""Share of executed warp instructions and stall samples per source region of an .ncu-rep source page.
usage: ncu_regions.py report.ncu-rep file:lo-hi=name ...   (lines outside every region -> 'other')"""
import csv, collections, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    loc, name = a.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    regions.append((f, int(lo), int(hi), name))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; hd = None
agg = collections.defaultdict(lambda: collections.Counter())
stall_cols = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hd = {}
        for i, h in enumerate(r):
            hd.setdefault(h, i)
        stall_cols = [h for h in hd if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hd is None or len(r) < 10: continue
    try: ln = int(r[0])
    except ValueError: continue
    if r[2] != '-': continue
    name = "other"
    for f, lo, hi, n in regions:
        if cur == f and lo <= ln <= hi: name = n; break
    a = agg[name]
    a["inst"] += int(r[hd["Instructions Executed"]] or 0)
    a["samples"] += int(r[hd["# Samples"]] or 0)
    for c in stall_cols:
        a[c] += int(r[hd[c]] or 0)
ti = sum(a["inst"] for a in agg.values()) or 1; ts = sum(a["samples"] for a in agg.values()) or 1
for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    top = sorted(((c, a[c]) for c in stall_cols), key=lambda x: -x[1])[:5]
    print(f"{n:28s} inst {100*a['inst']/ti:5.1f}%  samples {100*a['samples']/ts:5.1f}%  " +
          " ".join(f"{c[6:]}={100*v/max(a['samples'],1):.0f}%" for c, v in top))
