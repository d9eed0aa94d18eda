"""Join an ncu SASS source page (csv) with nvdisasm line info: per-source-line instruction / sample totals.
usage: sass_lines.py source.csv cubin kernel_mangled_name [top]"""
import csv, re, subprocess, sys, collections
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find the kernel section
start = next(i for i, l in enumerate(dis) if l.startswith("//---") and kname in l)
line_of = []   # per instruction in order: (line, inlined chain)
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---"):
        break
    m = re.search(r'//## File ".*?", line (\d+)(.*)', l)
    if m:
        cur = int(m.group(1))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        line_of.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; body = rows[2:]
assert len(body) == len(line_of), (len(body), len(line_of))
ex = collections.Counter(); sm = collections.Counter(); st = collections.Counter()
for r, ln in zip(body, line_of):
    e = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ex[ln] += e; sm[ln] += s; st[ln] += 1
te, ts = sum(ex.values()), sum(sm.values())
srcfile = [l.rstrip("\n") for l in open(sys.argv[5])] if len(sys.argv) > 5 else None
print(f"total executed {te}, samples {ts}, static {len(body)}")
for ln, e in ex.most_common(top):
    text = srcfile[ln - 1].strip()[:90] if srcfile and ln else ""
    print(f"line {ln}: static {st[ln]:5d} exec {100*e/te:5.2f}% samples {100*sm[ln]/ts:5.2f}%  {text}")

# ---- aggregate by source line ranges given as name:lo-hi (argv[6:])
if len(sys.argv) > 6:
    print("\nby region:")
    for spec in sys.argv[6:]:
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        e = sum(v for k, v in ex.items() if k and lo <= k <= hi); s = sum(v for k, v in sm.items() if k and lo <= k <= hi)
        n = sum(v for k, v in st.items() if k and lo <= k <= hi)
        print(f"{name:16s} static {n:6d} exec {100*e/te:6.2f}% samples {100*s/ts:6.2f}%")
