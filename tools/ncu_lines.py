"""Join an `ncu --page source --csv --print-source sass` export with `nvdisasm --print-line-info` of the same cubin:
executed warp instructions and stall samples per source line (inlined callee lines attributed to the innermost line).
Usage: python tools/ncu_lines.py src_sass.csv all.sass '<mangled kernel name>' [topN]"""
import csv, re, sys
from collections import defaultdict

sass_csv, disasm, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows, hdr, k, need = [], None, -1, False
with open(sass_csv) as f:
    for r in csv.reader(f):
        if r and r[0] == "Kernel Name":
            k += 1; need = True; continue
        if k != 0: continue
        if need: hdr = r; need = False; continue
        rows.append(dict(zip(hdr, r)))
lines = []
cur = None
infunc = False
with open(disasm) as f:
    for ln in f:
        if ln.startswith(".text."):
            infunc = ln.strip() == f".text.{kname}:"
            continue
        if not infunc: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
print("sass rows", len(rows), "disasm insts", len(lines))
n = min(len(rows), len(lines))
agg = defaultdict(lambda: [0, 0, 0])
for i in range(n):
    a = agg[lines[i]]
    a[0] += int(rows[i]["Instructions Executed"]); a[1] += int(rows[i]["# Samples"]); a[2] += 1
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"{'file:line':28s} {'warp-insts':>10s} {'%':>6s} {'samples':>8s} {'%':>6s} {'sass':>5s}")
for key, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{str(key[0])+':'+str(key[1]):28s} {a[0]:10d} {100*a[0]/ti:6.2f} {a[1]:8d} {100*a[1]/ts:6.2f} {a[2]:5d}")
