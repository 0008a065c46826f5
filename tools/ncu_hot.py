"""Summarise an `ncu --page source --csv --print-source sass` export: top SASS instructions by stall samples, per reason.
Usage: python tools/ncu_hot.py src_sass.csv [topN] [kernel-index]"""
import csv, sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows, hdr, k = [], None, -1
with open(path) as f:
    for r in csv.reader(f):
        if r and r[0] == "Kernel Name":
            k += 1
            need_hdr = True
            continue
        if k != kidx:
            continue
        if need_hdr:
            hdr = r
            need_hdr = False
            continue
        rows.append(dict(zip(hdr, r)))
tot = sum(int(r["# Samples"]) for r in rows)
print("kernel", kidx, "instructions", len(rows), "samples", tot, "warp-insts", sum(int(r["Instructions Executed"]) for r in rows))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[h]) for r in rows) for h in reasons}
print({h: v for h, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
op = defaultdict(lambda: [0, 0])
for r in rows:
    m = r["Source"].split()
    name = m[1] if m[0].startswith("@") else m[0]
    op[name.split(".")[0]][0] += int(r["Instructions Executed"])
    op[name.split(".")[0]][1] += int(r["# Samples"])
print("opcode mix (warp insts, samples):")
for n, (c, s) in sorted(op.items(), key=lambda x: -x[1][0])[:25]:
    print(f"  {n:10s} {c:10d} {s:8d}")
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"]))[:top]
for i in sorted(idx):
    r = rows[i]
    main = max(reasons, key=lambda h: int(r[h]))
    print(f"{i:6d} {int(r['# Samples']):6d} {100.0*int(r['# Samples'])/tot:5.2f}% {int(r['Instructions Executed']):8d} {main:18s} {r['Source'][:90]}")
