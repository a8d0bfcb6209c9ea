"""Dynamic opcode mix and stall profile of one kernel from an `ncu --set full --import-source on` report:

    ncu -i X.ncu-rep --page source --csv -k regex:KERNEL > src.csv ; python tools/sass_dynamic.py src.csv [top]

Aggregates `Instructions Executed` (warp-level) and stall samples by SASS opcode."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr = rows[1]
i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ex, smp, stalls = defaultdict(int), defaultdict(int), defaultdict(int)
tot = 0
for r in rows[2:]:
    if len(r) <= i_ex:
        continue
    op = r[i_src].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
    o = o.split(".")[0].rstrip(";")
    n = int(r[i_ex] or 0)
    ex[o] += n
    smp[o] += int(r[i_smp] or 0)
    tot += n
    for i, h in stall_cols:
        if i < len(r) and r[i]:
            stalls[h] += int(r[i])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f"warp instructions executed: {tot}")
for o, n in sorted(ex.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{o:12s} {n:12d} {100.0 * n / tot:6.2f} %   samples {smp[o]}")
st = sum(stalls.values())
print("stall samples:", {h[6:]: f"{100.0 * v / st:.1f}%" for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]})
