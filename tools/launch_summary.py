"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X`).

    python tools/launch_summary.py X.csv [steps]

`steps`: number of fwd+bwd passes in the capture (default: inferred from the count of the pair forward kernel)."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    rows.append((r["Kernel Name"], us))
agg = OrderedDict()
for k, us in rows:
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + us)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else max([n for k, (n, t) in agg.items() if "pairs_fwd" in k] or [1])
tot = sum(t for n, t in agg.values())
print(f"# {path}: {len(rows)} launches, {steps} fwd+bwd passes; times are cold-cache / serialised under ncu: compare SHARES")
print(f"{'kernel':56s} {'n':>5s} {'total_us':>11s} {'share':>6s} {'avg_us':>9s} {'us/step':>9s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:56]:56s} {n:5d} {t:11.1f} {t / tot:6.3f} {t / n:9.1f} {t / steps:9.1f}")
print(f"total us/step {tot / steps:.2f}")
