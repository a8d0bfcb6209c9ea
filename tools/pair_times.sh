#!/bin/bash
# per-kernel durations of the four tcgen05 pair kernels (ncu, cold-cache serialised: compare between builds, not with bench.py)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pairs_ -c 16 --csv --log-file gpurun_out/pair_times.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/pair_times.csv') if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
acc = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("ns", "nsecond"): v /= 1e6
    elif r[ui] in ("us", "usecond"): v /= 1e3
    acc[r[ki][:40]].append(v)
for k, v in acc.items():
    v = sorted(v); print(f"{k:42s} n={len(v)} median={v[len(v)//2]:.3f} ms  min={v[0]:.3f}")
PY
