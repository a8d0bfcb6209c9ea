#!/bin/bash
# Round-2 evidence run on one B200 (results -> gpurun_out/, copied to profiles/ by hand):
#   bench line (ns64), forward-only line, recompute line, ncu launch list of one step, ncu --set full of the four pair kernels
tag=${1:-r02}
o=gpurun_out
python bench.py --steps 10 --warmup 3 > $o/${tag}_bench_ns64.json 2> $o/${tag}_bench_ns64.err
python bench.py --forward-only --steps 5 --warmup 3 > $o/${tag}_bench_fwdonly.json 2>> $o/${tag}_bench_ns64.err
python bench.py --forward-only --out-bf16 --steps 5 --warmup 3 >> $o/${tag}_bench_fwdonly.json 2>> $o/${tag}_bench_ns64.err
python bench.py --recompute --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_recompute.json 2>> $o/${tag}_bench_ns64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/${tag}_launches_ns64.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launches.log 2>&1
python tools/launch_summary.py $o/${tag}_launches_ns64.csv > $o/${tag}_launches_ns64_summary.txt
# one launch of each pair kernel from a warm step (skip the first two steps' launches of each)
ncu --set full --clock-control none --import-source on -k regex:pairs_ --launch-skip 8 -c 4 -f -o $o/${tag}_pairs python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
ncu -i $o/${tag}_pairs.ncu-rep --page raw --csv > $o/${tag}_pairs_raw.csv 2>/dev/null
ls -la $o/${tag}_pairs.ncu-rep
