#!/bin/bash
# Final evidence run of a round on one B200: GPU tests, smoke, bench lines of every config, forward-only / recompute / ODE lines,
# launch list + ncu full capture of the pair kernels (ns64).  Results -> gpurun_out/<tag>_*
tag=${1:-r02f}
o=gpurun_out
python -m pytest tests -q -m gpu > $o/${tag}_gputests.log 2>&1; tail -2 $o/${tag}_gputests.log
python __graft_entry__.py smoke > $o/${tag}_smoke.log 2>&1; tail -3 $o/${tag}_smoke.log
python bench.py --steps 10 --warmup 3 > $o/${tag}_bench_ns64.json 2> $o/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_arm.json 2>> $o/${tag}_bench.err
: > $o/${tag}_bench_other_configs.jsonl
for c in plane64 sphere sw192 ihc; do python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline >> $o/${tag}_bench_other_configs.jsonl 2>> $o/${tag}_bench.err; done
python bench.py --forward-only --steps 5 --warmup 3 > $o/${tag}_bench_forward_only.jsonl 2>> $o/${tag}_bench.err
python bench.py --forward-only --out-bf16 --steps 5 --warmup 3 >> $o/${tag}_bench_forward_only.jsonl 2>> $o/${tag}_bench.err
python bench.py --recompute --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_recompute.json 2>> $o/${tag}_bench.err
python bench.py --ode --steps 10 --warmup 3 > $o/${tag}_bench_ode.json 2>> $o/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $o/${tag}_launches_ns64.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_launches.log 2>&1
python tools/launch_summary.py $o/${tag}_launches_ns64.csv > $o/${tag}_launches_ns64_summary.txt
ncu --set full --clock-control none --import-source on -k regex:pairs_ --launch-skip 8 -c 4 -f -o $o/${tag}_pairs python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $o/${tag}_ncu_full.log 2>&1
ncu -i $o/${tag}_pairs.ncu-rep --page raw --csv > $o/${tag}_pairs_raw.csv 2>/dev/null
wc -c $o/${tag}_*.json* | tail -12
