o=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pairs_ --launch-skip 8 -c 4 -f -o $o/r02_ihc_pairs python bench.py --config ihc --steps 1 --warmup 3 --no-cpu-baseline > $o/r02_ihc_ncu_full.log 2>&1
ncu -i $o/r02_ihc_pairs.ncu-rep --page raw --csv > $o/r02_ihc_pairs_raw.csv 2>/dev/null
ncu -i $o/r02_ihc_pairs.ncu-rep --page source --csv -k regex:pairs_bwd_tc_q > $o/r02_ihc_q_source.csv 2>/dev/null
ls -la $o/r02_ihc_pairs.ncu-rep
