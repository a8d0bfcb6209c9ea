#!/bin/bash
# Strong-scaling and query-partition runs of bench.py on one 8-GPU box: results -> gpurun_out/r02_scaling.jsonl
# usage (on the GPU box): bash tools/scaling_runs.sh
out=${1:-gpurun_out/r02_scaling.jsonl}
: > $out
run() {   # n, extra args...
  n=$1; shift
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline "$@" >> $out 2>> gpurun_out/r02_scaling.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" >> $out 2>> gpurun_out/r02_scaling.err
  fi
}
for n in 1 2 4 8; do run $n --scaling strong; done                                   # ns64, global batch 32 fields split over N
run 8                                                                                 # weak (the driver's SCALE line) for reference
for n in 1 8; do run $n --config sw192 --scaling strong --partition queries; done     # 4 fields < 8 ranks: shard the queries
for n in 1 8; do run $n --config ihc --scaling strong --partition queries --steps 3; done   # 1 field: "8-GPU sharded"
for n in 1 8; do run $n --config plane64 --scaling strong; done
wc -l $out
