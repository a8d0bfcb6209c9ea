"""Key metrics of an `ncu --set full` report exported with `ncu -i X.ncu-rep --page raw --csv > X_raw.csv`.

    python tools/ncu_summary.py X_raw.csv > profiles/X_ncu.txt
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"   {k:100s} {r[i]:>18s} {units[i]}")
