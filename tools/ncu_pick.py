#!/usr/bin/env python
"""Print selected metrics from an `ncu --page raw --csv` dump.  Usage: ncu_pick.py raw.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
keys = sys.argv[2:] or ["gpu__time_duration.sum", "pipe_fp64", "issue_active.avg.pct", "warps_active.avg.pct",
                        "registers_per_thread", "grid_size", "block_size", "warp_issue_stalled", "dram__bytes_read.sum ",
                        "dram__bytes_write.sum ", "inst_executed.sum ", "thread_inst_executed_per_inst",
                        "pipe_lsu", "pipe_xu", "pipe_alu", "pipe_fma", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "subpipe_dmma", "pipe_tensor_cycles_active", "wavefronts_mem_shared.sum", "bank_conflicts_pipe_lsu_mem_shared.sum", "occupancy", "sm__throughput", "sm__cycles_elapsed.max", "warp_cycles_per_issued"]
for row in rows[2:]:
    print("==", row[hdr.index("Kernel Name")][:50] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, row):
        if any(k in h for k in keys):
            print(f"  {h} [{u}] = {v}")
