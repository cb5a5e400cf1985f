#!/bin/bash
# Key metrics of one `ncu --set full` capture as text:  tools/ncu_summary.sh gpurun_out/x.ncu-rep > profiles/rNN_ncu_full_x.txt
rep="$1"
echo "# ncu --set full --clock-control none, one launch, from $rep"
ncu -i "$rep" --page raw --csv 2>/dev/null | python3 -c '
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "launch__cluster", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu", "smsp__pcsamp_warps_issue_stalled", "smsp__average_warp")
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want):
        print(f"{h:100s} {v} {u}")
'
