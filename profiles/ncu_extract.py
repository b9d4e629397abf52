"""Extract the tracked summary of an `ncu --set full` report:  python profiles/ncu_extract.py REPORT.ncu-rep > summary.txt
(reads `ncu -i REPORT --page raw --csv`; one `metric unit value` line per metric kept)."""
import csv
import io
import subprocess
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__icc_request_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
print("kernel", vals[hdr.index("Kernel Name")])
for h, u, v in sorted(zip(hdr, units, vals)):
    if h in KEEP:
        print(h, u, v)
    elif h.startswith(STALL) and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio"):
        try:
            if float(v.replace(",", "")) >= 0.05:
                print("stall_" + h[len(STALL):].replace("_per_issue_active.ratio", "").replace(".ratio", ""), u, v)
        except ValueError:
            pass
