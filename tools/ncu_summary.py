"""dev aid: key metrics of the first kernel in an ncu report.  usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
for h, u, v in zip(hdr, units, vals):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h and float(v or 0) > 0.2):
        print(f"{h:80s} {u:12s} {v}")
