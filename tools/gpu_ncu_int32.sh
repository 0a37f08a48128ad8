#!/bin/bash
# ncu counters (a few metrics, 1-3 replay passes) of the int32 regimes: 296 single-CTA pairs of 100 kb, and 8 pairs of 200 kb on clusters
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
A="python tools/c4_probe.py 18 100000 296"
B="python tools/c4_probe.py 4 200000 8"
timeout 100 $A > gpurun_out/int32_100k_plain.log 2>&1 && \
timeout 200 ncu --metrics $M --clock-control none -k regex:aw_align_kernel -c 1 --csv --log-file gpurun_out/int32_100k_ncu.csv $A > gpurun_out/int32_100k_ncu.log 2>&1
tail -3 gpurun_out/int32_100k_plain.log | cut -c1-250
timeout 100 $B > gpurun_out/int32_cluster_200k_plain.log 2>&1 && \
timeout 200 ncu --metrics $M --clock-control none -k regex:aw_align_kernel -c 1 --csv --log-file gpurun_out/int32_cluster_200k_ncu.csv $B > gpurun_out/int32_cluster_200k_ncu.log 2>&1
tail -3 gpurun_out/int32_cluster_200k_plain.log | cut -c1-250
