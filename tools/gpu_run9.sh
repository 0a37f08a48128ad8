#!/bin/bash
mkdir -p gpurun_out
P="python tools/perf_probe.py C2 60 2368"
$P > gpurun_out/r2_9_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:aw_align_kernel -s 1 -c 1 -o gpurun_out/prof_r2_9_align $P > gpurun_out/r2_9_ncu.log 2>&1
tail -2 gpurun_out/r2_9_ncu.log
