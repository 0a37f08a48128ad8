#!/bin/bash
# round-2 GPU call 2: occupancy variants (register budgets) + ncu of the new default kernel
mkdir -p gpurun_out
P="python tools/perf_probe.py C2 60 2368"
for v in r96 r80 r64; do
  ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_$v.so $P > gpurun_out/r2_2_probe_$v.log 2>&1
done
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_r64.so $P ctas_per_sm=6 > gpurun_out/r2_2_probe_r64_cta6.log 2>&1
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_r80.so $P ctas_per_sm=5 > gpurun_out/r2_2_probe_r80_cta5.log 2>&1
grep -H "pairs/s" gpurun_out/r2_2_probe_*.log | grep "it=1"
$P > gpurun_out/r2_2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:aw_align_kernel -s 1 -c 1 -o gpurun_out/prof_r2_2_align $P > gpurun_out/r2_2_ncu.log 2>&1
tail -2 gpurun_out/r2_2_ncu.log
