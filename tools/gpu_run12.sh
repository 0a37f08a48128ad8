#!/bin/bash
mkdir -p gpurun_out
P="python tools/perf_probe.py C2 120 14208"
$P > gpurun_out/r2_12_probe_default.log 2>&1
for v in lp0 lp0r96 lp0r80cmp; do
  ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_$v.so $P > gpurun_out/r2_12_probe_$v.log 2>&1
done
grep -H "pairs/s" gpurun_out/r2_12_probe_*.log | grep "it=1"
