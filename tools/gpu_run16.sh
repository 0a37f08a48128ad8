#!/bin/bash
mkdir -p gpurun_out
C="python tools/c4_probe.py 18 100000 296"
for v in pf1 pf8 pf16; do
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_$v.so $C > gpurun_out/r2_16_c4_100k_$v.log 2>&1; grep "^C4" gpurun_out/r2_16_c4_100k_$v.log | cut -c1-200; grep sha1 gpurun_out/r2_16_c4_100k_$v.log
done
