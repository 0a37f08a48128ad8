#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_run14.sh
C="python tools/c4_probe.py 18 100000 296"
$C > gpurun_out/r2_15_c4_100k_pf4.log 2>&1; grep "^C4" gpurun_out/r2_15_c4_100k_pf4.log | cut -c1-200; grep sha1 gpurun_out/r2_15_c4_100k_pf4.log
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_pf1.so $C > gpurun_out/r2_15_c4_100k_pf1.log 2>&1; grep "^C4" gpurun_out/r2_15_c4_100k_pf1.log | cut -c1-200; grep sha1 gpurun_out/r2_15_c4_100k_pf1.log
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_pf8.so $C > gpurun_out/r2_15_c4_100k_pf8.log 2>&1; grep "^C4" gpurun_out/r2_15_c4_100k_pf8.log | cut -c1-200; grep sha1 gpurun_out/r2_15_c4_100k_pf8.log
