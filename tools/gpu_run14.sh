#!/bin/bash
mkdir -p gpurun_out
P="python tools/perf_probe.py C2 120 14208"
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_lp0.so $P > gpurun_out/r2_14_probe_lp0.log 2>&1
AW_SMEM_PAD=11264 ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_lp0.so $P > gpurun_out/r2_14_probe_lp0_pad11k.log 2>&1
AW_SMEM_PAD=20480 ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_lp0.so $P > gpurun_out/r2_14_probe_lp0_pad20k.log 2>&1
$P > gpurun_out/r2_14_probe_default.log 2>&1
AW_SMEM_PAD=8192 $P > gpurun_out/r2_14_probe_default_pad8k.log 2>&1
grep -H "pairs/s" gpurun_out/r2_14_probe_*.log | grep "it=1"
