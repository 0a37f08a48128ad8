#!/bin/bash
mkdir -p gpurun_out
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_cyc.so python tools/perf_probe.py C2 60 2368 > gpurun_out/r2_7_probe_cyc.log 2>&1; tail -3 gpurun_out/r2_7_probe_cyc.log
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_cyc.so python tools/perf_probe.py C5 80 4736 > gpurun_out/r2_7_probe_cyc_c5.log 2>&1; tail -3 gpurun_out/r2_7_probe_cyc_c5.log
