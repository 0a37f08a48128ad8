#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2_6_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_6_tests.log
tail -25 gpurun_out/r2_6_tests.log
P="python tools/perf_probe.py C2 60 2368"
$P > gpurun_out/r2_6_probe_default.log 2>&1
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_cmp.so $P threads_per_cta=256 > gpurun_out/r2_6_probe_cmp_nt256.log 2>&1
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_cmp.so $P ctas_per_sm=2 > gpurun_out/r2_6_probe_cmp_cta2.log 2>&1
grep -H "pairs/s" gpurun_out/r2_6_probe_*.log | grep "it=1"
timeout 600 python bench.py --scaling strong --gpus 1 --steps 1 --warmup 1 --pairs 37888 > gpurun_out/r2_6_strong_C2_n1.json 2> gpurun_out/r2_6_strong_C2_n1.err; cut -c1-200 gpurun_out/r2_6_strong_C2_n1.json; tail -3 gpurun_out/r2_6_strong_C2_n1.err
