#!/bin/bash
# round-2 GPU call 1: parity suite, kernel variants on C2, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_1_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_1_tests.log
tail -3 gpurun_out/r2_1_tests.log
P="python tools/perf_probe.py C2 60 2368"
$P > gpurun_out/r2_1_probe_default.log 2>&1
$P threads_per_cta=256 > gpurun_out/r2_1_probe_nt256.log 2>&1
$P ctas_per_sm=3 > gpurun_out/r2_1_probe_cta3.log 2>&1
$P ctas_per_sm=2 > gpurun_out/r2_1_probe_cta2.log 2>&1
for v in pf0 cmp cmp_pf0; do
  ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_$v.so $P > gpurun_out/r2_1_probe_$v.log 2>&1
done
ALLWAVE_CUDA_LIB=allwave_b200/liballwave_cuda_cmp.so $P ctas_per_sm=3 > gpurun_out/r2_1_probe_cmp_cta3.log 2>&1
grep -H "pairs/s" gpurun_out/r2_1_probe_*.log | grep "it=1"
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; tail -c 1500 gpurun_out/r2_1_bench.json
