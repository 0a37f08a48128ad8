#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "stream or cancel or sentinel or orientation or divergence or run_job or c1_full" > gpurun_out/r2_4_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_4_tests.log
tail -30 gpurun_out/r2_4_tests.log
timeout 600 python bench.py --scaling strong --config C3 --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_4_strong_C3_n1.json 2> gpurun_out/r2_4_strong_C3_n1.err; cut -c1-200 gpurun_out/r2_4_strong_C3_n1.json; tail -3 gpurun_out/r2_4_strong_C3_n1.err
timeout 600 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_4_bench_C3.json 2> gpurun_out/r2_4_bench_C3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_4_bench_C3.json')); print('C3 value', d['value'], 'e2e', d['e2e'])"; tail -3 gpurun_out/r2_4_bench_C3.err
timeout 600 python bench.py --scaling strong --gpus 1 --steps 1 --warmup 1 --pairs 37888 > gpurun_out/r2_4_strong_C2_n1.json 2> gpurun_out/r2_4_strong_C2_n1.err; cut -c1-200 gpurun_out/r2_4_strong_C2_n1.json; tail -3 gpurun_out/r2_4_strong_C2_n1.err
