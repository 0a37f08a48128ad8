#!/bin/bash
# round-2 GPU call 3: full parity suites (old + new), bench lines for C2 (default), C1, C3, C5, strong-scaling probe on 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_3_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_3_tests.log
tail -15 gpurun_out/r2_3_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_3_bench_C2.json 2> gpurun_out/r2_3_bench_C2.err; tail -c 600 gpurun_out/r2_3_bench_C2.json; tail -3 gpurun_out/r2_3_bench_C2.err
for c in C1 C3 C5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2_3_bench_$c.json 2> gpurun_out/r2_3_bench_$c.err; tail -c 900 gpurun_out/r2_3_bench_$c.json; tail -3 gpurun_out/r2_3_bench_$c.err
done
timeout 600 python bench.py --scaling strong --gpus 1 --steps 1 --warmup 1 --pairs 37888 > gpurun_out/r2_3_strong_C2_n1.json 2> gpurun_out/r2_3_strong_C2_n1.err; cat gpurun_out/r2_3_strong_C2_n1.json; tail -3 gpurun_out/r2_3_strong_C2_n1.err
timeout 600 python bench.py --scaling strong --config C3 --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_3_strong_C3_n1.json 2> gpurun_out/r2_3_strong_C3_n1.err; cat gpurun_out/r2_3_strong_C3_n1.json; tail -3 gpurun_out/r2_3_strong_C3_n1.err
