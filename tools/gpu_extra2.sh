#!/bin/bash
# where a short job's wall time goes: CLI start-up phases on C3, and the load / stream split of the C1 end-to-end step
mkdir -p gpurun_out
ALLWAVE_TIMING=1 timeout 300 python tools/cli_e2e.py C3 1415 > gpurun_out/extra2_cli_C3_timing.log 2>&1; cat gpurun_out/extra2_cli_C3_timing.log
AW_BENCH_TRACE=1 timeout 300 python bench.py --config C1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/extra2_bench_C1.json 2> gpurun_out/extra2_bench_C1.err; tail -5 gpurun_out/extra2_bench_C1.err
