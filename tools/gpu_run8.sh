#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_8_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_8_tests.log
tail -25 gpurun_out/r2_8_tests.log
python tools/perf_probe.py C2 60 2368 > gpurun_out/r2_8_probe_default.log 2>&1; tail -1 gpurun_out/r2_8_probe_default.log
python tools/perf_probe.py C5 80 4736 > gpurun_out/r2_8_probe_c5.log 2>&1; tail -1 gpurun_out/r2_8_probe_c5.log
python tools/perf_probe.py C1 16 240 > gpurun_out/r2_8_probe_c1.log 2>&1; tail -1 gpurun_out/r2_8_probe_c1.log
