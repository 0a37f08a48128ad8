#!/bin/bash
# two GPUs: the tests that need more than one device
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "multi_gpu or run_job or stream_blocks" > gpurun_out/two_gpu_tests.log 2>&1; echo "exit $?" >> gpurun_out/two_gpu_tests.log; tail -4 gpurun_out/two_gpu_tests.log
