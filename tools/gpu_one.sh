#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -k "legacy_align or aligner_api or match_score" > gpurun_out/one_tests.log 2>&1; echo "exit $?" >> gpurun_out/one_tests.log; tail -15 gpurun_out/one_tests.log
