#!/bin/bash
# One B200: the CLI end to end (FASTA in, PAF file out) on C3 / C5-shaped inputs.
mkdir -p gpurun_out
for i in 1 2; do timeout 300 python tools/cli_e2e.py C3 1415 > gpurun_out/extra_cli_C3_$i.log 2>&1; tail -2 gpurun_out/extra_cli_C3_$i.log; done
timeout 300 python tools/cli_e2e.py C5 1000 > gpurun_out/extra_cli_C5.log 2>&1; tail -2 gpurun_out/extra_cli_C5.log
