#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/one_smoke.log 2>&1; tail -2 gpurun_out/one_smoke.log
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/one_reference.json 2> gpurun_out/one_reference.err; cut -c1-200 gpurun_out/one_reference.json; tail -2 gpurun_out/one_reference.err
