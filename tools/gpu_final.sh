#!/bin/bash
# Round-end measurement on one B200: parity suites, the bench lines of every config (weak form, N=1), the reference arm,
# and the ncu evidence of the default bench command (launch list + full capture of the alignment kernel).
# Results: gpurun_out/final_*; copy into profiles/r02_runs/ and run tools/make_profiles.py + tools/fill_baseline_table.py.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_C2.json 2> gpurun_out/final_bench_C2.err; cut -c1-140 gpurun_out/final_bench_C2.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/final_bench_C2_reference.json 2> gpurun_out/final_bench_C2_reference.err; cut -c1-140 gpurun_out/final_bench_C2_reference.json
for c in C1 C3 C5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err; cut -c1-140 gpurun_out/final_bench_$c.json
done
timeout 900 python bench.py --config C4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final_bench_C4.json 2> gpurun_out/final_bench_C4.err; cut -c1-140 gpurun_out/final_bench_C4.json; tail -2 gpurun_out/final_bench_C4.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/final_ncu_launches.log 2>&1
# NOTE: the full-set capture replays the 4 s kernel ~40 times with a save/restore of its multi-GB workspace: it took 32 minutes of box
# time in round 2 (everything above it: 10 minutes).  `--batch 2368` would cut that sixfold at the cost of a longer tail in the profile.
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$B1 > gpurun_out/final_plain1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:aw_align_kernel -s 3 -c 1 -o gpurun_out/prof_final_align $B1 > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_ncu_full.log
