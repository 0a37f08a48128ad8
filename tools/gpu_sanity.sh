#!/bin/bash
# Last check of a round: the GPU suite (with the vector-test plumbing fed from the ORACLE's answers, build/oracle_vectors.tsv.gz
# made by tools/wfa2_vectors/dump_oracle_vectors.py), smoke(), and short bench runs of both arms.
mkdir -p gpurun_out
AW_WFA2_VECTORS=$PWD/build/oracle_vectors.tsv.gz timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/sanity_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/sanity_tests.log; tail -4 gpurun_out/sanity_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/sanity_smoke.log 2>&1; tail -2 gpurun_out/sanity_smoke.log
timeout 300 python bench.py --config C3 --steps 3 --warmup 3 > gpurun_out/sanity_bench_C3.json 2> gpurun_out/sanity_bench_C3.err; cut -c1-160 gpurun_out/sanity_bench_C3.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/sanity_bench_C2.json 2> gpurun_out/sanity_bench_C2.err; cut -c1-160 gpurun_out/sanity_bench_C2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/sanity_bench_C2_reference.json 2> gpurun_out/sanity_bench_C2_reference.err; cut -c1-160 gpurun_out/sanity_bench_C2_reference.json
