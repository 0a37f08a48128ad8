#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "cluster" > gpurun_out/r2_5_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_5_tests.log
tail -30 gpurun_out/r2_5_tests.log
if grep -q "tests exit 0" gpurun_out/r2_5_tests.log; then
  timeout 600 python tools/c4_probe.py 8 200000 8 > gpurun_out/r2_5_c4_200k_cluster.log 2>&1; tail -3 gpurun_out/r2_5_c4_200k_cluster.log
  timeout 600 python tools/c4_probe.py 8 200000 8 cluster_min_len=0 > gpurun_out/r2_5_c4_200k_single.log 2>&1; tail -3 gpurun_out/r2_5_c4_200k_single.log
  timeout 900 python tools/c4_probe.py 6 1000000 4 > gpurun_out/r2_5_c4_1m_cluster.log 2>&1; tail -3 gpurun_out/r2_5_c4_1m_cluster.log
  timeout 900 python tools/c4_probe.py 6 1000000 4 solo_len=16384 > gpurun_out/r2_5_c4_1m_cluster_solo16k.log 2>&1; tail -3 gpurun_out/r2_5_c4_1m_cluster_solo16k.log
  timeout 900 python tools/c4_probe.py 6 1000000 4 solo_len=262144 > gpurun_out/r2_5_c4_1m_cluster_solo256k.log 2>&1; tail -3 gpurun_out/r2_5_c4_1m_cluster_solo256k.log
fi
