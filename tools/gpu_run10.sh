#!/bin/bash
# 2 GPUs: multi-GPU host path (CLI --gpus 2, run_job n_gpus=2), strong scaling 1 -> 2 through the product path
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or run_job" > gpurun_out/r2_10_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_10_tests.log
tail -5 gpurun_out/r2_10_tests.log
for n in 1 2; do
  timeout 900 python bench.py --scaling strong --gpus $n --steps 1 --warmup 1 --pairs 75776 > gpurun_out/r2_10_strong_C2_n$n.json 2> gpurun_out/r2_10_strong_C2_n$n.err; cut -c1-160 gpurun_out/r2_10_strong_C2_n$n.json; tail -2 gpurun_out/r2_10_strong_C2_n$n.err
done
python -c "
import json
a=json.load(open('gpurun_out/r2_10_strong_C2_n1.json')); b=json.load(open('gpurun_out/r2_10_strong_C2_n2.json'))
print('N1', a['value'], 'N2', b['value'], 'ratio', b['value']/a['value'], 'digest equal', a['paf_digest']==b['paf_digest'], 'imbalance', b['gpu_imbalance_max_over_mean'])"
