#!/bin/bash
# 8 GPUs: strong scaling through the product path (one process, N contexts, shared chunk queue): C2 1/2/4/8, C5 and C3 1/8, full C4 at 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # config gpus pairs extra
  timeout 1500 python bench.py --scaling strong --config $1 --gpus $2 --steps 1 --warmup 1 --pairs $3 $4 > gpurun_out/r2_13_strong_$1_n$2.json 2> gpurun_out/r2_13_strong_$1_n$2.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_13_strong_$1_n$2.json')); print('$1 N=$2', 'pairs', d['config']['pairs_total'], 'pairs/s', round(d['value'],1), 's', round(d['ms_per_step']/1e3,2), 'imb', round(d['gpu_imbalance_max_over_mean'],3), 'digest', d['paf_digest'], 'setup', round(d['setup_seconds'],1))
except Exception as e:
    print('$1 N=$2 FAILED', e); print(open('gpurun_out/r2_13_strong_$1_n$2.err').read()[-600:])
PY
}
run C2 1 151552 ""
run C2 2 151552 "--no-checksum"
run C2 4 151552 "--no-checksum"
run C2 8 151552 ""
run C3 1 2000810 "--no-checksum"
run C3 8 2000810 "--no-checksum"
run C5 1 303104 "--no-checksum"
run C5 8 303104 "--no-checksum"
run C4 8 0 "--no-checksum"
