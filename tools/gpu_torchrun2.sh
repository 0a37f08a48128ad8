#!/bin/bash
# Two B200s of one box: the driver's weak-scaling launch of bench.py (one rank per GPU under torchrun), both arms.
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $L bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/extra_weak_C2_n2.json 2> gpurun_out/extra_weak_C2_n2.err; cut -c1-200 gpurun_out/extra_weak_C2_n2.json; tail -2 gpurun_out/extra_weak_C2_n2.err
timeout 900 $L bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/extra_weak_C2_n2_reference.json 2> gpurun_out/extra_weak_C2_n2_reference.err; cut -c1-200 gpurun_out/extra_weak_C2_n2_reference.json
timeout 600 python bench.py --scaling strong --config C5 --gpus 2 --steps 1 --warmup 1 > gpurun_out/extra_strong_C5_n2.json 2> gpurun_out/extra_strong_C5_n2.err; cut -c1-200 gpurun_out/extra_strong_C5_n2.json
