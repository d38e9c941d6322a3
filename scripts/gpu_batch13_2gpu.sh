#!/usr/bin/env bash
# config 3 on 1 and 2 GPUs (run with gpurun --gpus 2)
set -u
mkdir -p gpurun_out
echo "=== config 3, 1 GPU"; timeout 900 python scripts/bench_config3_mgpu.py --frames 10 2> gpurun_out/b13_c3_n1.err | tee gpurun_out/r2_config3_n1.json | cut -c1-700
echo "=== config 3, 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_config3_mgpu.py --frames 10 2> gpurun_out/b13_c3_n2.err | tee gpurun_out/r2_config3_n2.json | cut -c1-700
tail -n 5 gpurun_out/b13_c3_n2.err | cut -c1-300
