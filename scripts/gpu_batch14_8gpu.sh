#!/usr/bin/env bash
# 8 GPUs: the headline bench (phase breakdown per rank) and config 3 (run with gpurun --gpus 8)
set -u
mkdir -p gpurun_out
echo "=== bench N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-realtime 2> gpurun_out/b14_n8.err | tee gpurun_out/r2_bench_n8.json | cut -c1-400
echo "=== bench N=4"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline --no-realtime 2> gpurun_out/b14_n4.err | tee gpurun_out/r2_bench_n4.json | cut -c1-400
echo "=== config 3, 8 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 scripts/bench_config3_mgpu.py --frames 10 2> gpurun_out/b14_c3_n8.err | tee gpurun_out/r2_config3_n8.json | cut -c1-700
tail -n 3 gpurun_out/b14_n8.err gpurun_out/b14_c3_n8.err | cut -c1-300
