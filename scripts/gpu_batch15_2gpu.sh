#!/usr/bin/env bash
# 2 GPUs: the C++ multi-GPU host tests (incl. the realtime frame) and the bench line with the NVML clock sampler
set -u
mkdir -p gpurun_out
echo "=== C++ multi-GPU host tests"; timeout 900 python -m pytest tests/test_mgpu_host.py -q -m gpu > gpurun_out/b15_mgpu.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/b15_mgpu.log | cut -c1-300
echo "=== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-realtime 2> gpurun_out/b15_n2.err | tee gpurun_out/r2_bench_n2_nvml.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['value'], d['ms_per_step'], d['clocks'], d['phases']['ms_max_over_ranks'])"
