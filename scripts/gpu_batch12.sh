#!/usr/bin/env bash
# motion vectors / block heuristics on the GPU (first run), then the whole suite (loadSurface became a template), then the bench line against the round-2 ncu capture
set -u
mkdir -p gpurun_out
echo "=== motion vector tests"; timeout 900 python -m pytest tests/test_motion_vectors.py -q -m gpu > gpurun_out/b12_mv.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/b12_mv.log | cut -c1-400
echo "=== gpu suite"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/b12_gpu.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/b12_gpu.log | cut -c1-400
echo "=== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/b12_bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2_bench_n1_b.json
