#!/usr/bin/env bash
# motion vectors / block heuristics on the GPU (first run), then the whole suite (loadSurface became a template), then the bench line against the round-2 ncu capture
set -u
mkdir -p gpurun_out
echo "=== motion vector tests"; timeout 900 python -m pytest tests/test_motion_vectors.py -q -m gpu > gpurun_out/b12_mv.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/b12_mv.log | cut -c1-400
echo "=== config-3 split over two ranks (one GPU)"; timeout 900 python -m pytest tests/test_gpu_realtime.py -q -m gpu -k two_ranks > gpurun_out/b12_split.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/b12_split.log | cut -c1-400
echo "=== gpu suite"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/b12_gpu.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/b12_gpu.log | cut -c1-400
echo "=== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/b12_bench.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2_bench_n1_b.json
echo "=== A/B: next-item prefetch in k_shade (off / L2 / L1), twice each"
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-realtime"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); k=d['roofline']['kernel_ms_per_frame']; print('%-28s %8.1f Mrays/s %7.3f ms/frame | closest %.3f shadow %.3f shade %.3f' % (sys.argv[1], d['value'], d['ms_per_step'], k['trace_closest'], k['trace_shadow'], k['shade']))" "$1"; }
for rep in 1 2; do for v in base pf1 pf2; do
  if [ $v = base ]; then $B 2>gpurun_out/b12_$v.err | pick $v; else RTXPT_LIB=$PWD/rtxpt_b200/csrc/_build/librtxpt_b200_var_$v.so $B 2>gpurun_out/b12_$v.err | pick $v; fi
done; done | tee gpurun_out/b12_ab.txt
