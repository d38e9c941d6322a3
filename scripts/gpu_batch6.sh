#!/usr/bin/env bash
# round 2, sixth GPU batch: the tiled HitDistReconstruction under compute-sanitizer (first run faulted with "illegal instruction"), LPT shadow ordering A/B, full suite
set -u
mkdir -p gpurun_out
echo "=== tiled hit-dist under compute-sanitizer"
RTXPT_REBLUR_TILED=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_reblur.py -x -q -m gpu -k "static_camera and True" > gpurun_out/b6_sanitizer.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/b6_sanitizer.log | head -60
echo "=== tiled hit-dist plain"; RTXPT_REBLUR_TILED=1 timeout 600 python -m pytest tests/test_gpu_reblur.py -x -q -m gpu > gpurun_out/b6_tiled.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/b6_tiled.log
echo "=== gpu suite (tiled off)"; timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/b6_gpu.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/b6_gpu.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-realtime"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); k=d['roofline']['kernel_ms_per_frame']; print('%-34s %8.1f Mrays/s %7.3f ms/frame | closest %.3f shadow %.3f shade %.3f' % (sys.argv[1], d['value'], d['ms_per_step'], k['trace_closest'], k['trace_shadow'], k['shade']))" "$1"; }
echo "=== LPT A/B"
for v in 1 0; do RTXPT_SHADOW_LPT=$v $B 2>>gpurun_out/b6.err | pick "N=1 shadow LPT=$v"; done | tee gpurun_out/b6_lpt.txt
for v in 1 0; do RTXPT_BENCH_EMULATE_WORLD=8 RTXPT_SHADOW_LPT=$v $B 2>>gpurun_out/b6.err | pick "rank of 8, shadow LPT=$v"; done | tee -a gpurun_out/b6_lpt.txt
echo "=== config3"; python scripts/profile_config3.py 2>&1 | tail -n 2; RTXPT_SHADOW_LPT=0 python scripts/profile_config3.py 2>&1 | tail -n 1
