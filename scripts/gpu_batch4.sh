#!/usr/bin/env bash
# round 2, fourth GPU batch: the promoted -m gpu suite, realtime-shade occupancy A/B on the config-3 frame, per-rank frame time of an 8-GPU job with 1/2/4 pipeline lanes
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1800 python -m pytest tests -q -m gpu --durations=12 > gpurun_out/b4_gpu.log 2>&1; echo "rc=$?"; tail -n 22 gpurun_out/b4_gpu.log
echo "=== config3: default (3 rt-shade CTAs/SM, fast-math ReBLUR)"; python scripts/profile_config3.py 2>&1 | tail -n 2
echo "=== config3: 4 rt-shade CTAs/SM"; RTXPT_LIB=$PWD/rtxpt_b200/csrc/_build/librtxpt_b200_var_rt4.so python scripts/profile_config3.py 2>&1 | tail -n 2
echo "=== one rank of an 8-GPU job (emulated on one GPU), lanes 1/2/4"
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-realtime"
for l in 1 2 4; do RTXPT_BENCH_EMULATE_WORLD=8 RTXPT_LANES=$l $B 2>>gpurun_out/b4.err | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); k=d['roofline']['kernel_ms_per_frame']; print('world 8 rank 0, lanes=%s: %.3f ms/frame (serialised kernels: closest %.3f shadow %.3f shade %.3f other %.3f)' % (sys.argv[1], d['ms_per_step'], k['trace_closest'], k['trace_shadow'], k['shade'], k['other']))" $l; done | tee gpurun_out/b4_world8.txt
for l in 1 2; do RTXPT_BENCH_EMULATE_WORLD=2 RTXPT_LANES=$l $B 2>>gpurun_out/b4.err | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('world 2 rank 0, lanes=%s: %.3f ms/frame' % (sys.argv[1], d['ms_per_step']))" $l; done | tee -a gpurun_out/b4_world8.txt
