#!/usr/bin/env bash
# round 2, second GPU batch: the whole GPU suites without -x (every failure in one go), A/B timing of the traversal variants and of the pipeline lanes, ncu of the ReBLUR passes
set -u
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-realtime"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); k=d['roofline']['kernel_ms_per_frame']; print('%-28s %8.1f Mrays/s %7.3f ms/frame | closest %.3f shadow %.3f shade %.3f' % (sys.argv[1], d['value'], d['ms_per_step'], k['trace_closest'], k['trace_shadow'], k['shade']))" "$1"; }
echo "=== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/b2_gpu.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/b2_gpu.log
echo "=== unverified suite"; timeout 900 python -m pytest tests -q -m gpu_unverified > gpurun_out/b2_unverified.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/b2_unverified.log
echo "=== A/B (lanes = default 2)"
for v in base f0i1 f0i0 f1i0; do
  if [ $v = base ]; then $B 2>gpurun_out/b2_$v.err | pick $v; else RTXPT_LIB=$PWD/rtxpt_b200/csrc/_build/librtxpt_b200_var_$v.so $B 2>gpurun_out/b2_$v.err | pick $v; fi
done | tee gpurun_out/b2_ab.txt
echo "=== lanes"
for l in 1 2 4; do RTXPT_LANES=$l RTXPT_LIB=$PWD/rtxpt_b200/csrc/_build/librtxpt_b200_var_f0i0.so $B 2>>gpurun_out/b2_lanes.err | pick "f0i0 lanes=$l"; done | tee gpurun_out/b2_lanes.txt
for l in 1 2 4; do RTXPT_LANES=$l $B 2>>gpurun_out/b2_lanes.err | pick "base lanes=$l"; done | tee -a gpurun_out/b2_lanes.txt
echo "=== ncu reblur"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rb_ -s 168 -c 24 -f -o gpurun_out/r2_reblur python scripts/profile_reblur.py > gpurun_out/b2_ncu_reblur.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/b2_ncu_reblur.log
ls -la gpurun_out | head -40
