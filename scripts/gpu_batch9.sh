#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "=== BC probe"; ./scripts/probes/bc_probe
echo "=== tiled hit-dist (aligned boxes): reblur tests, strict + fast"; timeout 600 python -m pytest tests/test_gpu_reblur.py -q -m gpu > gpurun_out/b9_reblur.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/b9_reblur.log
echo "=== config3 timing: tiled on / off"; python scripts/profile_config3.py 2>&1 | tail -n 2; RTXPT_REBLUR_TILED=0 python scripts/profile_config3.py 2>&1 | tail -n 2
echo "=== sanitizer on the tiled pass"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_reblur.py -x -q -m gpu -k "static_camera and True" 2>&1 | tail -n 4
