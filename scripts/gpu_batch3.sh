#!/usr/bin/env bash
# round 2, third GPU batch: re-run of the suites with the measured tolerances, launch lists of the config-3 frame (with / without NEE-AT feedback), bench line of the measured defaults
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/b3_gpu.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/b3_gpu.log
echo "=== unverified suite"; timeout 900 python -m pytest tests -q -m gpu_unverified > gpurun_out/b3_unverified.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/b3_unverified.log
echo "=== config3 launch list"
FRAMES=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_config3_launches.csv python scripts/profile_config3.py > gpurun_out/b3_c3.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/b3_c3.log
NEEAT=0 FRAMES=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_config3_nofeedback_launches.csv python scripts/profile_config3.py > gpurun_out/b3_c3nf.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/b3_c3nf.log
echo "=== config3 unprofiled"; python scripts/profile_config3.py 2>&1 | tail -n 6; NEEAT=0 python scripts/profile_config3.py 2>&1 | tail -n 3
echo "=== bench"; python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/b3_bench.err; echo "rc=$?"; python -c "import json; d=json.loads(open('gpurun_out/r2_bench_a.json').read().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('config3'))"
ls -la gpurun_out | head -50
