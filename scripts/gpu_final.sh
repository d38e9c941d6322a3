#!/usr/bin/env bash
# what the driver runs at round end, in that order: GPU tests, smoke, the bench line (+ the reference arm)
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_gpu.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/final_gpu.log | cut -c1-300
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/final_bench.err; echo "rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches']); print({k: d['roofline'][k] for k in ('bound','frac','traffic_capture_commit','issue_active','thread_inst_per_ray') if k in d['roofline']}); print(d['cpu_baseline']['value'], d.get('config3',{}).get('frame_ms'))"
