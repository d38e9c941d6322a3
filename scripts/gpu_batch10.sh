#!/usr/bin/env bash
# Round-2 verification batch: full GPU suite on the current tree, the bench line, the launch list and one `ncu --set full` capture of the wavefront kernels.
set -u
mkdir -p gpurun_out
echo "=== gpu suite"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/b10_gpu.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/b10_gpu.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
echo "=== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/b10_bench.err; echo "rc=$?"; cut -c1-600 gpurun_out/r2_bench_n1.json
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-realtime > gpurun_out/b10_ll.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2_ncu_launches.csv
echo "=== ncu full (first iterations of one frame)"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_generate|k_trace_|k_shade|k_commit" -s 70 -c 35 -o gpurun_out/r2_full -f python scripts/profile_wavefront.py > gpurun_out/b10_ncu.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r2_full.ncu-rep
