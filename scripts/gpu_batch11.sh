#!/usr/bin/env bash
# Round-2 artifacts: config-3 full-size parity test (first run), the bench line, the launch list, `ncu --set full` of one whole frame (summarised on the box: the report is too large to bring back)
set -u
mkdir -p gpurun_out
echo "=== config3 full-size parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k config3 -s > gpurun_out/b11_c3.log 2>&1; echo "rc=$?"; grep -a "config3 frame\|passed\|failed\|Error" gpurun_out/b11_c3.log | cut -c1-900
echo "=== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/b11_bench.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2_bench_n1.json
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-realtime > gpurun_out/b11_ll.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2_ncu_launches.csv
echo "=== ncu full, one frame"; timeout 1500 ncu --set full --clock-control none -k regex:"k_generate|k_trace_|k_shade|k_commit" -s 70 -c 35 -o /tmp/r2_full -f python scripts/profile_wavefront.py > gpurun_out/b11_ncu.log 2>&1; echo "rc=$?"
python scripts/ncu_summary.py /tmp/r2_full.ncu-rep gpurun_out/r2_ncu_full_summary.json 3977823 | tail -n 40
echo "=== ncu full with source, k_shade x2 + closest x1 (small report to read here)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace_closest" -s 46 -c 3 -o gpurun_out/r2_shade -f python scripts/profile_wavefront.py > gpurun_out/b11_ncu2.log 2>&1; echo "rc=$?"; ls -la gpurun_out/
