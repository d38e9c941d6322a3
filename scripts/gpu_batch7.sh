#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "=== TMA probe"; for w in 1 2 3; do WHICH=$w ./scripts/probes/tma_probe 2>&1 | tail -n 2; done
echo "=== BC textures"; timeout 600 python -m pytest tests/test_dds.py -q -m gpu > gpurun_out/b7_bc.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/b7_bc.log
