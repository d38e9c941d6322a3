#!/usr/bin/env bash
# round 2, 2-GPU batch: the C++ multi-GPU host against the single-context frame, the realtime fast-build tests, the bench at N=2 with its per-phase breakdown
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== multi-GPU + realtime tests"; timeout 900 python -m pytest tests/test_mgpu_host.py tests/test_gpu_realtime.py tests/test_gpu_reblur.py -q -m gpu > gpurun_out/b5_tests.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/b5_tests.log
echo "=== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/b5_bench.err; echo "rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n2.json').read().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['phases'])"
echo "=== C++ example on 2 GPUs"; python - <<'PY'
import os, sys, subprocess, tempfile
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gltf_export
from rtxpt_b200 import scenes, lib
d = tempfile.mkdtemp(); path = gltf_export.export(scenes.cornell_builder(), os.path.join(d, 'c.gltf'), camera=dict(position=(2.78, 2.73, -8.0), direction=(0, 0, 1), up=(0, 1, 0), yfov=0.66, znear=0.1, zfar=1e7))
exe = os.path.join(os.path.dirname(lib.LIB_PATH), 'multigpu_gltf')
for g in ('1', '2'):
    r = subprocess.run([exe, path, os.path.join(d, 'o%s.pfm' % g), g, '512', '512', '64', '4'], capture_output=True, text=True); print(g, r.returncode, r.stderr.strip())
a = open(os.path.join(d, 'o1.pfm'), 'rb').read(); b = open(os.path.join(d, 'o2.pfm'), 'rb').read(); print('2-GPU frame == 1-GPU frame:', a == b)
PY
