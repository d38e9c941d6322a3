#!/usr/bin/env bash
# One gpurun call that takes the GPU-unverified code of round 1 through its first runs, cheapest and most isolated first, each step under its own timeout so that a hang or a
# fault costs one step.  Logs land in gpurun_out/ (merged back by gpurun).  Usage on the GPU box:   bash scripts/gpu_verify_round2.sh [per-step timeout seconds, default 240]
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_verify_round2.sh'
# Order = DESIGN.md §9: host-verified bodies first (their kernels add only launch shapes / atomics / libdevice), then the path-tracer integrations, then timing.
set -u
T=${1:-240}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$T" "$@" > "gpurun_out/r2_$name.log" 2>&1; echo "rc=$? ($name)"; tail -n 3 "gpurun_out/r2_$name.log"; }
PY="python -m pytest -x -q -m gpu_unverified"
run verified      python -m pytest tests -x -q -m gpu                                   # the round-1 bar first: nothing regressed
run guide_filter  $PY tests/test_gpu_reblur.py -k spec_hit_t
run envbake       $PY tests/test_gpu_envbake.py
run tonemap       $PY tests/test_gpu_tonemap.py
run refit         $PY tests/test_gpu_refit.py
run skinning      $PY tests/test_gpu_skinning.py
run reblur        $PY tests/test_gpu_reblur.py -k "static_camera or moving_camera or reset"
run realtime_rest $PY tests/test_gpu_realtime.py
run denoise_e2e   $PY tests/test_gpu_reblur.py -k denoise_realtime
run neeat_baker   $PY tests/test_gpu_neeat.py -k baker_passes
run neeat_api     $PY tests/test_gpu_neeat.py -k api_errors
run neeat_loop    $PY tests/test_gpu_neeat.py -k "unbiased or reference_mode"
run cpp_example   $PY tests/test_gltf_loader.py -k realtime_example
run bench         python bench.py --steps 4 --warmup 3                                   # its realtime child times realtime mode, the denoised frame and the NEE-AT loop
run sanitizer     compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest -x -q -m gpu_unverified tests/test_gpu_reblur.py -k "static_camera and True" tests/test_gpu_neeat.py -k baker_passes
echo "=== done"; grep -h "^rc=" /dev/null; ls -la gpurun_out | head -40
