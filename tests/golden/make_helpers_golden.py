"""Generates tests/golden/helpers_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/PathTracerHelpers.hlsli (lines 26-66, 155-219: ComputeRayOrigin, the grazing-angle
falloff, BalanceHeuristic, the ray-cone growth functions, ComputeNewScatterFireflyFilterK, FireflyFilter, FireflyFilterShort; 221-270: MatrixRotateFromTo) compiled in place as C++ through
oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_bsdf, mode "helpers").  Run in the build container only:   make -C oracle ref && python tests/golden/make_helpers_golden.py
  helpers_in [M,8] uniforms   helpers_out [M,16]   layout: oracle/ref_kat_bsdf_main.cpp"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(20260925)
    u = rng.random((3000, 8), dtype=np.float32)
    u[:40, 0:3] = 0.5                                # positions at the origin (inside the |p| < 1/16 switch on every axis)
    u[40:80, 3] = 0.0; u[80:120, 4] = 0.0            # zero bounce pdf; zero firefly K
    out = run("helpers", u, 16)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "helpers_golden.npz"), helpers_in=u, helpers_out=out,
                        source=np.array("Rtxpt/Shaders/PathTracer/PathTracerHelpers.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape, "nan:", int(np.isnan(out).sum()))
