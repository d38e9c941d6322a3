"""Generates tests/golden/utils_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/Utils/Utils.hlsli (lines 68-92, 115-169, 392-499, 510-517: LuminanceClamp, Reinhard, the
octahedral encodings incl. the 32- and 30-bit packings, EvalMIS, RelativelyEqual, FastSqrt / FastACos, WeightedAverage) compiled in place as C++ through oracle/ref_hlsl_shim.h
(oracle/_ref/ref_kat_bsdf, mode "utils").  Run in the build container only:   make -C oracle ref && python tests/golden/make_utils_golden.py
  utils_in [M,8] uniforms   utils_out [M,24]   layout: oracle/ref_kat_bsdf_main.cpp"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(20260924)
    u = rng.random((3000, 8), dtype=np.float32)
    u[:32, 0:3] = 0.5; u[32:64, 2] = 0.5 - 1e-3; u[64:96, 0] = 0.0; u[96:128, 1] = 1.0; u[128:160, 5] = 0.0; u[160:192, 7] = 0.0        # axis directions, z near 0, zero pdfs
    out = run("utils", u, 24)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "utils_golden.npz"), utils_in=u, utils_out=out,
                        source=np.array("Rtxpt/Shaders/PathTracer/Utils/Utils.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape, "nan:", int(np.isnan(out).sum()))
