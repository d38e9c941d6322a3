"""Generates tests/golden/interior_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/Rendering/Materials/InteriorList.hlsli (the two-slot stack of nested dielectrics)
compiled in place as C++ through oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_bsdf, mode "interior").  Run in the build container only:
    make -C oracle ref && python tests/golden/make_interior_golden.py
  interior_in [M,48]: 12 crossings x (material, nested priority, entering, probe priority)   interior_out [M,72]: after each crossing: slots (2 bit patterns), top priority,
  top material, next material, isTrueIntersection( probe ) before the crossing"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(77); n = 4000
    u = np.zeros((n, 12, 4), np.float32)
    u[..., 0] = rng.integers(0, 6, (n, 12)); u[..., 1] = rng.integers(0, 16, (n, 12)); u[..., 2] = rng.integers(0, 2, (n, 12)); u[..., 3] = rng.integers(0, 16, (n, 12))
    # well-formed enter / leave pairs on half of the records (what closed meshes produce); the rest random (overflowing stacks, leaving what was never entered)
    for i in range(n // 2):
        mats = rng.integers(0, 5, 6); pri = rng.integers(1, 16, 6)
        seq = [(mats[k], pri[k], 1) for k in range(6)] + [(mats[k], pri[k], 0) for k in rng.permutation(6)]
        for k, (m, p, e) in enumerate(seq): u[i, k, 0:3] = (m, p, e)
    u = u.reshape(n, 48)
    out = run("interior", u, 72)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "interior_golden.npz"), interior_in=u, interior_out=out,
                        source=np.array("Rtxpt/Shaders/PathTracer/Rendering/Materials/InteriorList.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape)
