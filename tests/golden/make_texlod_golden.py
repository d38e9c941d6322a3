"""Generates tests/golden/texlod_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/Rendering/Materials/TexLODHelpers.hlsli (lines 40-161: SafeLog2, the fp16-packed RayCone
with propagateDistance / addToSpreadAngle / computeLOD, computeRayConeTriangleLODValue) compiled in place as C++ through oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_bsdf, mode
"texlod").  Run in the build container only:   make -C oracle ref && python tests/golden/make_texlod_golden.py
  texlod_in [M,40]: 3 vertices, 3 uvs, 3x3 transform, cone width, spread angle, hitT, ray dir, normal, extra angle, a positive number   texlod_out [M,8]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(31); n = 3000
    u = np.zeros((n, 40), np.float32)
    u[:, 0:9] = (rng.random((n, 9)) - 0.5) * np.float32(6); u[:, 9:15] = rng.random((n, 6)) * np.float32(4)
    rot = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(n)]).astype(np.float32) * rng.uniform(0.3, 3.0, (n, 1, 1)).astype(np.float32); u[:, 15:24] = rot.reshape(n, 9)
    u[:, 24] = rng.random(n) * 0.2; u[:, 25] = rng.random(n) * 0.01; u[:, 26] = rng.random(n) * 80; u[:, 27:33] = rng.normal(size=(n, 6)); u[:, 33] = rng.random(n) * 0.3; u[:, 34] = np.exp(rng.uniform(-90, 90, n))
    u[:40, 9:15] = 0.25                                                   # degenerate texture triangle: Ta = 0 -> SafeLog2 clamps
    u[40:80, 24:26] = 0                                                   # a fresh cone
    out = run("texlod", u, 8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "texlod_golden.npz"), texlod_in=u, texlod_out=out,
                        source=np.array("Rtxpt/Shaders/PathTracer/Rendering/Materials/TexLODHelpers.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape, "nan:", int(np.isnan(out).sum()), "inf:", int(np.isinf(out).sum()))
