"""Generates tests/golden/lights_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/Lighting/PolymorphicLight.hlsli (+ LightShaping.hlsli, Utils/Geometry.hlsli,
Utils/Packing.hlsli:16-51, PolymorphicLight.h) compiled in place as C++ through oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_bsdf, mode "lights"): an emissive triangle through
TriangleLight::Store (what LightsBaker writes: PackColor + the half-packed edges, including the float3 round trip of the packed words), TriangleLight::Create, CalcSample,
CalcSolidAnglePdfForMIS, GetPower.  Run in the build container only:   make -C oracle ref && python tests/golden/make_lights_golden.py
  lights_in [M,24]: base 3, edge1 3, edge2 3, radiance 3, random 2, viewer 3, pad   lights_out [M,24]: 8 record words (bit patterns), sample 10, pdf for MIS, power, edge1 3, area"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(20260926); n = 3000
    u = np.zeros((n, 24), np.float32)
    u[:, 0:3] = (rng.random((n, 3)) - 0.5) * np.float32(60)
    u[:, 3:6] = (rng.random((n, 3)) - 0.5) * np.float32(4); u[:, 6:9] = (rng.random((n, 3)) - 0.5) * np.float32(4)
    u[:200, 3:9] *= np.float32(0.02)                                     # centimetre-sized emitters
    u[200:400, 3:9] *= np.float32(25)                                    # 50 m edges: large half exponents in the packed words
    u[:, 9:12] = rng.gamma(2.0, 3.0, (n, 3)).astype(np.float32); u[400:450, 9:12] = 0; u[450:500, 9:12] *= np.float32(1e4)
    u[:, 12:14] = rng.random((n, 2)); u[:, 14:17] = (rng.random((n, 3)) - 0.5) * np.float32(80)
    out = run("lights", u, 24)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lights_golden.npz"), lights_in=u, lights_out=out,
                        source=np.array("Rtxpt/Shaders/PathTracer/Lighting/PolymorphicLight.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape, "nan:", int(np.isnan(out).sum()))
