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
    # analytic sphere / spot lights (mode "spheres"): record assembled from the reference's own packers, then SphereLight::Create, PolymorphicLight::CalcSample (incl.
    # evaluateLightShaping), CalcSolidAnglePdfForMIS, GetPower.  in: centre 3, radius, radiance 3, spot?, min-falloff?, axis 3, cos cone, softness, random 2, viewer 3
    rng = np.random.default_rng(20260927)
    v = np.zeros((n, 24), np.float32)
    v[:, 0:3] = (rng.random((n, 3)) - 0.5) * np.float32(40); v[:, 3] = np.float32(0.02) + rng.random(n).astype(np.float32) * np.float32(0.6)
    v[:, 4:7] = rng.gamma(2.0, 30.0, (n, 3)).astype(np.float32); v[:, 7] = rng.random(n); v[:, 8] = rng.random(n)
    v[:, 9:12] = rng.normal(size=(n, 3)).astype(np.float32); v[:, 12] = np.cos(np.radians(rng.uniform(5, 80, n))).astype(np.float32); v[:, 13] = np.float32(0.02) + rng.random(n).astype(np.float32) * np.float32(0.3)
    v[:, 14:16] = rng.random((n, 2)); v[:, 16:19] = (rng.random((n, 3)) - 0.5) * np.float32(60)
    v[:100, 16:19] = v[:100, 0:3] + np.float32(0.01)                     # viewer inside the sphere
    sout = run("spheres", v, 24)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sphere_lights_golden.npz"), spheres_in=v, spheres_out=sout,
                        source=np.array("Rtxpt/Shaders/PathTracer/Lighting/{PolymorphicLight,LightShaping}.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(v.shape, sout.shape, "nan:", int(np.isnan(sout).sum()))
