"""Generates tests/golden/bsdf_golden.npz from the UNMODIFIED reference material headers
  /root/reference/Rtxpt/Shaders/PathTracer/Rendering/Materials/{Fresnel,Microfacet,BxDF,StandardBSDF,IBSDF}.hlsli, Utils/Math/MathHelpers.hlsli, Scene/ShadingData.hlsli ...
compiled in place as C++ (oracle/ref_hlsl_shim.h + oracle/ref_hlsl_tu.sh -> oracle/_ref/ref_kat_bsdf).  Run in the build container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_bsdf_golden.py
The committed vectors pin the floating-point material model: StandardBSDF eval / evalPdf / sample / getLobes / evalDeltaLobes / estimateSpecDiffBSDF on
seeded records (tests/bsdf_records.py) and the scalar building blocks (Fresnel, GGX NDF / masking / bounded-VNDF sampling, hemisphere / disk sampling, octahedral maps).
  bsdf_in  [N,36]  bsdf_out  [N,40]   layouts: oracle/ref_kat_bsdf_main.cpp
  funcs_in [M,8]   funcs_out [M,40]"""
import os
import subprocess
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bsdf_records import make_records  # noqa: E402


def run(mode, arr, width, exe="ref_kat_bsdf"):
    exe = os.path.join(ROOT, "oracle", "_ref", exe)
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "in.f32"), os.path.join(d, "out.f32")
        np.ascontiguousarray(arr, np.float32).tofile(a)
        subprocess.run([exe, mode, a, b], check=True)
        return np.fromfile(b, np.float32).reshape(-1, width)


def generate():
    rng = np.random.default_rng(20260923)
    rec = np.concatenate([make_records(rng, 2200), make_records(rng, 1000, "transmissive"), make_records(rng, 400, "opaque")])
    # edge cases the reference tests by construction: grazing view, wo in the lower hemisphere / on the horizon, roughness at the delta threshold, eta = 1, single active lobes
    e = make_records(rng, 400)
    e[:50, 21] = np.float16(0.08)                               # alpha = 0.0064: the kMinGGXAlpha boundary (fp16 roughness on either side)
    e[50:100, 21] = np.float16(0.0799); e[100:150, 31] = 1.0    # just below; eta == 1 switches rough transmission to the delta lobe
    e[150:200, 12:15] = -e[150:200, 3:6]                        # wo = -N
    e[200:250, 12:15] = e[200:250, 6:9]                         # wo on the horizon (cos = 0 < kMinCosTheta)
    for k, lobes in enumerate((0x01, 0x02, 0x04, 0x10, 0x20, 0x40, 0x11, 0x22, 0x44, 0x0F, 0xF0, 0x33)): e[250 + 12 * k: 262 + 12 * k, 33] = lobes
    rec = np.concatenate([rec, e]).astype(np.float32)
    u = rng.random((2000, 8), dtype=np.float32)
    u[:16, 1] = 0.0; u[16:32, 2] = 0.0; u[32:48, 0] = 0.0; u[48:64, 4] = 0.0      # grazing cosines, alpha at its floor, eta = 1
    return rec, run("bsdf", rec, 40), u, run("funcs", u, 40)


if __name__ == "__main__":
    rec, out, u, fout = generate()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bsdf_golden.npz"), bsdf_in=rec, bsdf_out=out, funcs_in=u, funcs_out=fout,
                        source=np.array("Rtxpt/Shaders/PathTracer/Rendering/Materials/*.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(rec.shape, out.shape, u.shape, fout.shape, "nan in bsdf_out:", int(np.isnan(out).sum()), "nan in funcs_out:", int(np.isnan(fout).sum()))
