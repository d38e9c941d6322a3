"""Generates tests/golden/tonemap_golden.npz from the UNMODIFIED Rtxpt/ToneMapper/ToneMapping.ps.hlsli (lines 31-129: calcLuminance, the six operators, toneMap) and
ToneMapping_cb.h compiled in place as C++ through oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_bsdf, mode "tonemap").  Run in the build container only:
    make -C oracle ref && python tests/golden/make_tonemap_golden.py
  tonemap_in [M,8]: rgb, operator 0-5, whiteMaxLuminance, whiteScale, pad   tonemap_out [M,4]: rgb, luminance"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(5); n = 3000
    u = np.zeros((n, 8), np.float32); u[:, 0:3] = rng.gamma(1.5, 1.2, (n, 3)).astype(np.float32); u[:, 3] = rng.integers(0, 6, n); u[:, 4] = 1.0 + rng.random(n) * 5; u[:, 5] = 5 + rng.random(n) * 10
    u[:60, 0:3] = 0; u[60:120, 0:3] *= np.float32(1e-3); u[120:180, 0:3] *= np.float32(200)          # black (0 / 0 in Reinhard: NaN in the reference too), toe, far shoulder
    out = run("tonemap", u, 4)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tonemap_golden.npz"), tonemap_in=u, tonemap_out=out,
                        source=np.array("Rtxpt/ToneMapper/ToneMapping.ps.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape, "nan:", int(np.isnan(out).sum()))
