"""Generates tests/golden/baker_golden.npz from the UNMODIFIED Rtxpt/Lighting/LightsBaker.hlsl (NEE-AT's feedback passes ProcessFeedbackHistoryP0, P1a, P1b, P2 / FillTile and
ClearFeedbackHistory - the passes whose threads are independent) compiled in place as C++ through oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_baker).  Run in the build container only:
    make -C oracle ref && python tests/golden/make_baker_golden.py
  baker_in [M,3056], baker_out [M,4241]: layouts in oracle/ref_kat_baker_main.cpp.  One record = one frame end on a 16 x 16 image (3 x 3 tiles, 8 x 8 blended image, 16 lights):
  the reservoirs NEE filled, last frame's depth, this frame's depth and motion vectors, last frame's tile lists, the global proxies, a past-to-current light table."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402
from make_sampler_golden import pack_tile  # noqa: E402

if __name__ == "__main__":
    rng = np.random.default_rng(31337); n = 200; P = 256
    r = np.zeros((n, 3056), np.float32)
    for i in range(n):
        total = rng.integers(4, 17); hist = rng.integers(4, 17)
        r[i, 0] = total; r[i, 1] = hist; r[i, 2] = rng.integers(0, 1000); r[i, 3:7] = rng.integers(0, 8, 4); r[i, 7] = rng.random() < 0.85; r[i, 8] = rng.random() < 0.75
        r[i, 10] = 1.5; r[i, 11] = rng.random() < 0.8; r[i, 12] = rng.choice(np.float32([0.005, 0.05, 0.5]))
        remap = rng.integers(0, 18, 16).astype(np.uint32); remap[remap >= 16] = 0xFFFFFFFF          # entries past this frame's list and lights that are gone
        ident = rng.random() < 0.5
        if ident: remap = np.arange(16, dtype=np.uint32)
        r[i, 32:48] = remap.view(np.float32)
        counts = rng.integers(0, 9, total) * (rng.random(total) < 0.8)
        if counts.sum() == 0: counts[0] = 3
        while counts.sum() > 64: counts[np.argmax(counts)] -= 1
        idx = np.repeat(np.arange(total), counts); r[i, 9] = len(idx); r[i, 48:48 + len(idx)] = idx
        # reservoirs: empty, screen-space coherent (flag bit 31), world-space coherent, lights past the historic count
        w = rng.gamma(1.0, 2.0, P).astype(np.float32) * (rng.random(P) < 0.7); cand = rng.integers(0, 18, P).astype(np.uint32) | ((rng.random(P) < 0.6).astype(np.uint32) << 31)
        cand[rng.random(P) < 0.05] = 0xFFFFFFFF
        r[i, 112:368] = w; r[i, 368:624] = cand.view(np.float32)
        hd = np.exp(rng.uniform(0, 3, P)).astype(np.float32); r[i, 624:880] = hd; d = hd * np.where(rng.random(P) < 0.8, 1 + rng.normal(0, 0.05, P), rng.uniform(1.6, 3, P)); r[i, 880:1136] = d.astype(np.float32)
        mv = np.zeros((P, 3), np.float32); mv[:, :2] = np.float16(rng.normal(0, 2.5, (P, 2)) * (rng.random((P, 1)) < 0.7)); mv[:, 2] = np.float16(rng.normal(0, 0.1, P)); r[i, 1136:1904] = mv.reshape(-1)
        tiles = np.stack([pack_tile(rng.choice(rng.choice(hist, rng.integers(1, min(hist, 8) + 1), replace=False), 128)) for _ in range(9)]); r[i, 1904:3056] = tiles.reshape(-1).view(np.float32)
    out = run("feedback", r, 4241, exe="ref_kat_baker")
    # ComputeProxyCounts (mode "counts"): light count, feedback available?, total feedback slots, use weight, sampling type, weights (some zero, some huge), usage counters
    c = np.zeros((1000, 64), np.float32)
    for i in range(len(c)):
        k = rng.integers(1, 17); c[i, 0] = k; c[i, 1] = rng.random() < 0.8; c[i, 3] = rng.choice(np.float32([0.0, 0.25, 0.75, 1.0])); c[i, 4] = 0 if rng.random() < 0.1 else 2
        w = (rng.gamma(0.7, 3.0, k) * (rng.random(k) < 0.85)).astype(np.float32); w[rng.random(k) < 0.05] *= np.float32(1e4)
        if w.sum() == 0: w[0] = 1
        c[i, 8:8 + k] = w; c[i, 5] = np.add.reduce(w, dtype=np.float32)
        use = rng.integers(0, 60, k + 1); c[i, 24:25 + k] = use; c[i, 2] = use.sum() if rng.random() < 0.9 else use[k]           # (the last case: no valid feedback at all)
    cout = run("counts", c, 40, exe="ref_kat_baker")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "baker_golden.npz"), baker_in=r, baker_out=out, counts_in=c, counts_out=cout,
                        source=np.array("Rtxpt/Lighting/LightsBaker.hlsl (ProcessFeedbackHistoryP0 / P1a / P1b / P2 / P3, ClearFeedbackHistory) at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_baker"))
    print(r.shape, out.shape, "nan:", int(np.isnan(out).sum()), os.path.getsize(os.path.join(ROOT, "tests", "golden", "baker_golden.npz")))
