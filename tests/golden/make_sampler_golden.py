"""Generates tests/golden/sampler_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/Lighting/LightSampler.hlsli (with LightingTypes.hlsli's LightFeedbackReservoir and
LightingAlgorithms.hlsli's LocalLightBinarySearch) compiled in place as C++ through oracle/ref_hlsl_shim.h (oracle/_ref/ref_kat_bsdf, mode "sampler").  Run in the build
container only:
    make -C oracle ref && python tests/golden/make_sampler_golden.py
  sampler_in [M,680]: header (proxy count, tile jitter x / y, candidate samples, full samples, local-to-global ratio, screen-space threshold, -), 16 proxy counters, 64 proxy
      indices, 2 x 2 tiles x 128 packed (light << 9 | count - 1) words as bit patterns, 8 queries x (pixel x, y, rnd, light, flags [1 screen-space coherent, 2 from the local
      sampler, 4 sampleable by the BSDF], contribution, feedback rnd, bsdf pdf, solid-angle pdf, selection pdf)
  sampler_out [M,128]: per query: SampleGlobal (light, pdf), SampleLocal (light, pdf), SampleGlobalPDF, SampleLocalPDF, local / global candidate counts, MIS for a BSDF hit,
      other sampler's pdf, this / other count, MIS for the light sample, feedback reservoir (total weight, candidate bits) after InsertFeedbackFromNEE, the coherence heuristic"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_bsdf_golden import run  # noqa: E402


def pack_tile(lights):
    lights = np.sort(np.asarray(lights, np.uint32)); _, inv, cnt = np.unique(lights, return_inverse=True, return_counts=True)
    return (lights << np.uint32(9)) | (cnt[inv].astype(np.uint32) - np.uint32(1))


def make(rng, n, quirk=False):
    rec = np.zeros((n, 680), np.float32)
    for i in range(n):
        r = rec[i]
        counters = rng.integers(0, 9, 16) * (rng.random(16) < 0.7)
        if counters.sum() == 0: counters[rng.integers(0, 16)] = 3
        while counters.sum() > 64: counters[np.argmax(counters)] -= 1
        idx = np.repeat(np.arange(16), counters)
        r[0] = len(idx); r[1], r[2] = rng.integers(0, 8, 2); r[3] = rng.integers(1, 9); r[4] = rng.integers(1, 3)
        r[5] = rng.choice(np.float32([0.0, 0.35, 0.65, 1.0, rng.random()])); r[6] = 0.3
        r[8:24] = counters; r[24:24 + len(idx)] = idx
        tiles = np.zeros((4, 128), np.uint32)
        for t in range(4):
            pool = rng.choice(16, rng.integers(1, 9), replace=False); tiles[t] = pack_tile(rng.choice(pool, 128))
        q = r[600:].reshape(8, 10)
        q[:, 0:2] = rng.integers(0, 8, (8, 2)); q[:, 2] = rng.random(8, dtype=np.float32); q[:, 3] = rng.integers(0, 16, 8); q[:, 4] = rng.integers(0, 8, 8)
        q[:, 5] = np.exp(rng.uniform(-8, 6, 8)); q[:, 6] = rng.random(8, dtype=np.float32)
        q[:, 7] = np.float16(np.exp(rng.uniform(-4, 6, 8))) * (rng.random(8) < 0.9); q[:, 8] = np.exp(rng.uniform(-6, 8, 8)); q[:, 9] = rng.integers(1, 129, 8) / np.float32(128)
        q[0, 2] = 0.0; q[1, 2] = np.nextafter(np.float32(1), np.float32(0)); q[2, 0:2] = q[3, 0:2]             # rnd at both ends; two inserts into one reservoir
        if i % 9 == 0 and counters[int(q[4, 3])] > 0: q[4, 5] = 0.0      # a zero contribution (not for a light without global proxies: 0 / pow(0, .65) is a NaN, and min( 1e12, NaN ) is 1e12 in DXIL but NaN in the shim)
        if quirk:
            # LocalLightBinarySearch runs its 8 steps without an empty-range test (LightingAlgorithms.hlsli:655-685): a light below every key of the tile makes step 8 read the
            # word just before the tile - the previous tile's last entry, or (tile 0) an out-of-range address that reads as 0, i.e. "light 0, count 1"
            tiles[0] = pack_tile(rng.choice(np.arange(1, 6), 128)); tiles[0][-1] = (np.uint32(5) << 9) | (tiles[0][-1] & np.uint32(0x1FF))
            tiles[1] = pack_tile(rng.choice(np.arange(8, 14), 128))
            r[1] = 0; r[2] = 0
            q[5, 0:2] = (4, 3); q[5, 3] = 0; q[5, 4] = 1             # tile 0, light 0 (absent, below all keys)
            q[6, 0:2] = (4, 3); q[6, 3] = 1 + i % 5                  # tile 0, a light that is there
            q[7, 0:2] = (7, 2); q[7, 3] = 5; q[7, 4] = 1; r[1] = 1   # jitter 1: pixel 7 lands in tile 1; light 5 = tile 0's last key, below all of tile 1's
        r[88:600] = tiles.reshape(-1).view(np.float32)
    return rec


if __name__ == "__main__":
    rng = np.random.default_rng(4242)
    u = np.concatenate([make(rng, 1100), make(rng, 100, quirk=True)])
    out = run("sampler", u, 128)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler_golden.npz"), sampler_in=u, sampler_out=out,
                        source=np.array("Rtxpt/Shaders/PathTracer/Lighting/LightSampler.hlsli, LightingTypes.hlsli, LightingAlgorithms.hlsli at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    print(u.shape, out.shape, os.path.getsize(os.path.join(ROOT, "tests", "golden", "sampler_golden.npz")))
