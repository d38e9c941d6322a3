"""Random StandardBSDF test records (36 floats: V N T B wo u[3] diffuse rough specular metallic transmission diffTrans specTrans eta thin lobes pad pad)."""
import numpy as np


def _f16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


def orthonormal_frames(rng, n):
    N = rng.normal(size=(n, 3)); N /= np.linalg.norm(N, axis=1, keepdims=True)
    a = rng.normal(size=(n, 3)); T = a - N * (a * N).sum(1, keepdims=True); T /= np.linalg.norm(T, axis=1, keepdims=True)
    B = np.cross(N, T)
    return N.astype(np.float32), T.astype(np.float32), B.astype(np.float32)


def sphere_dirs(rng, n):
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d.astype(np.float32)


def make_records(rng, n, kind="mixed", wo=None):
    N, T, B = orthonormal_frames(rng, n)
    # view direction in the upper hemisphere of the shading frame (the path tracer guarantees this through adjustShadingNormal)
    l = sphere_dirs(rng, n); l[:, 2] = np.abs(l[:, 2]) * 0.98 + 0.02; l /= np.linalg.norm(l, axis=1, keepdims=True)
    V = (T * l[:, 0:1] + B * l[:, 1:2] + N * l[:, 2:3]).astype(np.float32)
    if wo is None:
        wo = sphere_dirs(rng, n)
    rec = np.zeros((n, 36), np.float32)
    rec[:, 0:3], rec[:, 3:6], rec[:, 6:9], rec[:, 9:12], rec[:, 12:15] = V, N, T, B, wo
    rec[:, 15:18] = rng.random((n, 3), dtype=np.float32)
    base = rng.random((n, 3), dtype=np.float32)
    metal = (rng.random(n) < 0.25).astype(np.float32) * rng.random(n, dtype=np.float32)
    rough = rng.random(n, dtype=np.float32) ** 1.5
    rough[rng.random(n) < 0.1] = 0.02                  # delta lobes (alpha < kMinGGXAlpha)
    trans = np.zeros(n, np.float32); dtrans = np.zeros(n, np.float32); thin = np.ones(n, np.float32)
    if kind in ("mixed", "transmissive"):
        tm = rng.random(n) < (0.35 if kind == "mixed" else 1.0)
        trans[tm] = rng.random(tm.sum(), dtype=np.float32) * 0.9 + 0.1
        dm = rng.random(n) < 0.1
        dtrans[dm] = rng.random(dm.sum(), dtype=np.float32)
        thin = (rng.random(n) < 0.5).astype(np.float32)
        thin[(trans == 0) & (dtrans == 0)] = 1.0
    ior = np.where(rng.random(n) < 0.8, 1.5, 1.0 + rng.random(n) * 1.2).astype(np.float32)
    f0 = ((ior - 1) / (ior + 1)) ** 2
    front = rng.random(n) < 0.7
    eta = np.where((thin == 0) & ~front, ior, 1.0 / ior)
    rec[:, 18:21] = _f16(base * (1 - metal[:, None]))
    rec[:, 21] = _f16(rough)
    rec[:, 22:25] = _f16(f0[:, None] * (1 - metal[:, None]) + base * metal[:, None])
    rec[:, 25] = _f16(metal)
    rec[:, 26:29] = _f16(base)
    rec[:, 29] = _f16(dtrans * (1 - metal)); rec[:, 30] = _f16(trans * (1 - metal)); rec[:, 31] = _f16(eta)
    rec[:, 32] = thin; rec[:, 33] = 255.0
    return rec
