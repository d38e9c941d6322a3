"""CPU: the environment-map baking path: the oracle's restatement of EnvMapBaker (oracle/pt_envbake.h) held to what the bake guarantees, and the product's
pass bodies (rtxpt_b200/csrc/envbake.cuh, compiled for the host by tests/emu) equal to the oracle.  GPU: tests/test_gpu_envbake.py (-m gpu)."""
import ctypes as C
import numpy as np
import pytest


def _bake(fn_lib, fn_name, cube_dim, source=None, scale=(1.0, 1.0, 1.0), lights=()):
    from rtxpt_b200 import lib as product_lib
    src, st, w, h, lt = product_lib.env_bake_arguments(cube_dim, source, None, scale, lights)
    total = sum(6 * (cube_dim >> m) ** 2 * 4 for m in range(cube_dim.bit_length()))
    out = np.zeros(total, np.float32); sc = np.float32(scale)
    f = getattr(fn_lib, fn_name); f.argtypes = [C.c_uint32] * 4 + [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    assert f(cube_dim, st, w, h, None if src is None else src.ctypes.data, sc.ctypes.data, len(lt), lt.ctypes.data if len(lt) else None, out.ctypes.data) == 0
    return product_lib.split_env_mips(out, cube_dim)


def _dirs(n):
    """Texel-centre directions of an n^2 cube, as CubemapGetDirectionFor lays the faces out."""
    u = (np.arange(n) + 0.5) / n; cx, cy = np.meshgrid(u * 2 - 1, 1 - u * 2)
    one = np.ones_like(cx)
    faces = [np.stack(v, -1) for v in ((one, cy, -cx), (-one, cy, cx), (cx, one, -cy), (cx, -one, cy), (cx, cy, one), (-cx, cy, -one))]
    d = np.stack(faces); return d / np.linalg.norm(d, axis=-1, keepdims=True)


def _solid_angles(n):
    e = np.arange(n + 1) * 2.0 / n - 1; X, Y = np.meshgrid(e, e)
    A = np.arctan2(X * Y, np.sqrt(X * X + Y * Y + 1))
    return np.abs(A[:-1, :-1] - A[1:, :-1] - A[:-1, 1:] + A[1:, 1:])


def test_oracle_bake_properties(oracle):
    L = oracle.lib()
    n = 64
    # a constant source stays constant through every MIP (weights normalise), scaled by ScaleColor, alpha 1
    src = np.zeros((32, 64, 4), np.float32); src[..., :3] = (0.5, 0.25, 1.0)
    mips = _bake(L, "oracle_bake_env_map", n, src, scale=(2.0, 1.0, 0.5))
    assert len(mips) == 7 and all(m.shape == (6, n >> i, n >> i, 4) for i, m in enumerate(mips))
    for m in mips: assert np.allclose(m[..., :3], (1.0, 0.25, 0.5), rtol=2e-3) and np.allclose(m[..., 3], 1.0)
    # orientation: an equirectangular image that encodes its own direction comes back as the cube's directions (world_to_latlong_map: centred on -z, y up)
    H, W = 256, 512
    v = (np.arange(H) + 0.5) / H; u = (np.arange(W) + 0.5) / W; U, V = np.meshgrid(u, v)
    theta = V * np.pi; phi = (U - 0.5) * 2 * np.pi
    d = np.stack([np.sin(theta) * np.sin(phi), np.cos(theta), -np.sin(theta) * np.cos(phi)], -1)
    src = np.zeros((H, W, 4), np.float32); src[..., :3] = d * 0.5 + 0.5
    mips = _bake(L, "oracle_bake_env_map", n, src)
    assert np.abs(mips[0][..., :3] * 2 - 1 - _dirs(n)).max() < 0.03
    # a directional light: energy = intensity x colour (W/sr x sr), centred on -direction, confined to its cone (+ the one-texel anti-aliasing fringe)
    n = 256
    sun_dir = np.float32([0.3, -0.8, 0.52]); sun_dir /= np.linalg.norm(sun_dir)
    ang = 0.2
    mips = _bake(L, "oracle_bake_env_map", n, None, lights=[((1.0, 0.5, 0.25), 7.0, sun_dir, ang)])
    sa = _solid_angles(n)[None, :, :, None]
    energy = (mips[0][..., :3] * sa).sum((0, 1, 2))
    assert np.allclose(energy, 7.0 * np.float32([1.0, 0.5, 0.25]), rtol=0.05), energy
    lit = mips[0][..., 0] > 0
    cosang = (_dirs(n) * (-sun_dir)).sum(-1)
    assert np.degrees(np.arccos(cosang[lit])).max() < np.degrees(ang / 2) + 1.0 and lit.sum() > 50
    # solid-angle weighted MIPs conserve the integral
    for m in range(1, 5):
        e = (mips[m][..., :3] * _solid_angles(n >> m)[None, :, :, None]).sum((0, 1, 2))
        assert np.allclose(e, energy, rtol=0.02), (m, e, energy)
    # fp16 storage and its clamp
    src = np.zeros((8, 16, 4), np.float32); src[..., :3] = 1e6
    m0 = _bake(L, "oracle_bake_env_map", 8, src)[0]
    assert (m0[..., :3] == 65504.0).all() and np.array_equal(m0, m0.astype(np.float16).astype(np.float32))


def test_product_bodies_equal_the_oracle(oracle):
    """Equirectangular and cube sources, lights, scale: the host build of envbake.cuh reproduces the oracle bit for bit (same libm, no contraction on either side)."""
    import host_build_lib as emu
    Lo, Le = oracle.lib(), emu.lib()
    rng = np.random.default_rng(4)
    eq = rng.gamma(2.0, 0.5, (48, 96, 4)).astype(np.float32)
    cube = rng.gamma(2.0, 0.5, (6, 16, 16, 4)).astype(np.float32)
    lights = [((1.0, 0.9, 0.7), 20.0, (0.2, -0.9, 0.4), 0.05), ((0.2, 0.3, 1.0), 3.0, (-0.7, -0.1, -0.7), 0.4)]
    lights = [(c, i, tuple(np.float32(d) / np.linalg.norm(d)), a) for c, i, d, a in lights]
    for n, src, sc, lt in ((32, eq, (1.0, 1.0, 1.0), ()), (64, eq, (0.5, 2.0, 1.5), lights), (32, cube, (1.0, 1.0, 1.0), lights[:1]), (16, None, (1.0, 1.0, 1.0), lights)):
        a = _bake(Lo, "oracle_bake_env_map", n, src, sc, lt); b = _bake(Le, "emu_bake_env_map", n, src, sc, lt)
        for m, (x, y) in enumerate(zip(a, b)): assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (n, m, np.abs(x - y).max())
