"""CPU: tone mapping / auto exposure: the oracle's restatement of RTXPT's ToneMapper (oracle/pt_tonemap.h) held to what the operators and the colour pipeline
guarantee, and the product's host constants + pixel bodies (rtxpt_b200/csrc/tonemap.cuh, host build in tests/emu) equal to the oracle.  GPU: tests/test_gpu_tonemap.py."""
import ctypes as C
import numpy as np
import pytest
from rtxpt_b200 import structs as S


def _run(L, fn, params, img):
    rgba = np.ascontiguousarray(img, np.float32).reshape(-1, 4); out = np.zeros((len(rgba), 4), np.uint8); aux = np.zeros(4, np.float32)
    f = getattr(L, fn); f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    assert f(C.byref(params), rgba.ctypes.data, len(rgba), out.ctypes.data, aux.ctypes.data) == 0
    return out.reshape(img.shape[:-1] + (4,)), aux


def _srgb_decode(u8): v = u8.astype(np.float64) / 255.0; return np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)


def test_oracle_tone_mapping_properties(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    img = np.concatenate([rng.gamma(1.5, 0.6, (40, 60, 3)), rng.random((40, 60, 1))], -1).astype(np.float32)
    # disabled: a pure sRGB encode of the clamped input; alpha is stored linearly
    out, aux = _run(L, "oracle_tone_map", S.make_tone_mapping_params(enabled=False), img)
    assert np.abs(_srgb_decode(out[..., :3]) - np.clip(img[..., :3], 0, 1)).max() < 0.012 and np.array_equal(out[..., 3], (np.clip(img[..., 3], 0, 1) * 255 + 0.5).astype(np.uint8))
    # Linear operator + exposure compensation of one stop = the input doubled
    out2, _ = _run(L, "oracle_tone_map", S.make_tone_mapping_params(op=0, exposure_compensation=1.0), img)
    assert np.abs(_srgb_decode(out2[..., :3]) - np.clip(img[..., :3] * 2, 0, 1)).max() < 0.012
    # ACES: monotone per channel, maps 0 to 0, saturates below 1
    ramp = np.zeros((1, 256, 4), np.float32); ramp[0, :, :3] = np.linspace(0, 16, 256)[:, None]; ramp[..., 3] = 1
    a, _ = _run(L, "oracle_tone_map", S.make_tone_mapping_params(op=5), ramp)
    assert (np.diff(a[0, :, 0].astype(int)) >= 0).all() and a[0, 0, 0] == 0 and a[0, -1, 0] == 255
    # every operator gives finite, in-range output and keeps black black
    for op in range(6):
        o, _ = _run(L, "oracle_tone_map", S.make_tone_mapping_params(op=op), ramp)
        assert o[0, 0, :3].max() <= 1 and (np.diff(o[0, :, 1].astype(int)) >= -1).all()
    # auto exposure: the average luminance is the geometric mean, and scaling the image leaves the mapped result unchanged (within the exposure bounds)
    p = S.make_tone_mapping_params(op=1, auto_exposure=True)
    o1, aux1 = _run(L, "oracle_tone_map", p, img); img4 = img.copy(); img4[..., :3] *= 4
    o4, aux4 = _run(L, "oracle_tone_map", p, img4)
    lum = img[..., :3] @ np.float32([0.299, 0.587, 0.114])
    assert np.isclose(aux1[0], np.exp2(np.log2(np.maximum(lum, 1e-4)).mean()), rtol=1e-4) and np.isclose(aux4[0], 4 * aux1[0], rtol=1e-4)
    assert np.abs(o1.astype(int) - o4.astype(int)).max() <= 1
    # white balance: D65 is the identity; a warm white point pushes a grey towards blue (the transform undoes the illuminant), pre-exposed grey inverts the transform
    grey = np.full((4, 4, 4), 0.05, np.float32)
    d65, _ = _run(L, "oracle_tone_map", S.make_tone_mapping_params(op=0, white_balance=True, white_point=6500.0), grey)
    warm, auxw = _run(L, "oracle_tone_map", S.make_tone_mapping_params(op=0, white_balance=True, white_point=3200.0), grey)
    assert np.array_equal(d65[..., 0], d65[..., 2]) and abs(int(d65[0, 0, 0]) - int(d65[0, 0, 1])) <= 1
    lin = _srgb_decode(warm[0, 0, :3]); assert lin[2] > 1.3 * lin[0] and lin.max() < 1.0, lin
    _, aux0 = _run(L, "oracle_tone_map", S.make_tone_mapping_params(op=0, exposure_compensation=2.0), grey)
    assert np.allclose(aux0[1:], 0.18 / 4, rtol=1e-5)


def test_product_bodies_equal_the_oracle(oracle):
    import host_build_lib as emu
    Lo, Le = oracle.lib(), emu.lib()
    rng = np.random.default_rng(6)
    img = np.concatenate([rng.gamma(1.2, 0.9, (50, 70, 3)), rng.random((50, 70, 1))], -1).astype(np.float32); img[0, 0, :3] = 0; img[1, 1, :3] = 1e4
    for kw in (dict(), dict(op=0), dict(op=1, auto_exposure=True), dict(op=2, white_max_luminance=3.0, exposure_compensation=-1.5), dict(op=3), dict(op=4, white_scale=6.0, clamped=False),
               dict(op=5, white_balance=True, white_point=4200.0, film_speed=400.0, f_number=2.8, shutter=60.0), dict(enabled=False, auto_exposure=True, exposure_value_min=-2.0, exposure_value_max=1.0)):
        p = S.make_tone_mapping_params(**kw)
        a, auxa = _run(Lo, "oracle_tone_map", p, img); b, auxb = _run(Le, "emu_tone_map", p, img)
        assert np.array_equal(a, b), (kw, int((a != b).sum())); assert np.array_equal(auxa, auxb), (kw, auxa, auxb)
