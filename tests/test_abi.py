"""CPU: the C-ABI shared library loads and exports every entry point include/rtxpt_b200.h declares; without a device it fails loudly."""
import ctypes as C
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rtxpt_b200.h")).read()
    return sorted(set(re.findall(r"RTXPT_API\s+[\w\s\*]+?\b(rtxpt_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(product):
    names = declared_symbols()
    assert len(names) >= 20
    for path in (product.LIB_PATH, product.LIB_PATH_STRICT):
        L = C.CDLL(path)
        for n in names:
            assert hasattr(L, n), (path, n)
    assert set(product.EXPORTED_SYMBOLS) == set(names)


def test_struct_layouts_match_reference_tables():
    from rtxpt_b200 import structs as S
    assert C.sizeof(S.GeometryData) == 64       # c_SizeOfGeometryData, bindless.h:83
    assert C.sizeof(S.InstanceData) == 112      # c_SizeOfInstanceData, bindless.h:84
    assert C.sizeof(S.SubInstanceData) == 32
    assert C.sizeof(S.MaterialData) == 128
    assert C.sizeof(S.CameraData) == 112
    assert S.MaterialData.IoR.offset == 100 and S.MaterialData.VolumeAttenuationColor.offset == 112


def test_no_cpu_fallback_without_device(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(product.RtxptError, match="no CUDA device"):
        product.Context()


def test_product_does_not_reference_oracle():
    # the product (rtxpt_b200/) must never import, link or execute anything under oracle/
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rtxpt_b200")):
        if "_build" in dirpath or "__pycache__" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("the oracle's", ""), os.path.join(dirpath, f)


def test_camera_matrices_helper_matches_the_python_mirror():
    """rtxpt_b200_camera_matrices (what a C++ host without Donut's PlanarView uses for set_view / the realtime, denoiser and ReBLUR constants) against scene_builder's functions, which
    the parity tests feed to the oracle and the product alike: equal to float rounding (the two differ in where they round to float, not in the formula)."""
    import ctypes as C
    import numpy as np
    from rtxpt_b200 import lib, scene_builder as sb
    f = lib.load().rtxpt_b200_camera_matrices; f.argtypes = [C.c_void_p] * 4; f.restype = C.c_int
    for cam in (sb.bridge_camera(640, 360, pos=(1, 2, -3), direction=(0.2, -0.1, 1), up=(0, 1, 0), fov_y=0.7),
                sb.bridge_camera(1920, 1080, pos=(-40.5, 7.25, 13), direction=(-0.7, 0.3, -0.4), up=(0, 1, 0), fov_y=1.1, near_z=0.05, far_z=5000.0)):
        wv, vc, wc = (np.zeros(16, np.float32) for _ in range(3))
        assert f(C.byref(cam), wv.ctypes.data, vc.ctypes.data, wc.ctypes.data) == 0
        for got, want in ((wv, sb.world_to_view(cam)), (vc, sb.view_to_clip(cam)), (wc, sb.world_to_clip(cam))):
            assert np.allclose(got.reshape(4, 4), want, rtol=2e-6, atol=2e-6)
        assert f(C.byref(cam), None, None, wc.ctypes.data) == 0
