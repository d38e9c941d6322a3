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
