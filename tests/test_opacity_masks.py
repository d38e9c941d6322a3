"""CPU: the opacity-mask baker (rtxpt_b200/csrc/opacity_masks.{h,cpp}; the equivalent of the reference's OMM bake, Rtxpt/OpacityMicroMap/) through its host-only C-ABI hooks.
The property that makes the masks safe: a micro-triangle is marked opaque / transparent only when the alpha test gives that answer at EVERY point of it, so the traversal may use the
mask instead of the texture fetch without changing a single hit.  Checked here against a float bilinear mip-0 sampler (wrap addressing) at thousands of points per triangle."""
import ctypes as C
import numpy as np
import pytest

TRANSPARENT, OPAQUE, UNKNOWN = 0, 1, 2


def _lib(product):
    L = C.CDLL(product.LIB_PATH)
    L.rtxpt_b200_host_bake_opacity_mask.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]; L.rtxpt_b200_host_bake_opacity_mask.restype = C.c_int
    L.rtxpt_b200_host_opacity_micro_index.argtypes = [C.c_float, C.c_float]; L.rtxpt_b200_host_opacity_micro_index.restype = C.c_uint32
    return L


def _bilinear_alpha(img, uv):
    """mip 0, wrap, texel centres at integer + 0.5: what tex2DLod(..., 0).w returns up to the filter's fixed-point weights."""
    h, w = img.shape[:2]
    x = uv[:, 0].astype(np.float64) * w - 0.5; y = uv[:, 1].astype(np.float64) * h - 0.5
    x0 = np.floor(x).astype(np.int64); y0 = np.floor(y).astype(np.int64); fx = x - x0; fy = y - y0
    a = img[..., 3].astype(np.float64) / (255.0 if img.dtype == np.uint8 else 1.0)
    s = lambda xx, yy: a[yy % h, xx % w]
    return (s(x0, y0) * (1 - fx) + s(x0 + 1, y0) * fx) * (1 - fy) + (s(x0, y0 + 1) * (1 - fx) + s(x0 + 1, y0 + 1) * fx) * fy


def _states(mask):
    m = np.asarray(mask, np.uint32)
    return np.array([(int(m[i >> 4]) >> ((i & 15) * 2)) & 3 for i in range(64)])


def test_micro_index_covers_the_triangle_exactly_once(product):
    L = _lib(product)
    rng = np.random.default_rng(3)
    uv = rng.random((200000, 2), dtype=np.float32); uv = uv[uv.sum(1) <= 1.0]
    idx = np.array([L.rtxpt_b200_host_opacity_micro_index(float(u), float(v)) for u, v in uv[:40000]])
    assert idx.min() == 0 and idx.max() == 63 and len(np.unique(idx)) == 64
    counts = np.bincount(idx, minlength=64) / len(idx)
    assert np.abs(counts - 1 / 64).max() < 0.004                               # 64 micro-triangles of equal area
    for u, v in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (0.5, 0.5), (0.999999, 0.0), (1e-9, 0.9999999), (0.125, 0.125), (1.0000001, 0.0)):
        assert L.rtxpt_b200_host_opacity_micro_index(u, v) < 64              # corners, the hypotenuse, lattice points and an ulp outside


@pytest.mark.parametrize("kind", ["binary_leaves", "smooth", "float32"])
def test_known_states_agree_with_the_alpha_test_everywhere(product, kind):
    L = _lib(product)
    rng = np.random.default_rng({"binary_leaves": 1, "smooth": 2, "float32": 3}[kind])
    S = 64
    n = rng.random((S, S)); n = (n + np.roll(n, 1, 0) + np.roll(n, 1, 1) + np.roll(n, 2, 0) + np.roll(n, 3, 1)) / 5
    for _ in range(3): n = (n + np.roll(n, 1, 0) + np.roll(n, -1, 0) + np.roll(n, 1, 1) + np.roll(n, -1, 1)) / 5
    n = (n - n.min()) / (n.max() - n.min())
    if kind == "binary_leaves": alpha = np.where(n > 0.5, 255, 0).astype(np.uint8)
    else: alpha = np.clip(n * 255, 0, 255).astype(np.uint8)
    if kind == "float32":
        img = np.zeros((S, S, 4), np.float32); img[..., 3] = alpha / 255.0; fmt = 2
    else:
        img = np.zeros((S, S, 4), np.uint8); img[..., 3] = alpha; fmt = 0
    img = np.ascontiguousarray(img)
    seen = np.zeros(3, np.int64); seen_small = np.zeros(3, np.int64)
    for t in range(60):
        scale = [0.05, 0.2, 1.0, 3.0][t % 4]
        uv = ((rng.random((3, 2)) - 0.5) * scale + rng.random(2) * 4 - 2).astype(np.float32)                  # small to tiling triangles, also outside [0, 1): wrap
        cutoff = int(rng.integers(1, 255)) if kind != "binary_leaves" else 128
        mask = np.zeros(4, np.uint32)
        assert L.rtxpt_b200_host_bake_opacity_mask(img.ctypes.data, S, S, fmt, cutoff, uv.ctypes.data, mask.ctypes.data) == 0
        st = _states(mask); seen += np.bincount(st, minlength=3)[:3]
        if t % 4 == 0: seen_small += np.bincount(st, minlength=3)[:3]
        b = rng.random((6000, 2), dtype=np.float32); b = b[b.sum(1) <= 1.0]
        pts = (uv[0] * (1 - b[:, :1] - b[:, 1:2]) + uv[1] * b[:, :1] + uv[2] * b[:, 1:2]).astype(np.float32)
        passes = _bilinear_alpha(img, pts) >= cutoff / 255.0
        idx = np.array([L.rtxpt_b200_host_opacity_micro_index(float(u), float(v)) for u, v in b])
        s = st[idx]
        assert passes[s == OPAQUE].all(), (kind, t, "opaque micro-triangle with a failing point")
        assert (~passes[s == TRANSPARENT]).all(), (kind, t, "transparent micro-triangle with a passing point")
    assert seen[OPAQUE] > 0 and seen[TRANSPARENT] > 0 and seen[UNKNOWN] > 0                                    # the fixture exercises all three states
    if kind == "binary_leaves": assert seen_small[UNKNOWN] < 0.7 * seen_small.sum(), seen_small                 # micro-triangles a few texels across over a leaf texture: a good share is decided


def test_degenerate_inputs(product):
    L = _lib(product)
    img = np.full((8, 8, 4), 200, np.uint8); mask = np.zeros(4, np.uint32)
    uv = np.zeros((3, 2), np.float32)                                            # geometry without texture coordinates: every micro-triangle samples texel (0, 0)
    assert L.rtxpt_b200_host_bake_opacity_mask(img.ctypes.data, 8, 8, 0, 128, uv.ctypes.data, mask.ctypes.data) == 0 and (_states(mask) == OPAQUE).all()
    assert L.rtxpt_b200_host_bake_opacity_mask(img.ctypes.data, 8, 8, 0, 200, uv.ctypes.data, mask.ctypes.data) == 0 and (_states(mask) == UNKNOWN).all()     # at the cutoff: never trusted
    uv[:] = np.nan
    assert L.rtxpt_b200_host_bake_opacity_mask(img.ctypes.data, 8, 8, 0, 128, uv.ctypes.data, mask.ctypes.data) == 0 and (_states(mask) == UNKNOWN).all()
    uv = (np.array([[0, 0], [500, 0], [0, 500]], np.float32))                    # a footprint of thousands of texels per micro-triangle: left to the texture test
    assert L.rtxpt_b200_host_bake_opacity_mask(img.ctypes.data, 8, 8, 0, 128, uv.ctypes.data, mask.ctypes.data) == 0 and (_states(mask) == UNKNOWN).all()
    assert L.rtxpt_b200_host_bake_opacity_mask(None, 8, 8, 0, 128, uv.ctypes.data, mask.ctypes.data) != 0
