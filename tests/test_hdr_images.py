"""Host: OpenEXR and Radiance .hdr readers (rtxpt_b200/csrc/hdr_images.cpp, rtxpt_b200_load_hdr_image) against files written by OpenCV's OpenEXR / RGBE writers
(tests/golden/make_hdr_image_golden.py) and against hand-built edge cases.  The reference reads these formats through tinyexr / stb_image in Donut's TextureCache
(External/Donut/src/engine/TextureCache.cpp:200-236) for its environment maps (Rtxpt/Sample.cpp:110-118)."""
import os, struct, zlib
import numpy as np
import pytest
from rtxpt_b200 import lib as L

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _file(name):
    with open(os.path.join(G, name), "rb") as f: return f.read()


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(G, "hdr_image_golden.npz"))


def test_exr_float_and_half_scanline_files(golden):
    rgb, rgba, gray = golden["rgb"], golden["rgba"], golden["gray"]
    a = L.load_hdr_image(_file("exr_zip_float.exr"))
    assert a.shape == rgb.shape[:2] + (4,) and np.array_equal(a[..., :3], rgb) and (a[..., 3] == 1).all()              # FLOAT channels: bit for bit
    a = L.load_hdr_image(_file("exr_zips_half.exr"))
    assert np.array_equal(a[..., :3], rgb.astype(np.float16).astype(np.float32))                                       # HALF channels: the half the writer rounded to
    a = L.load_hdr_image(_file("exr_rle_half_rgba.exr")); assert np.array_equal(a, rgba.astype(np.float16).astype(np.float32))
    a = L.load_hdr_image(_file("exr_none_float_rgba.exr")); assert np.array_equal(a, rgba)
    a = L.load_hdr_image(_file("exr_zip_gray.exr"))
    assert np.array_equal(a[..., 0], gray) and np.array_equal(a[..., 1], gray) and np.array_equal(a[..., 2], gray)     # a lone Y channel fills R, G and B


def test_exr_refusals_say_what_is_wrong():
    with pytest.raises(L.RtxptError, match="PIZ"): L.load_hdr_image(_file("exr_piz_half.exr"))
    good = bytearray(_file("exr_zip_float.exr"))
    with pytest.raises(L.RtxptError, match="truncated|offset|inflate"): L.load_hdr_image(bytes(good[:len(good) // 2]))
    bad = bytearray(good); bad[4:8] = struct.pack("<I", 2 | 0x200)
    with pytest.raises(L.RtxptError, match="tiled"): L.load_hdr_image(bytes(bad))
    bad = bytearray(good); bad[4:8] = struct.pack("<I", 2 | 0x1000)
    with pytest.raises(L.RtxptError, match="multi-part"): L.load_hdr_image(bytes(bad))
    with pytest.raises(L.RtxptError, match="signature"): L.load_hdr_image(b"not an image at all")
    bad = bytearray(good); i = bad.find(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4; bad[i + 8:i + 12] = struct.pack("<i", 1 << 30)
    with pytest.raises(L.RtxptError, match="dataWindow"): L.load_hdr_image(bytes(bad))


def _exr(channels, compression, width, height, rows, x_min=0, y_min=0, decreasing=False):
    """Minimal scan-line EXR writer for edge cases: channels = [(name, type)], rows[y] = bytes of one scan line in file layout (channel after channel)."""
    def attr(name, typ, data): return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<I", len(data)) + data
    chl = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", t, 0, 0, 0, 0, 1, 1) for n, t in channels) + b"\0"
    box = struct.pack("<iiii", x_min, y_min, x_min + width - 1, y_min + height - 1)
    hdr = struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chl) + attr("compression", "compression", bytes([compression])) + attr("dataWindow", "box2i", box) + \
        attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", bytes([1 if decreasing else 0])) + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + \
        attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    per = 16 if compression == 3 else 1; chunks = []
    for y0 in range(0, height, per):
        raw = b"".join(rows[y0:y0 + per])
        if compression in (2, 3):
            t = np.frombuffer(raw, np.uint8); half = (len(t) + 1) // 2; s = np.concatenate([t[0::2], t[1::2]]).astype(np.int32); assert len(s[:half]) == half
            d = s.copy(); d[1:] = (s[1:] - s[:-1] + 128 + 256) % 256
            z = zlib.compress(d.astype(np.uint8).tobytes()); data = z if len(z) < len(raw) else raw
        else: data = raw
        chunks.append((y_min + y0, data))
    if decreasing: chunks = chunks[::-1]
    table_at = len(hdr); at = table_at + 8 * len(chunks); offs = {}; body = b""
    for y, data in chunks: offs[y] = at + len(body); body += struct.pack("<iI", y, len(data)) + data
    table = b"".join(struct.pack("<Q", offs[y]) for y in sorted(offs))
    return hdr + table + body


def test_exr_edge_cases_written_by_hand():
    rng = np.random.default_rng(5); W, H = 19, 35
    A = rng.random((H, W)).astype(np.float16); B = rng.random((H, W)).astype(np.float32); R = (rng.random((H, W)) * 1000).astype(np.uint32); Z = rng.random((H, W)).astype(np.float32)
    # channels in file (alphabetical) order with mixed types, an unknown channel (Z) to skip, a data window that does not start at 0, chunks stored bottom-up
    ch = [("A", 1), ("B", 2), ("R", 0), ("Z", 2)]
    rows = [A[y].tobytes() + B[y].tobytes() + R[y].tobytes() + Z[y].tobytes() for y in range(H)]
    for comp in (0, 2, 3):
        img = L.load_hdr_image(_exr(ch, comp, W, H, rows, x_min=-7, y_min=11, decreasing=True))
        assert img.shape == (H, W, 4)
        assert np.array_equal(img[..., 3], A.astype(np.float32)) and np.array_equal(img[..., 2], B) and np.array_equal(img[..., 0], R.astype(np.float32)) and (img[..., 1] == 0).all()
    # a constant image: ZIP blocks far smaller than the data; and half special values
    h = np.zeros((3, 4), np.float16); h.view(np.uint16)[0] = (0x0001, 0x03FF, 0x7C00, 0xFC00); h.view(np.uint16)[1] = (0x8000, 0x3C00, 0x7BFF, 0x0400)
    img = L.load_hdr_image(_exr([("Y", 1)], 3, 4, 3, [h[y].tobytes() for y in range(3)]))
    assert np.array_equal(img[..., 0], h.astype(np.float32)) and np.array_equal(np.signbit(img[1, 0, 0]), True)
    with pytest.raises(L.RtxptError, match="no R, G, B or Y"): L.load_hdr_image(_exr([("Z", 2)], 0, 4, 2, [bytes(16)] * 2))


def test_radiance_hdr(golden):
    rgb = golden["rgb"]
    a = L.load_hdr_image(_file("radiance_rle.hdr"))
    assert a.shape == rgb.shape[:2] + (4,) and (a[..., 3] == 1).all()
    big = rgb.max(-1, keepdims=True)
    assert (np.abs(a[..., :3] - rgb) <= big / 128 + 1e-30).all()                       # 8-bit mantissa shared by the three components
    assert np.array_equal(a[..., :3], golden["radiance_opencv_decode"])                 # and exactly OpenCV's decode of the same file
    # flat (not run-length coded) scan lines, narrow image, zero exponent
    px = np.array([[[128, 64, 32, 129], [1, 2, 3, 0], [255, 0, 0, 136]]], np.uint8)
    f = b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 1 +X 3\n" + px.tobytes()
    out = L.load_hdr_image(f)
    assert np.array_equal(out[0, :, :3], np.array([[1.0, 0.5, 0.25], [0, 0, 0], [255.0, 0, 0]], np.float32))
    with pytest.raises(L.RtxptError, match="truncated"): L.load_hdr_image(f[:-3])
    with pytest.raises(L.RtxptError, match="resolution"): L.load_hdr_image(f.replace(b"-Y 1 +X 3", b"+X 3 -Y 1"))


def test_compare_script_reads_exr_against_pfm(tmp_path, golden):
    """scripts/compare_hdr_images.py: an EXR 'reference dump' against our PFM of the same pixels passes the 1e-3 gate, a perturbed one fails it."""
    import subprocess, sys
    from rtxpt_b200.imageio import write_pfm, read_pfm
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rgb = golden["rgb"]; p = str(tmp_path / "ours.pfm"); write_pfm(p, rgb); assert np.array_equal(read_pfm(p), rgb)
    cmd = [sys.executable, os.path.join(root, "scripts", "compare_hdr_images.py"), os.path.join(G, "exr_zip_float.exr")]
    ok = subprocess.run(cmd + [p], capture_output=True, text=True); assert ok.returncode == 0 and "per-pixel L2 0.000e+00" in ok.stdout, ok.stdout + ok.stderr
    write_pfm(p, rgb * 1.1 + 0.05); bad = subprocess.run(cmd + [p], capture_output=True, text=True); assert bad.returncode == 1, bad.stdout


def test_hostile_image_sizes_are_refused_before_allocation():
    """Headers that claim gigapixel images (a 17 GB allocation in RGBA32F) are refused by their size, in all three decoders."""
    import io
    from PIL import Image
    good = bytearray(_file("exr_zip_float.exr")); i = good.find(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4
    good[i:i + 16] = struct.pack("<iiii", 0, 0, 32767, 32767)
    with pytest.raises(L.RtxptError, match="2\\^28|dataWindow"): L.load_hdr_image(bytes(good))
    with pytest.raises(L.RtxptError, match="2\\^28|dimensions"): L.load_hdr_image(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y 32768 +X 32768\n" + bytes(64))
    b = io.BytesIO(); Image.new("RGB", (16, 16)).save(b, "JPEG"); j = bytearray(b.getvalue()); k = j.find(b"\xff\xc0"); j[k + 5:k + 9] = struct.pack(">HH", 32768, 32768)
    with pytest.raises(L.RtxptError, match="2\\^28|dimensions"): L.decode_jpeg(bytes(j))
