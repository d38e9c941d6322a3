"""Host: the JPEG decoder of the glTF loader (rtxpt_b200/csrc/jpeg.cpp) against Pillow's (libjpeg) decode of files Pillow wrote: baseline, grey, 4:4:4 / 4:2:2 / 4:2:0 chroma,
odd sizes, restart intervals; progressive files are refused.  glTF 2.0 allows PNG and JPEG images; the reference reads both through stb_image (Donut TextureCache)."""
import io
import numpy as np
import pytest
from rtxpt_b200 import lib as L


def _image(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([127 + 110 * np.sin(x / 7.0 + seed), 127 + 110 * np.cos(y / 5.0), 127 + 90 * np.sin((x + y) / 11.0)], -1) + rng.normal(0, 6, (h, w, 3))
    img[h // 3: h // 2, w // 4: w // 2] = (250, 20, 30)                      # a saturated block with hard edges
    return np.clip(img, 0, 255).astype(np.uint8)


def _jpeg(arr, **kw):
    from PIL import Image
    b = io.BytesIO(); Image.fromarray(arr).save(b, "JPEG", **kw); return b.getvalue()


def _pil_decode(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


@pytest.mark.parametrize("w,h,kw", [(64, 48, dict(quality=92, subsampling=0)), (67, 45, dict(quality=85, subsampling=1)), (70, 51, dict(quality=75, subsampling=2)),
                                    (33, 17, dict(quality=95, subsampling=2)), (128, 96, dict(quality=60, subsampling=2, optimize=True)), (8, 8, dict(quality=90, subsampling=0))])
def test_jpeg_matches_libjpeg(w, h, kw):
    data = _jpeg(_image(w, h, w + h), **kw)
    ours = L.decode_jpeg(data); ref = _pil_decode(data)
    assert ours.shape == (h, w, 4) and (ours[..., 3] == 255).all()
    d = np.abs(ours[..., :3].astype(np.int32) - ref.astype(np.int32))
    sub = kw.get("subsampling", 0)
    # 4:4:4: only IDCT / colour-conversion rounding differs; subsampled chroma: libjpeg's triangle filter vs bilinear at texel centres differ at hard chroma edges
    assert d.mean() < (0.15 if sub == 0 else 0.6), d.mean()                    # measured: 0.04 (4:4:4), 0.33 (4:2:2 / 4:2:0)
    assert np.percentile(d, 99) <= 2 and d.max() <= 6, (np.percentile(d, 99), d.max())


def test_jpeg_grey_and_restart_intervals():
    from PIL import Image
    g = _image(61, 40, 3)[..., 0]
    b = io.BytesIO(); Image.fromarray(g).save(b, "JPEG", quality=90); data = b.getvalue()
    ours = L.decode_jpeg(data); ref = np.asarray(Image.open(io.BytesIO(data)))
    assert np.abs(ours[..., 0].astype(int) - ref.astype(int)).max() <= 2 and (ours[..., 0] == ours[..., 1]).all() and (ours[..., 1] == ours[..., 2]).all()
    # restart markers every 2 MCUs (Pillow: restart_marker_blocks)
    data = _jpeg(_image(80, 64, 9), quality=88, subsampling=2, restart_marker_blocks=2)
    assert b"\xff\xdd" in data
    ours = L.decode_jpeg(data); ref = _pil_decode(data)
    d = np.abs(ours[..., :3].astype(int) - ref.astype(int)); assert d.mean() < 0.6 and np.percentile(d, 99) <= 2


def test_jpeg_refusals():
    prog = _jpeg(_image(32, 32, 1), quality=80, progressive=True)
    with pytest.raises(L.RtxptError, match="progressive"): L.decode_jpeg(prog)
    good = _jpeg(_image(32, 32, 1), quality=80)
    with pytest.raises(L.RtxptError, match="JPEG"): L.decode_jpeg(good[: len(good) // 3])          # truncated: a message, not a crash
    with pytest.raises(L.RtxptError, match="not a JPEG"): L.decode_jpeg(b"\x89PNG\r\n\x1a\n" + bytes(32))


def test_gltf_with_jpeg_texture_loads(product, tmp_path):
    """A glTF whose texture is a JPEG file loads (the exporter writes PNGs; the image is swapped for a JPEG of the same pixels)."""
    import json, os
    import gltf_export
    from test_gltf_loader import _textured_builder
    path = gltf_export.export(_textured_builder(), str(tmp_path / "s.gltf")); doc = json.load(open(path))
    from PIL import Image
    for im in doc["images"]:
        png = tmp_path / im["uri"]; rgb = np.asarray(Image.open(png).convert("RGB")); jpg = png.with_suffix(".jpg")
        Image.fromarray(rgb).save(jpg, "JPEG", quality=95, subsampling=0); im["uri"] = jpg.name; im.pop("mimeType", None); os.remove(png)
    json.dump(doc, open(path, "w"))
    gl = product.GltfScene(path); assert gl.desc.textureCount == len(doc["images"]) or gl.desc.textureCount >= 1
    t = gl.desc.textures[0]; assert t.width > 0 and t.mipLevels >= 1
    gl.close()
