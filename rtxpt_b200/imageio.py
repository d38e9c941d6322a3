"""Tiny image helpers for tests/bench (no third-party imaging library in the image): tonemapped PNG writer, error metrics."""
import struct
import zlib
import numpy as np


def write_png(path, rgb_linear, exposure=1.0):
    img = np.clip(np.asarray(rgb_linear, np.float32)[..., :3] * exposure, 0, None)
    img = img / (1.0 + img)                                   # Reinhard
    img = np.where(img <= 0.0031308, img * 12.92, 1.055 * np.power(np.maximum(img, 1e-8), 1 / 2.4) - 0.055)
    u8 = np.clip(np.rint(img * 255), 0, 255).astype(np.uint8)
    h, w = u8.shape[:2]
    raw = b"".join(b"\x00" + u8[y].tobytes() for y in range(h))
    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def per_pixel_l2(a, b):
    """Mean over pixels of the squared RGB distance — the per-pixel L2 metric BASELINE.json's tolerance gate is quoted in."""
    d = np.asarray(a, np.float64)[..., :3] - np.asarray(b, np.float64)[..., :3]
    return float((d * d).sum(-1).mean())


def rel_mse(a, b, eps=1e-2):
    a = np.asarray(a, np.float64)[..., :3]; b = np.asarray(b, np.float64)[..., :3]
    return float((((a - b) ** 2) / (b * b + eps)).mean())


def read_pfm(path):
    """PFM ('PF' RGB or 'Pf' grey, scale < 0 = little endian, bottom row first - what examples/render_gltf.cpp writes) -> H x W x 3 (or H x W) float32, top row first."""
    with open(path, "rb") as f:
        kind = f.readline().strip(); w, h = (int(v) for v in f.readline().split()); scale = float(f.readline())
        if kind not in (b"PF", b"Pf"): raise ValueError("%s: not a PFM file" % path)
        ch = 3 if kind == b"PF" else 1
        a = np.frombuffer(f.read(w * h * ch * 4), "<f4" if scale < 0 else ">f4").astype(np.float32)
    a = a.reshape(h, w, ch)[::-1]
    return np.ascontiguousarray(a if ch == 3 else a[..., 0])


def write_pfm(path, rgb):
    a = np.asarray(rgb, np.float32)[..., :3]
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (a.shape[1], a.shape[0])); f.write(np.ascontiguousarray(a[::-1]).astype("<f4").tobytes())
