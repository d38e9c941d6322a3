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
