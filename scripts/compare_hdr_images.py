"""Compares two HDR images pixel by pixel: BASELINE.md's parity metric (per-pixel L2 of linear RGB, gate 1e-3 at 1024 spp) between, typically, an `AccumulatedRadiance` dump of an RTXPT
reference-mode run made on another machine (.exr; .hdr and HDR .dds read too) and this framework's accumulation (.pfm from examples/render_gltf, or any of the former).
    python scripts/compare_hdr_images.py rtxpt_dump.exr ours.pfm [--crop x0 y0 x1 y1]
Host only (the readers are rtxpt_b200_load_hdr_image in librtxpt_b200.so; no GPU is touched)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def read(path):
    from rtxpt_b200 import lib as L
    from rtxpt_b200.imageio import read_pfm
    if path.lower().endswith(".pfm"):
        a = read_pfm(path); return a if a.ndim == 3 else np.repeat(a[..., None], 3, -1)
    with open(path, "rb") as f: a = L.load_hdr_image(f.read())
    if a.ndim == 4: raise SystemExit("%s is a cube map: compare faces one at a time" % path)
    return a[..., :3]


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("reference"); ap.add_argument("ours"); ap.add_argument("--crop", type=int, nargs=4)
    args = ap.parse_args()
    from rtxpt_b200.imageio import per_pixel_l2
    a, b = read(args.reference)[..., :3], read(args.ours)[..., :3]
    if a.shape != b.shape: raise SystemExit("sizes differ: %s vs %s" % (a.shape, b.shape))
    if args.crop: x0, y0, x1, y1 = args.crop; a, b = a[y0:y1, x0:x1], b[y0:y1, x0:x1]
    rel = np.abs(a - b) / (np.abs(a) + 1e-2)
    print("size %dx%d  mean %.6f vs %.6f  per-pixel L2 %.3e (gate 1e-3)  pixels within 5 %%: %.4f  max abs diff %.4g" %
          (a.shape[1], a.shape[0], a.mean(), b.mean(), per_pixel_l2(a, b), (rel.max(-1) < 0.05).mean(), np.abs(a - b).max()))
    return 0 if per_pixel_l2(a, b) <= 1e-3 else 1


if __name__ == "__main__":
    sys.exit(main())
