"""Writes the small EXR / Radiance fixtures of tests/test_hdr_images.py with OpenCV's writers (OpenEXR library and OpenCV's RGBE coder: implementations independent of
rtxpt_b200/csrc/hdr_images.cpp) and stores the pixels that went in.  Run here (OpenCV is in this image), commit the outputs:
    OPENCV_IO_ENABLE_OPENEXR=1 python tests/golden/make_hdr_image_golden.py"""
import os
os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import numpy as np, cv2

here = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(77)
W, H = 37, 29                                                           # not a multiple of ZIP's 16-line blocks
img = (rng.gamma(1.5, 0.7, (H, W, 3))).astype(np.float32); img[3:9, 5:20] = 0.25; img[0, 0] = (0, 0, 0); img[1, 1] = (1e-6, 700.0, 65504.0)
rgba = np.concatenate([img, rng.random((H, W, 1), np.float32)], -1)
gray = img[..., 0].copy()
out = {"rgb": img, "rgba": rgba, "gray": gray}
T, C = cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_COMPRESSION
cases = {"exr_zip_float.exr": (img, [T, cv2.IMWRITE_EXR_TYPE_FLOAT, C, cv2.IMWRITE_EXR_COMPRESSION_ZIP]),
         "exr_zips_half.exr": (img, [T, cv2.IMWRITE_EXR_TYPE_HALF, C, cv2.IMWRITE_EXR_COMPRESSION_ZIPS]),
         "exr_rle_half_rgba.exr": (rgba, [T, cv2.IMWRITE_EXR_TYPE_HALF, C, cv2.IMWRITE_EXR_COMPRESSION_RLE]),
         "exr_none_float_rgba.exr": (rgba, [T, cv2.IMWRITE_EXR_TYPE_FLOAT, C, cv2.IMWRITE_EXR_COMPRESSION_NO]),
         "exr_zip_gray.exr": (gray, [T, cv2.IMWRITE_EXR_TYPE_FLOAT, C, cv2.IMWRITE_EXR_COMPRESSION_ZIP]),
         "exr_piz_half.exr": (img, [T, cv2.IMWRITE_EXR_TYPE_HALF, C, cv2.IMWRITE_EXR_COMPRESSION_PIZ])}        # refused, with a message
for name, (a, flags) in cases.items():
    src = a[..., [2, 1, 0] + ([3] if a.shape[-1] == 4 else [])] if a.ndim == 3 else a                         # OpenCV's channel order is BGR(A)
    assert cv2.imwrite(os.path.join(here, name), np.ascontiguousarray(src), flags), name
assert cv2.imwrite(os.path.join(here, "radiance_rle.hdr"), np.ascontiguousarray(img[..., ::-1]))
back = cv2.imread(os.path.join(here, "radiance_rle.hdr"), cv2.IMREAD_UNCHANGED)[..., ::-1]                   # OpenCV's own RGBE decode, for the record (it adds no half-step bias either)
out["radiance_opencv_decode"] = np.ascontiguousarray(back)
np.savez_compressed(os.path.join(here, "hdr_image_golden.npz"), **out)
for n in sorted(os.listdir(here)):
    if n.endswith((".exr", ".hdr")): print(n, os.path.getsize(os.path.join(here, n)))
