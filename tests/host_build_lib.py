"""ctypes binding of tests/emu/_build/libreblur_emu.so - TEST INFRASTRUCTURE ONLY: host builds of the product's __host__ __device__ kernel bodies (ReBLUR passes, guide filter, denoiser interface, NEE-AT feedback passes, environment bake, BVH refit, tone mapping)."""
import ctypes as C
import os
import subprocess
import numpy as np
from rtxpt_b200 import structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "tests", "emu", "_build", "libreblur_emu.so")
_lib = None


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s", "_build/libreblur_emu.so"], check=True)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.rb_emu_create.restype = C.c_void_p; L.rb_emu_destroy.argtypes = [C.c_void_p]
        L.rb_emu_denoise.restype = C.c_int
        L.rb_emu_denoise.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(S.ReblurFrame)] + [C.c_void_p] * 9
        _lib = L
    return _lib


class ReblurPort:
    """One instance (history persists between calls) of the host-compiled product passes; same call shape as oracle_lib.Reblur.denoise."""
    def __init__(self): self.h = lib().rb_emu_create()

    def close(self):
        if self.h: lib().rb_emu_destroy(self.h); self.h = None

    def __del__(self): self.close()

    def denoise(self, frame, view_z, normal_roughness, diff, spec, motion=None, disocclusion_mix=None):
        H, W = view_z.shape
        c = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
        vz, nr, d, s, mv, mix = c(view_z, np.float32), c(normal_roughness, np.uint32), c(diff, np.float16), c(spec, np.float16), c(motion, np.float16), c(disocclusion_mix, np.uint8)
        od = np.zeros((H, W, 4), np.float16); os_ = np.zeros((H, W, 4), np.float16); frames = np.zeros((H, W, 2), np.float32)
        ptr = lambda a: None if a is None else a.ctypes.data
        rc = lib().rb_emu_denoise(self.h, W, H, C.byref(frame), ptr(vz), ptr(nr), ptr(mv), ptr(mix), ptr(d), ptr(s), ptr(od), ptr(os_), ptr(frames))
        assert rc == 0
        return od, os_, frames


def denoise_spec_hit_t(depth, spec_hit_t):
    """The product's DenoiseSpecHitT pixel function (rtxpt_b200/csrc/guides_filter.cuh) compiled for the host, ping + pong."""
    H, W = depth.shape
    d = np.ascontiguousarray(depth, np.float32); out = np.array(spec_hit_t, np.float32, copy=True, order="C")
    L = lib(); L.emu_denoise_spec_hit_t.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    assert L.emu_denoise_spec_hit_t(W, H, d.ctypes.data, out.ctypes.data) == 0
    return out


class NeeatPort:
    """The product's NEE-AT feedback passes (rtxpt_b200/csrc/neeat.cuh + neeat_host.h) compiled for the host; state persists between frames like the context's does."""
    def __init__(self, W, H, weights, weights_sum, nee_type=2):
        L = lib(); L.neeat_emu_create.restype = C.c_void_p; L.neeat_emu_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_float, C.c_uint32]
        L.neeat_emu_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]; L.neeat_emu_update_begin.argtypes = [C.c_void_p]; L.neeat_emu_update_end.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.neeat_emu_set_feedback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; L.neeat_emu_destroy.argtypes = [C.c_void_p]
        w = np.ascontiguousarray(weights, np.float32); self.W, self.H, self.n = W, H, len(w)
        self.h = L.neeat_emu_create(W, H, len(w), w.ctypes.data, float(weights_sum), nee_type)

    def close(self):
        if self.h: lib().neeat_emu_destroy(self.h); self.h = None

    def __del__(self): self.close()

    def set_boost(self, flags, light_records, world_to_clip):
        L = lib(); L.neeat_emu_set_boost.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        r = np.ascontiguousarray(light_records, np.uint32); m = None if world_to_clip is None else np.ascontiguousarray(world_to_clip, np.float32)
        assert L.neeat_emu_set_boost(self.h, flags, r.ctypes.data, None if m is None else m.ctypes.data) == 0

    def set_feedback(self, weight, candidate):
        w = np.ascontiguousarray(weight, np.float32); c = np.ascontiguousarray(candidate, np.uint32); assert lib().neeat_emu_set_feedback(self.h, w.ctypes.data, c.ctypes.data) == 0

    def update_begin(self): assert lib().neeat_emu_update_begin(self.h) == 0

    def update_end(self, depth, motion=None):
        d = np.ascontiguousarray(depth, np.float32); m = None if motion is None else np.ascontiguousarray(motion, np.float16)
        assert lib().neeat_emu_update_end(self.h, d.ctypes.data, None if m is None else m.ctypes.data) == 0

    def raw(self, what, dtype, count):
        a = np.zeros(count, dtype); n = lib().neeat_emu_get(self.h, what, a.ctypes.data, a.nbytes); assert n >= 0, (what, n)
        return a[: n // a.itemsize]
