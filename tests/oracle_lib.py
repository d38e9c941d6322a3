"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY (never imported by rtxpt_b200/)."""
import ctypes as C
import os
import subprocess
import numpy as np
from rtxpt_b200 import structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle.so")


class RenderStats(C.Structure):
    _fields_ = [("scatterRays", C.c_uint64), ("shadowRays", C.c_uint64), ("nodeVisits", C.c_uint64), ("triTests", C.c_uint64),
                ("seconds", C.c_double), ("threads", C.c_int)]


def build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "_build/liboracle.so"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_hash32.restype = C.c_uint32; L.oracle_hash32.argtypes = [C.c_uint32]
        L.oracle_hash32_combine.restype = C.c_uint32; L.oracle_hash32_combine.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_hash32_to_float.restype = C.c_float; L.oracle_hash32_to_float.argtypes = [C.c_uint32]
        L.oracle_sobol.restype = C.c_uint32; L.oracle_sobol.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_owen_scramble.restype = C.c_uint32; L.oracle_owen_scramble.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_f32tof16.restype = C.c_uint32; L.oracle_f32tof16.argtypes = [C.c_float]
        L.oracle_f16tof32.restype = C.c_float; L.oracle_f16tof32.argtypes = [C.c_uint32]
        L.oracle_pack_snorm8.restype = C.c_uint32; L.oracle_pack_snorm8.argtypes = [C.c_float]
        L.oracle_unpack_snorm8.restype = C.c_float; L.oracle_unpack_snorm8.argtypes = [C.c_uint32]
        L.oracle_create.restype = C.c_void_p; L.oracle_create.argtypes = [C.POINTER(S.SceneDesc)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_bvh_build_seconds.restype = C.c_double; L.oracle_bvh_build_seconds.argtypes = [C.c_void_p]
        L.oracle_triangle_count.restype = C.c_uint32; L.oracle_triangle_count.argtypes = [C.c_void_p]
        L.oracle_set_constants.argtypes = [C.c_void_p, C.POINTER(S.PathTracerConstants)]
        L.oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        L.oracle_get_lights.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.oracle_set_view.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_render_guides.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_get_lights_ex.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        L.oracle_get_sub_instances.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.oracle_render.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                    C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_int, C.POINTER(RenderStats)]
        L.oracle_render_realtime.argtypes = [C.c_void_p, C.POINTER(S.RealtimeConstants)] + [C.c_uint32] * 4 + [C.c_void_p] * 8 + [C.c_int]
        for f in ("oracle_branch_advance", "oracle_branch_vertex_index", "oracle_generic_ts_line_stride", "oracle_generic_ts_plane_stride"):
            getattr(L, f).restype = C.c_uint32
        L.oracle_branch_advance.argtypes = [C.c_uint32] * 2; L.oracle_branch_vertex_index.argtypes = [C.c_uint32]
        L.oracle_branch_on_stable_path.restype = C.c_uint32; L.oracle_branch_on_stable_path.argtypes = [C.c_uint32] * 4
        L.oracle_generic_ts_address.restype = C.c_uint32; L.oracle_generic_ts_address.argtypes = [C.c_uint32] * 5
        L.oracle_generic_ts_line_stride.argtypes = [C.c_uint32] * 2; L.oracle_generic_ts_plane_stride.argtypes = [C.c_uint32] * 2
        L.oracle_pack_ortho.argtypes = [C.c_void_p, C.c_void_p]; L.oracle_unpack_ortho.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_denoiser_prepare_inputs.argtypes = [C.c_void_p, C.POINTER(S.RealtimeConstants), C.POINTER(S.DenoiserConstants), C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_denoiser_final_merge.argtypes = [C.c_void_p, C.POINTER(S.RealtimeConstants), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_reblur_spatial.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint32] + [C.c_void_p] * 4
        L.oracle_reblur_create.restype = C.c_void_p; L.oracle_reblur_destroy.argtypes = [C.c_void_p]
        L.oracle_reblur_denoise.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint32, C.c_int] + [C.c_void_p] * 9
        L.oracle_tri_info.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_rng.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.oracle_bsdf.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, scene):
        self.scene = scene          # keeps the numpy storage alive
        self.h = lib().oracle_create(C.byref(scene.desc))
        self.consts = None

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def set_constants(self, consts):
        self.consts = consts
        assert lib().oracle_set_constants(self.h, C.byref(consts)) == 0

    def set_view(self, world_to_clip):
        m = np.ascontiguousarray(world_to_clip, np.float32).reshape(16)
        assert lib().oracle_set_view(self.h, m.ctypes.data) == 0

    def render_guides(self, sub_sample, threads=0):
        W, H = self.consts.imageWidth, self.consts.imageHeight
        depth = np.zeros((H, W), np.float32); thp = np.zeros((H, W), np.uint32)
        assert lib().oracle_render_guides(self.h, sub_sample, 0, 0, W, H, depth.ctypes.data, thp.ctypes.data, threads) == 0
        return depth, thp

    def trace_rays(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        hits = np.zeros(len(rays), dtype=[("t", "f4"), ("u", "f4"), ("v", "f4"), ("inst", "u4"), ("geom", "u4"), ("prim", "u4")])
        lib().oracle_trace_rays(self.h, rays.ctypes.data, len(rays), int(any_hit), hits.ctypes.data)
        return hits

    def lights(self):
        n, m = C.c_uint32(0), C.c_uint32(0)
        lib().oracle_get_lights(self.h, None, C.byref(n), None, None, C.byref(m))
        infos = np.zeros((n.value, 8), np.uint32); counters = np.zeros(n.value, np.uint32); proxies = np.zeros(m.value, np.uint32)
        lib().oracle_get_lights(self.h, infos.ctypes.data, C.byref(n), counters.ctypes.data, proxies.ctypes.data, C.byref(m))
        return infos, counters, proxies

    def lights_ex(self):
        n = C.c_uint32(0)
        lib().oracle_get_lights_ex(self.h, None, C.byref(n))
        ex = np.zeros((n.value, 4), np.uint32)
        if n.value:
            lib().oracle_get_lights_ex(self.h, ex.ctypes.data, C.byref(n))
        return ex

    def render(self, first_sub_sample, count, accum=None, accum_count=0, rect=None, threads=0, want_primary=False):
        W, H = self.consts.imageWidth, self.consts.imageHeight
        if accum is None:
            accum = np.zeros((H, W, 4), np.float32)
        x0, y0, x1, y1 = rect if rect else (0, 0, W, H)
        last = np.zeros((H, W, 3), np.float32)
        primary = np.zeros((H, W, 4), np.float32) if want_primary else None
        n = C.c_uint32(accum_count)
        st = RenderStats()
        rc = lib().oracle_render(self.h, first_sub_sample, count, x0, y0, x1, y1, accum.ctypes.data, C.byref(n), last.ctypes.data,
                                 primary.ctypes.data if want_primary else None, threads, C.byref(st))
        assert rc == 0
        return accum, n.value, last, primary, st

    # ---- NEE-AT temporal feedback (oracle/pt_neeat.h); frame order: set_constants; neeat_update_begin; [render_realtime does update_end after its BUILD pass] or neeat_update_end; render
    def set_lights(self, lights):
        """lights: ctypes array of structs.LightDesc (or None / empty): the scene's analytic lights of this frame."""
        n = len(lights) if lights is not None else 0
        assert lib().oracle_set_lights(C.c_void_p(self.h), lights if n else None, n) == 0

    def neeat_reset(self): assert lib().oracle_neeat_reset(C.c_void_p(self.h)) == 0

    def neeat_update_begin(self): assert lib().oracle_neeat_update_begin(C.c_void_p(self.h)) == 0

    def neeat_update_end(self, depth, motion=None):
        d = np.ascontiguousarray(depth, np.float32); m = None if motion is None else np.ascontiguousarray(motion, np.float16)
        assert lib().oracle_neeat_update_end(C.c_void_p(self.h), C.c_void_p(d.ctypes.data), C.c_void_p(None if m is None else m.ctypes.data)) == 0

    def neeat_get(self):
        """All NEE-AT state as numpy arrays."""
        W, H = self.consts.imageWidth, self.consts.imageHeight
        L = lib(); L.oracle_neeat_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]; L.oracle_neeat_get.restype = C.c_int
        ctl = np.zeros(8, np.uint32); assert L.oracle_neeat_get(self.h, 8, ctl.ctypes.data, ctl.nbytes) > 0
        bw, bh = (W + 1) // 2, (H + 1) // 2
        out = dict(tiles=(int(ctl[0]), int(ctl[1])), jitter=(int(ctl[2]), int(ctl[3])), sampling_proxy_count=int(ctl[4]), update_counter=int(ctl[5]), available=bool(ctl[6]), valid_feedback=int(ctl[7]))
        for key, what, shape, dt in (("weight", 0, (H, W), np.float32), ("candidate", 1, (H, W), np.uint32), ("scratch_weight", 2, (H, W), np.float32), ("scratch_candidate", 3, (H, W), np.uint32),
                                     ("blended_weight", 4, (bh, bw), np.float32), ("blended_candidate", 5, (bh, bw), np.uint32), ("local", 6, (int(ctl[1]), int(ctl[0]), 128), np.uint32)):
            a = np.zeros(shape, dt); assert L.oracle_neeat_get(self.h, what, a.ctypes.data, a.nbytes) == a.nbytes, key; out[key] = a
        return out

    def neeat_raw(self, what, dtype, count):
        L = lib(); L.oracle_neeat_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]; L.oracle_neeat_get.restype = C.c_int
        a = np.zeros(count, dtype); n = L.oracle_neeat_get(self.h, what, a.ctypes.data, a.nbytes); assert n >= 0, (what, n)
        return a[: n // a.itemsize]

    def neeat_proxy_counters(self, light_count):
        L = lib(); L.oracle_neeat_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]; L.oracle_neeat_get.restype = C.c_int
        a = np.zeros(light_count, np.uint32); assert L.oracle_neeat_get(self.h, 7, a.ctypes.data, a.nbytes) == a.nbytes
        return a

    def neeat_set_feedback(self, weight, candidate):
        w = np.ascontiguousarray(weight, np.float32); c = np.ascontiguousarray(candidate, np.uint32)
        assert lib().oracle_neeat_set_feedback(C.c_void_p(self.h), C.c_void_p(w.ctypes.data), C.c_void_p(c.ctypes.data)) == 0

    def render_realtime(self, rt, rect=None, threads=0):
        """BUILD + rt.subSampleCount x FILL + no-denoiser merge.  Returns a dict of the realtime render targets (numpy)."""
        W, H = self.consts.imageWidth, self.consts.imageHeight
        plane_stride = lib().oracle_generic_ts_plane_stride(W, H)
        out = dict(planes=np.zeros(3 * plane_stride, S.STABLE_PLANE_DTYPE), header=np.zeros((4, H, W), np.uint32), stable_radiance=np.zeros((H, W, 4), np.float16),
                   depth=np.zeros((H, W), np.float32), motion=np.zeros((H, W, 4), np.float16), throughput=np.zeros((H, W), np.uint32), spec_hit_t=np.zeros((H, W), np.float32),
                   merged=np.zeros((H, W, 3), np.float32))
        assert out["planes"].itemsize == 80
        x0, y0, x1, y1 = rect if rect else (0, 0, W, H)
        rc = lib().oracle_render_realtime(self.h, C.byref(rt), x0, y0, x1, y1, *[out[k].ctypes.data for k in ("planes", "header", "stable_radiance", "depth", "motion", "throughput", "spec_hit_t", "merged")], threads)
        assert rc == 0
        return out

    # ---- RTXPT's side of the denoiser interface ----
    @staticmethod
    def _ptr_table(arrays):
        return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])

    def new_denoiser_targets(self):
        W, H = self.consts.imageWidth, self.consts.imageHeight
        return dict(view_z=np.zeros((H, W), np.float32), motion=np.zeros((H, W, 4), np.float16), normal_roughness=np.zeros((H, W), np.uint32), diff=np.zeros((H, W, 4), np.float16),
                    spec=np.zeros((H, W, 4), np.float16), disocclusion_mix=np.zeros((H, W), np.uint8), history_clamp_relax=np.zeros((H, W), np.uint8), output=np.zeros((H, W, 4), np.float16))

    def denoiser_prepare_inputs(self, rt, k, realtime, denoiser, plane, init_with_stable_radiance):
        r = self._ptr_table([realtime[n] for n in ("planes", "header", "stable_radiance", "depth", "motion", "throughput", "spec_hit_t")])
        d = self._ptr_table([denoiser[n] for n in ("view_z", "motion", "normal_roughness", "diff", "spec", "disocclusion_mix", "history_clamp_relax", "output")])
        assert lib().oracle_denoiser_prepare_inputs(self.h, C.byref(rt), C.byref(k), plane, int(init_with_stable_radiance), r, d) == 0

    def denoiser_final_merge(self, rt, realtime, denoiser, plane, denoised_diff, denoised_spec):
        r = self._ptr_table([realtime[n] for n in ("planes", "header", "stable_radiance", "depth", "motion", "throughput", "spec_hit_t")])
        d = self._ptr_table([denoiser[n] for n in ("view_z", "motion", "normal_roughness", "diff", "spec", "disocclusion_mix", "history_clamp_relax", "output")])
        assert lib().oracle_denoiser_final_merge(self.h, C.byref(rt), plane, r, d, denoised_diff.ctypes.data, denoised_spec.ctypes.data) == 0


def denoise_spec_hit_t(depth, spec_hit_t):
    """DenoisingGuidesBaker::DenoiseSpecHitT of the oracle; returns the filtered guide."""
    H, W = depth.shape
    d = np.ascontiguousarray(depth, np.float32); out = np.array(spec_hit_t, np.float32, copy=True, order="C")
    L = lib(); L.oracle_denoise_spec_hit_t.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    assert L.oracle_denoise_spec_hit_t(W, H, d.ctypes.data, out.ctypes.data) == 0
    return out


def reblur_spatial(world_to_view, view_to_clip, frame_index, view_z, normal_roughness, diff, spec, accumulated_frames=None, stages=0b1111):
    """oracle/reblur.h spatial chain on NRD inputs (numpy: view_z f32 HxW, normal_roughness u32 HxW, diff / spec f16 HxWx4).  Returns (diff, spec, hit distance for tracking, tiles)."""
    H, W = view_z.shape
    out_d = np.zeros((H, W, 4), np.float16); out_s = np.zeros((H, W, 4), np.float16); track = np.zeros((H, W), np.float32); tiles = np.zeros(((H + 15) // 16, (W + 15) // 16), np.uint8)
    m0 = np.ascontiguousarray(world_to_view, np.float32); m1 = np.ascontiguousarray(view_to_clip, np.float32)
    vz = np.ascontiguousarray(view_z, np.float32); nr = np.ascontiguousarray(normal_roughness, np.uint32); d = np.ascontiguousarray(diff, np.float16); s = np.ascontiguousarray(spec, np.float16)
    af = None if accumulated_frames is None else np.ascontiguousarray(accumulated_frames, np.float32)
    rc = lib().oracle_reblur_spatial(W, H, m0.ctypes.data, m1.ctypes.data, frame_index, vz.ctypes.data, nr.ctypes.data, d.ctypes.data, s.ctypes.data, None if af is None else af.ctypes.data,
                                     stages, out_d.ctypes.data, out_s.ctypes.data, track.ctypes.data, tiles.ctypes.data)
    assert rc == 0
    return out_d, out_s, track, tiles


class Reblur:
    """One REBLUR_DIFFUSE_SPECULAR instance of the oracle (history persists between denoise calls)."""
    def __init__(self):
        self.h = lib().oracle_reblur_create()

    def close(self):
        if self.h: lib().oracle_reblur_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def denoise(self, world_to_view, view_to_clip, frame_index, view_z, normal_roughness, diff, spec, prev_world_to_view=None, prev_view_to_clip=None, motion=None, disocclusion_mix=None, reset=False):
        H, W = view_z.shape
        c = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
        m0, m1 = c(world_to_view, np.float32), c(view_to_clip, np.float32)
        p0 = m0 if prev_world_to_view is None else c(prev_world_to_view, np.float32); p1 = m1 if prev_view_to_clip is None else c(prev_view_to_clip, np.float32)
        vz, nr, d, s, mv, mix = c(view_z, np.float32), c(normal_roughness, np.uint32), c(diff, np.float16), c(spec, np.float16), c(motion, np.float16), c(disocclusion_mix, np.uint8)
        od = np.zeros((H, W, 4), np.float16); os_ = np.zeros((H, W, 4), np.float16); frames = np.zeros((H, W, 2), np.float32)
        ptr = lambda a: None if a is None else a.ctypes.data
        rc = lib().oracle_reblur_denoise(self.h, W, H, ptr(m0), ptr(m1), ptr(p0), ptr(p1), frame_index, int(reset), ptr(vz), ptr(nr), ptr(mv), ptr(mix), ptr(d), ptr(s), ptr(od), ptr(os_), ptr(frames))
        assert rc == 0
        return od, os_, frames
