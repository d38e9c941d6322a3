"""Diagnostic for tests/test_gpu_realtime.py::test_config3_frame_split_over_two_ranks_equals_the_single_gpu_frame: where do a split frame and a single-context frame differ?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from rtxpt_b200 import lib as product, scene_builder as sb, scenes, structs as S, realtime_mgpu as M, tiles

W, H = 160, 128
scene, cam = scenes.cornell_box(W, H, delta_surfaces=True)
consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=3)
k = sb.make_denoiser_constants(cam)
for strict in (False, True):
    def make(rank, world, tile=32):
        c = product.Context(max_sub_samples_per_launch=1, tile_rank=rank, tile_world=world, tile_size=tile, strict=strict); c.upload_scene(scene); c.set_constants(consts); c.set_view(sb.world_to_clip(cam))
        c.set_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=2)); return c
    one, one64, two = make(0, 1), make(0, 1, 64), [make(0, 2), make(1, 2)]
    group = M.LocalGroup(two); tables, _ = tiles.gather_layout(W, H, 32, 2)
    own = np.zeros((H, W), np.int32)
    for r, t in enumerate(tables): own[t & 0xFFFF, t >> 16] = r
    frame = sb.make_reblur_frame(cam, cam, frame_index=0)
    for c in (one, one64):
        c.path_trace_realtime(False); c.synchronize()
    rt1, rt64 = one.readback_realtime(), one64.readback_realtime()
    for key in ("header", "stable_radiance", "depth", "motion", "spec_hit_t", "planes"):
        a, b = rt1[key], rt64[key]
        print("strict", strict, "tile 32 vs 64, after trace:", key, "equal" if a.tobytes() == b.tobytes() else "DIFFER (%d bytes)" % (np.frombuffer(a.tobytes(), np.uint8) != np.frombuffer(b.tobytes(), np.uint8)).sum())
    one.path_trace_realtime(False); one.synchronize(); again = one.readback_realtime()
    print("strict", strict, "replay of the same context: planes", "equal" if again["planes"].tobytes() == rt1["planes"].tobytes() else "DIFFER", "header", "equal" if again["header"].tobytes() == rt1["header"].tobytes() else "DIFFER")
    for c in two: c.path_trace_realtime(False)
    for c in two: c.synchronize()
    for r, c in enumerate(two):
        g = c.readback_realtime(); mine = own == r
        for key in ("stable_radiance", "depth", "motion", "spec_hit_t"):
            a, b = rt1[key], g[key]
            d = (a != b); d = d.any(-1) if d.ndim == 3 else d
            print("strict", strict, "rank", r, "after trace:", key, "own pixels differing:", int((d & mine).sum()), "of", int(mine.sum()))
        hd = (rt1["header"] != g["header"]).any(0); print("strict", strict, "rank", r, "header own differing:", int((hd & mine).sum()))
    one.denoise_realtime(k, frame); one.synchronize()
    # the split recipe from the guides on
    M.realtime_frame.__globals__  # noqa
    moved = 0
    group.exchange(M.GUIDES); group.each(lambda c: c.denoise_spec_hit_t())
    for c in two: c.synchronize()
    a = one.readback_realtime()["spec_hit_t"]  # after denoise_realtime the filtered guide
    for r, c in enumerate(two): print("strict", strict, "rank", r, "filtered spec_hit_t differing:", int((c.readback_realtime()["spec_hit_t"] != a).sum()), "depth differing:", int((c.readback_guides()[0] != one.readback_guides()[0]).sum()))
    first = True
    for plane in (2, 1, 0):
        group.each(lambda c: c.denoiser_prepare_inputs(plane, first, k)); group.exchange(M.NRD_INPUTS)
        group.each(lambda c: c.reblur_denoise(plane, frame)); group.each(lambda c: c.denoiser_final_merge(plane, identity=False)); first = False
    for c in two: c.synchronize()
    ref = one.readback_denoiser_inputs()
    for r, c in enumerate(two):
        got = c.readback_denoiser_inputs(); mine = own == r
        for name in ref:
            d = ref[name] != got[name]; d = d.any(-1) if d.ndim == 3 else d
            ys, xs = np.nonzero(d)
            print("strict", strict, "rank", r, "NRD input", name, "differing:", int(d.sum()), "own:", int((d & mine).sum()), "first:", list(zip(xs[:4].tolist(), ys[:4].tolist())))
    for c in [one, one64] + two: c.close()
