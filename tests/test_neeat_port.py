"""CPU: the product's NEE-AT feedback passes held to the oracle without a GPU.  rtxpt_b200/csrc/neeat.cuh keeps the bodies of LightsBaker's feedback passes as __host__ __device__
functions and neeat_host.h the per-frame bookkeeping; tests/emu compiles both for the host and runs them in the order the C ABI launches the kernels.  The oracle's path tracer
fills the feedback reservoirs (the product's path-tracer side needs a GPU); every frame both sides process the same reservoirs and must agree bit for bit on everything the
passes produce: reservoirs after PreFilter + P0, proxy counts and table, blended and processed reservoirs, sorted tile lists, the seeded reservoirs of the next frame."""
import numpy as np
import pytest


def _compare(o, port, W, H, n_lights, stage):
    P, B = W * H, ((W + 1) // 2) * ((H + 1) // 2)
    ctl_o, ctl_p = o.neeat_raw(8, np.uint32, 8), port.raw(8, np.uint32, 8)
    assert np.array_equal(ctl_o, ctl_p), (stage, ctl_o, ctl_p)                              # tiles, jitter, proxy total, counter, availability, valid feedback count
    T = int(ctl_o[0]) * int(ctl_o[1]) * 128
    for what, dt, n, name in ((0, np.uint32, P, "feedback weight"), (1, np.uint32, P, "feedback candidate"), (7, np.uint32, n_lights, "proxy counters"), (11, np.uint32, int(ctl_o[4]), "proxy table")) + \
                             (((2, np.uint32, P, "processed weight"), (3, np.uint32, P, "processed candidate"), (4, np.uint32, B, "blended weight"), (5, np.uint32, B, "blended candidate"),
                               (6, np.uint32, T, "tile lists")) if stage == "end" else ()):
        a, b = o.neeat_raw(what, dt, n), port.raw(what, dt, n)                              # floats compared by bit pattern
        assert len(a) == len(b) == n and np.array_equal(a, b), (stage, name, int((a != b).sum()))


@pytest.mark.parametrize("moving,boost", [(False, 0), (True, 0), (True, 3), (False, 1)])
def test_feedback_passes_equal_the_oracle(oracle, moving, boost):
    from rtxpt_b200 import scene_builder as sb, scenes
    import host_build_lib as emu
    W, H = 90, 58                                                                            # neither a multiple of the 8-pixel tile nor of the 2-pixel blend
    scene, cam0 = scenes.light_gallery(W, H, bays=7)
    guide = oracle.Oracle(scene)                                                             # depth / motion guides of every frame (no feedback involved)
    o = oracle.Oracle(scene)
    c = sb.make_constants(W, H, cam0, bounce_count=2, diffuse_bounce_count=2); c.NEEATFeedback = 1; c.NEEATImportanceBoost = boost
    o.set_constants(c); o.set_view(sb.world_to_clip(cam0)); o.neeat_reset()
    n_lights = int(o.neeat_raw(12, np.uint32, 1)[0])
    port = emu.NeeatPort(W, H, o.neeat_raw(9, np.float32, n_lights), float(o.neeat_raw(10, np.float32, 1)[0]))
    prev = None
    for f in range(6):
        cam = sb.bridge_camera(W, H, pos=(7.0 + (0.4 * f if moving else 0.0), 1.3, -7.5), direction=(0, -0.02, 1), up=(0, 1, 0), fov_y=0.8)
        cg = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2); guide.set_constants(cg); guide.set_view(sb.world_to_clip(cam))
        g = guide.render_realtime(sb.make_realtime_constants(W, H, cam, prev_cam=prev, bounce_count=2, sub_samples=1))
        c = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2, sample_base_index=f); c.NEEATFeedback = 1; c.NEEATImportanceBoost = boost
        o.set_constants(c); o.set_view(sb.world_to_clip(cam))
        port.set_boost(boost, o.neeat_raw(15, np.uint32, n_lights * 8), sb.world_to_clip(cam))
        if f > 0: port.set_feedback(o.neeat_raw(0, np.float32, W * H), o.neeat_raw(1, np.uint32, W * H))        # what the path tracer left in the reservoirs last frame
        o.neeat_update_begin(); port.update_begin(); _compare(o, port, W, H, n_lights, "begin")
        assert np.array_equal(o.neeat_raw(13, np.uint32, n_lights), port.raw(13, np.uint32, n_lights)) and np.array_equal(o.neeat_raw(14, np.uint32, 1), port.raw(14, np.uint32, 1))      # boosted weights, their sum
        if boost & 1: base = o.neeat_raw(9, np.float32, n_lights); bw = o.neeat_raw(13, np.float32, n_lights); assert (bw[base > 0] > base[base > 0] * 0.999).all() and bw.max() > base.max() * 1.5
        o.neeat_update_end(g["depth"], g["motion"]); port.update_end(g["depth"], g["motion"]); _compare(o, port, W, H, n_lights, "end")
        o.render(0, 1)                                                                       # reference-mode radiance pass: NEE draws local + global candidates and inserts feedback
        prev = cam
    # the sampler-side functions the NEEAT shade kernel calls, on the tile lists both sides now hold: same light, same pdfs, for lights inside and outside the tile
    import ctypes as C
    Lo, Le = oracle.lib(), emu.lib()
    for L_, fn in ((Lo, "oracle_neeat_sample_local"), (Le, "neeat_emu_sample_local")): getattr(L_, fn).argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, C.c_void_p]
    rng = np.random.default_rng(9); hits = 0
    for _ in range(400):
        px, py, rnd = int(rng.integers(0, W)), int(rng.integers(0, H)), float(rng.random()); a = np.zeros(3, np.float32); b = np.zeros(3, np.float32)
        probe = int(rng.integers(n_lights - 28, n_lights))
        assert Lo.oracle_neeat_sample_local(o.h, px, py, rnd, probe, a.ctypes.data) == 0 and Le.neeat_emu_sample_local(port.h, px, py, rnd, probe, b.ctypes.data) == 0
        assert np.array_equal(a, b) and 0 < a[1] <= 1 and a[0] < n_lights, (px, py, a, b)
        hits += a[2] > 0
    assert 0 < hits < 400                                                                   # the binary search both finds and misses
    st = o.neeat_get()
    # the binary search, every kind of tile, both sides, against a brute-force count - and, for a light below every key of the tile, against what the reference's search does
    # (LightingAlgorithms.hlsli:654-682: 8 steps, no empty-range test; pinned by tests/golden/sampler_golden.npz): its last step reads the word just before the tile, i.e. the
    # previous tile's last entry, or for tile (0,0) address 0x7FFFFFFF, which D3D reads as 0 = "light 0, count 1" (a raw pointer would fault: the round-1 GPU fault)
    lists = st["local"].reshape(st["tiles"][1], st["tiles"][0], 128); flat = lists.reshape(-1); jx, jy = st["jitter"]
    for px, py in [(0, 0), (7 - jx, 0), (0, 7 - jy), (W - 1, H - 1), (W // 2, H // 2), (8, 0), (0, 8)]:
        ty, tx = (py + jy) // 8, (px + jx) // 8; tile = lists[ty, tx]; keys = tile >> 9; base = (tx + ty * st["tiles"][0]) * 128
        before = int(flat[base - 1]) if base > 0 else 0
        for probe in (0, 1, int(keys.min()) - 1, int(keys.min()), int(keys.max()), int(keys.max()) + 1, n_lights - 1, before >> 9):
            a = np.zeros(3, np.float32); b = np.zeros(3, np.float32)
            assert Lo.oracle_neeat_sample_local(o.h, px, py, 0.5, probe, a.ctypes.data) == 0 and Le.neeat_emu_sample_local(port.h, px, py, 0.5, probe, b.ctypes.data) == 0
            expect = (keys == probe).sum() / 128.0
            if probe < int(keys.min()) and (before >> 9) == probe: expect = ((before & 0x1FF) + 1) / 128.0
            assert a[2] == b[2] == np.float32(expect), (px, py, probe, a, b)
    assert st["available"] and st["valid_feedback"] > 0.3 * W * H
    if moving: assert np.abs(g["motion"][..., 0].astype(np.float32)).mean() > 1.0                # the dolly really exercised reprojection (whole-pixel shifts)
    port.close(); o.close(); guide.close()
