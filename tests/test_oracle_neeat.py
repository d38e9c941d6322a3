"""CPU: NEE-AT temporal feedback in the oracle (oracle/pt_neeat.h): feedback reservoirs -> usage-weighted global proxy table + per-tile local samplers ->
local / global candidate mix with MIS in NEE.  The reference has no golden vectors for this loop (its float atomics and cross-group races make a frame's state order-dependent),
so the restatement is held to what the algorithm guarantees: structural invariants of the tile lists, the response of both samplers to a known feedback image, an unbiased
estimator (same mean as global-only sampling) and lower error once the loop has warmed up."""
import numpy as np
import pytest

INVALID = 0xFFFFFFFF
SSC = 0x80000000


def _setup(oracle, W, H, bays=8, feedback=True, bounces=2):
    from rtxpt_b200 import scene_builder as sb, scenes
    scene, cam = scenes.light_gallery(W, H, bays=bays)
    o = oracle.Oracle(scene)
    c = sb.make_constants(W, H, cam, bounce_count=bounces, diffuse_bounce_count=bounces); c.NEEATFeedback = 1 if feedback else 0
    o.set_constants(c); o.set_view(sb.world_to_clip(cam))
    if feedback: o.neeat_reset()
    return o, c, cam, sb.make_realtime_constants(W, H, cam, bounce_count=bounces, sub_samples=1)


def _frames(o, c, rt, n, feedback, first=0):
    out = []
    for f in range(first, first + n):
        c.sampleBaseIndex = f; o.set_constants(c)
        if feedback: o.neeat_update_begin()
        out.append(o.render_realtime(rt)["merged"].copy())
    return np.stack(out)


def test_tile_lists_and_counters_are_well_formed(oracle):
    W, H = 96, 64
    o, c, cam, rt = _setup(oracle, W, H)
    _frames(o, c, rt, 4, True)
    st = o.neeat_get()
    assert st["tiles"] == ((W + 7) // 8 + 1, (H + 7) // 8 + 1) and st["available"] and 0 < st["valid_feedback"] <= W * H
    assert all(0 <= j < 8 for j in st["jitter"])
    light_count = 5368 + 8 * 2 * 2                                                     # environment quad-tree slots + two triangles per lamp
    lists = st["local"].reshape(-1, 128); lights = lists >> 9; counts = (lists & 0x1FF) + 1
    assert (np.diff(lights.astype(np.int64), axis=1) >= 0).all()                       # sorted ascending: SampleLocalPDF binary-searches them
    assert lights.max() < light_count and lights.min() >= 5368                         # only real (emissive triangle) lights: the environment map is off
    for row_l, row_c in zip(lights[::7], counts[::7]):
        u, n = np.unique(row_l, return_counts=True)
        assert n.sum() == 128 and all((row_c[row_l == k] == m).all() for k, m in zip(u, n))     # every entry carries the run length of its light: pdf = count / 128
    cand = st["candidate"]; has = st["weight"] > 0
    assert ((cand[has] & np.uint32(0x7FFFFFFF)) < light_count).all()                                     # reservoirs the next frame starts from hold valid lights
    assert (st["scratch_candidate"] != INVALID).all() and (st["blended_candidate"] != INVALID).all()      # P1a / P1b never leave a hole for FillTile
    o.close()


def test_samplers_follow_a_known_feedback_image(oracle):
    """All left-half pixels report lamp A, all right-half pixels lamp B: the global table shifts its proxies to A and B (75 % usage weight), left tiles hold mostly A, right
    tiles mostly B, and nothing else gets more than the power-based remainder."""
    W, H = 96, 64
    o, c, cam, rt = _setup(oracle, W, H)
    c.sampleBaseIndex = 0; o.set_constants(c); o.neeat_update_begin()
    base = o.neeat_proxy_counters(5368 + 32).astype(np.int64)
    r = o.render_realtime(rt)
    A, B = 5368 + 3, 5368 + 20
    cand = np.full((H, W), A | SSC, np.uint32); cand[:, W // 2:] = B | SSC
    o.neeat_set_feedback(np.ones((H, W), np.float32), cand)
    c.sampleBaseIndex = 1; o.set_constants(c); o.neeat_update_begin()
    cnt = o.neeat_proxy_counters(5368 + 32).astype(np.int64)
    total = cnt.sum(); others = np.delete(cnt, [A, B])
    assert abs(cnt[A] / total - (0.25 * base[A] / base.sum() + 0.375)) < 0.01 and abs(cnt[B] / total - (0.25 * base[B] / base.sum() + 0.375)) < 0.01
    assert np.allclose(others / total, 0.25 * np.delete(base, [A, B]) / base.sum(), atol=2e-3)
    o.neeat_update_end(r["depth"], r["motion"])
    st = o.neeat_get()
    lights = st["local"] >> 9
    tx = st["tiles"][0]; jx = st["jitter"][0]
    left = lights[:, : (W // 2 - 16 + jx) // 8]; right = lights[:, (W // 2 + 16 + jx) // 8 + 1: tx - 1]
    assert (left == A).mean() > 0.9 and (right == B).mean() > 0.9                       # 64 window pixels + 64 top-up picks from the blended image around the tile
    assert st["valid_feedback"] == W * H
    o.close()


def test_feedback_is_unbiased_and_reduces_error(oracle):
    W, H = 112, 72
    o, c, cam, rt = _setup(oracle, W, H, bays=10, feedback=True)
    a = _frames(o, c, rt, 72, True); o.close()
    o, c, cam, rt = _setup(oracle, W, H, bays=10, feedback=False)
    b = _frames(o, c, rt, 144, False); o.close()
    ref = b.mean(0); warm = a[12:]
    assert abs(warm.mean() / b.mean() - 1) < 0.03                                       # same estimator mean as global-only sampling
    noise = np.abs(b[:72].mean(0) - b[72:].mean(0)).mean()
    assert np.abs(warm.mean(0) - ref).mean() < 1.6 * noise                              # and per pixel, within the noise of the comparison itself
    e_fb, e_gl = np.abs(warm - ref).mean(), np.abs(b - ref).mean()
    assert e_fb < 0.93 * e_gl, (e_fb, e_gl)                                             # local candidates find the bay's own lamps: lower per-frame error
    assert np.isfinite(a).all()


def test_reference_mode_loop_is_unbiased(oracle):
    """Reference mode (RTXPT's default NEEType 2): update_begin, update_end on the previous frame's guides, then the radiance pass.  Same mean as feedback-free sampling."""
    from rtxpt_b200 import scene_builder as sb, scenes
    W, H = 96, 64
    scene, cam = scenes.light_gallery(W, H, bays=8)
    guide = oracle.Oracle(scene); cg = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2); guide.set_constants(cg); guide.set_view(sb.world_to_clip(cam))
    g = guide.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=2, sub_samples=1)); guide.close()          # static camera: one set of depth / motion guides
    def frames(n, feedback):
        o = oracle.Oracle(scene); c = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2); c.NEEATFeedback = 1 if feedback else 0
        o.set_constants(c); o.set_view(sb.world_to_clip(cam))
        if feedback: o.neeat_reset()
        out = []
        for f in range(n):
            c.sampleBaseIndex = 2 * f; o.set_constants(c)
            if feedback: o.neeat_update_begin(); o.neeat_update_end(g["depth"], g["motion"])
            acc = o.render(0, 2)[0]; out.append(acc[..., :3].copy())
        st = o.neeat_get() if feedback else None
        o.close(); return np.stack(out), st
    fb, st = frames(60, True); gl, _ = frames(120, False)
    assert st["available"] and st["valid_feedback"] > 0.3 * W * H
    assert abs(fb[10:].mean() / gl.mean() - 1) < 0.03
    ref = gl.mean(0)
    assert np.abs(fb[10:] - ref).mean() < 0.97 * np.abs(gl - ref).mean()


def test_frustum_booster_moves_proxies_into_view(oracle):
    """ImportanceBooster: with the frustum booster on, lamps in (or within 5 units of) the view frustum take proxies from those far outside it."""
    from rtxpt_b200 import scene_builder as sb, scenes
    W, H = 64, 48
    scene, _ = scenes.light_gallery(W, H, bays=12)
    cam = sb.bridge_camera(W, H, pos=(3.0, 1.3, -2.0), direction=(0, 0, 1), up=(0, 1, 0), fov_y=0.5)          # close up on the leftmost bays; the right end is ~20 units off
    def counters(boost):
        o = oracle.Oracle(scene); c = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2); c.NEEATFeedback = 1; c.NEEATImportanceBoost = boost
        o.set_constants(c); o.set_view(sb.world_to_clip(cam)); o.neeat_reset(); o.neeat_update_begin()
        n = int(o.neeat_raw(12, np.uint32, 1)[0]); cnt = o.neeat_proxy_counters(n).astype(np.float64); rec = o.neeat_raw(15, np.uint32, n * 8).reshape(n, 8); o.close()
        return cnt, rec[:, 0].view(np.float32)
    plain, x = counters(0); boosted, _ = counters(1)
    tri = np.arange(len(x)) >= 5368                                                      # the lamps' triangles
    near, far = tri & (x < 6.0), tri & (x > 16.0)
    assert near.any() and far.any()
    share = lambda c, m: c[m].sum() / c[tri].sum()
    assert share(boosted, near) > 1.5 * share(plain, near) and share(boosted, far) < 0.7 * share(plain, far)
