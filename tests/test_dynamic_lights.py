"""Dynamic light lists in NEE-AT's feedback loop (LightsBaker.cpp:1086-1225, LightsBaker.hlsl u_historyRemapPastToCurrent / u_historyRemapCurrentToPast):
analytic lights are added to and dropped from the scene between frames (rtxpt_b200_update_lights), the emissive triangles' indices shift with them, and last frame's feedback
reservoirs, usage counters and tile samplers must keep naming the same PHYSICAL lights.  CPU: the oracle's bookkeeping against the light records themselves; GPU: the CUDA passes
against the oracle, state by state."""
import numpy as np
import pytest

W, H = 90, 58
E = 5368                      # environment quad-tree nodes come first in the light list


def _lights(n):
    from rtxpt_b200 import scene_builder as sb
    pos = [(1.5, 2.2, 2.0), (4.0, 2.4, 3.0), (6.5, 2.0, 1.5), (8.0, 2.5, 2.5)]
    return sb.make_light_array([sb.point_light(pos[i], (1.0, 0.9 - 0.1 * i, 0.7), 30.0 + 10 * i, 0.12) for i in range(n)])


def _oracle(oracle):
    from rtxpt_b200 import scene_builder as sb, scenes
    scene, cam = scenes.light_gallery(W, H, bays=7)
    consts = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2); consts.NEEATFeedback = 1; consts.NEEATImportanceBoost = 3
    o = oracle.Oracle(scene); o.set_lights(_lights(2)); o.set_constants(consts); o.set_view(sb.world_to_clip(cam)); o.neeat_reset()
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=2, sub_samples=1)
    return scene, cam, consts, o, rt


def _records(o):
    li, _, _ = o.lights(); return np.ascontiguousarray(li).view(np.uint8).reshape(len(li), -1)


def test_oracle_feedback_follows_lights_through_list_changes(oracle):
    scene, cam, consts, o, rt = _oracle(oracle)
    for f in range(5):
        consts.sampleBaseIndex = f; o.set_constants(consts); o.neeat_update_begin(); o.render_realtime(rt)
    for n_next in (3, 1, 2):                                   # add one at the end, drop two, add one back
        before = _records(o); cand_a = o.neeat_raw(1, np.uint32, W * H).copy(); w_a = o.neeat_raw(0, np.float32, W * H).copy()
        n_prev = len(before) - E - (len(before) - E - _analytic(before))
        o.set_lights(_lights(n_next)); consts.sampleBaseIndex += 1; o.set_constants(consts)
        after = _records(o); assert len(after) - len(before) == n_next - n_prev
        o.neeat_update_begin()                                 # P0: candidates remapped into the new list, world-space-coherent ones stripped
        cand_b = o.neeat_raw(1, np.uint32, W * H); w_b = o.neeat_raw(0, np.float32, W * H)
        # (the pre-filter lets a reservoir adopt a neighbour's candidate, so the comparison is between the SETS of physical lights named before and after)
        had = (w_a > 0) & (cand_a != 0xFFFFFFFF); ia = np.unique(cand_a[had] & 0x7FFFFFFF)
        kept = (w_b > 0) & (cand_b != 0xFFFFFFFF); ib = np.unique(cand_b[kept] & 0x7FFFFFFF)
        assert len(ib) > 20 and kept.sum() > 200
        named_before = {before[i].tobytes() for i in ia}
        assert all(after[i].tobytes() in named_before for i in ib)                   # every surviving reservoir names a physical light (identical 32-byte record) that was named before
        delta = n_next - n_prev
        tri_b = ib[ib >= E + n_next].astype(np.int64); assert len(tri_b) > 10 and np.isin(tri_b - delta, ia.astype(np.int64)).all()       # emissive triangles moved by the change of the analytic count
        if delta: assert np.mean([after[i].tobytes() == before[i].tobytes() for i in tri_b]) < 0.2        # (an identity remap would have named other triangles)
        gone = ia[(ia >= E + min(n_prev, n_next)) & (ia < E + n_prev)]                # lights that were dropped: nothing maps onto their old slots' successors by accident
        assert all(before[i].tobytes() not in {after[j].tobytes() for j in ib} for i in gone)
        o.render_realtime(rt)                                  # update_end + FILL on the new list: tile samplers hold valid current indices
        lists = o.neeat_raw(6, np.uint32, 1 << 22); assert (lists >> 9).max() < len(after)
        # a frame later the list stands still: identity again, still consistent
        consts.sampleBaseIndex += 1; o.set_constants(consts); c0 = o.neeat_raw(1, np.uint32, W * H).copy(); o.neeat_update_begin(); c1 = o.neeat_raw(1, np.uint32, W * H)
        assert np.isin(np.unique(c1[c1 != 0xFFFFFFFF] & 0x7FFFFFFF), np.unique(c0[c0 != 0xFFFFFFFF] & 0x7FFFFFFF)).all()        # nothing is renamed when nothing changed
        o.render_realtime(rt)
    o.close()


def _analytic(records):
    """Number of analytic (sphere / point) lights in a light list: type field (bits 24..27 of ColorTypeAndFlags) is neither environment quad (5) nor triangle."""
    t = (records[:, 12:16].copy().view(np.uint32)[:, 0] >> 24) & 0xF
    tri_type = t[-1]                                            # the list ends with emissive triangles in these scenes
    return int(((t != 5) & (t != tri_type)).sum())


@pytest.mark.gpu
def test_gpu_feedback_passes_follow_the_oracle_through_list_changes(product, oracle):
    """The CUDA baker passes after rtxpt_b200_update_lights: fed the oracle's reservoirs frame after frame (as in test_gpu_neeat.py), every buffer the passes derive - remapped
    reservoirs, usage counters, proxy table, blended reservoirs, tile lists - is bit-identical to the oracle's, across an insertion, a removal and a re-insertion."""
    from rtxpt_b200 import scene_builder as sb
    from test_gpu_neeat import _same_state
    scene, cam, consts, o, rt = _oracle(oracle)
    c = product.Context(max_sub_samples_per_launch=1, strict=True); c.upload_scene(scene); c.update_lights(_lights(2)); c.set_constants(consts); c.set_view(sb.world_to_clip(cam)); c.set_realtime(rt)
    li_p, _, _ = c.lights(); li_o, _, _ = o.lights(); assert np.array_equal(li_p, li_o)
    counts = [2, 2, 2, 3, 3, 1, 1, 2, 2]
    for f, n in enumerate(counts):
        if f and n != counts[f - 1]: o.set_lights(_lights(n)); c.update_lights(_lights(n))
        consts.sampleBaseIndex = f; c.set_constants(consts); o.set_constants(consts)
        li_p, ct_p, px_p = c.lights(); li_o, ct_o, px_o = o.lights(); assert np.array_equal(li_p, li_o) and np.array_equal(px_p, px_o)
        n_lights = len(li_o)
        if f > 0: c.neeat_set_feedback(o.neeat_raw(0, np.float32, W * H), o.neeat_raw(1, np.uint32, W * H))
        o.neeat_update_begin(); c.neeat_update_begin(); c.synchronize(); _same_state(c, o, W, H, n_lights, "begin")
        r = o.render_realtime(rt); c.path_trace_realtime(True); c.synchronize()
        assert np.allclose(c.readback_realtime()["depth"], r["depth"], rtol=1e-5, atol=1e-6)
    c.close(); o.close()
