"""CPU: RTXPT's side of the denoiser interface (DENOISER_PREPARE_INPUTS and DENOISER_FINAL_MERGE) - the product's pixel bodies (rtxpt_b200/csrc/denoiser_iface.cuh,
the functions k_dn_prepare_inputs / k_dn_final_merge wrap) compiled for the host and run on the stable planes an oracle frame produced, against the oracle's own restatement.
Same libm and no contraction on either side: every NRD input image and the merged colour must agree bit for bit."""
import ctypes as C
import numpy as np
import pytest


def _tables(realtime, denoiser):
    r = (C.c_void_p * 7)(*[realtime[n].ctypes.data for n in ("planes", "header", "stable_radiance", "depth", "motion", "throughput", "spec_hit_t")])
    d = (C.c_void_p * 8)(*[denoiser[n].ctypes.data for n in ("view_z", "motion", "normal_roughness", "diff", "spec", "disocclusion_mix", "history_clamp_relax", "output")])
    return r, d


@pytest.mark.parametrize("suppress", [0.0, 0.6])
def test_prepare_inputs_and_final_merge_equal_the_oracle(oracle, suppress):
    from rtxpt_b200 import scene_builder as sb, scenes
    import host_build_lib as emu
    W, H = 88, 72
    scene, cam = scenes.cornell_box(W, H, delta_surfaces=True)
    consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=3)
    o = oracle.Oracle(scene); o.set_constants(consts); o.set_view(sb.world_to_clip(cam))
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=2)
    r = o.render_realtime(rt)
    k = sb.make_denoiser_constants(cam, suppress_primary_indirect_specular_k=suppress)
    a = o.new_denoiser_targets(); b = {n: v.copy() for n, v in a.items()}
    f = emu.lib().emu_denoiser_interface
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    seen = 0
    for i, plane in enumerate((2, 1, 0)):
        o.denoiser_prepare_inputs(rt, k, r, a, plane, i == 0)
        rb, db = _tables(r, b)
        assert f(C.byref(consts), C.byref(rt), C.byref(k), plane, int(i == 0), 0, rb, db, None, None) == 0
        surf = a["view_z"] < 1e30; seen += int(surf.sum())
        for n in ("view_z", "normal_roughness", "disocclusion_mix", "history_clamp_relax", "output"): assert np.array_equal(a[n], b[n]), (plane, n, int((a[n] != b[n]).sum()))
        for n in ("motion", "diff", "spec"): assert np.array_equal(a[n][surf].view(np.uint16), b[n][surf].view(np.uint16)), (plane, n)      # non-surface texels are left as they were
        dd, ds = a["diff"].copy(), a["spec"].copy()
        o.denoiser_final_merge(rt, r, a, plane, dd, ds)
        assert f(C.byref(consts), C.byref(rt), None, plane, 0, 1, rb, db, dd.ctypes.data, ds.ctypes.data) == 0
        assert np.array_equal(a["output"].view(np.uint16), b["output"].view(np.uint16)), (plane, "merged")
    assert seen > 0.6 * W * H                                                       # the planes carried real surfaces
    o.close()
