"""CPU: the CUDA ReBLUR source held to the oracle without a GPU.  rtxpt_b200/csrc/reblur_passes.cuh keeps the body of every pass as a __host__ __device__ function; the kernels
wrap them one thread per pixel, and tests/emu/reblur_host_emu.cu compiles the same functions (plus the product's host constants, reblur_host.h) for the host and walks the pixels
in launchReblurFrame's pass order.  With the same libm and no FMA contraction on either side, the port and the oracle (oracle/reblur.h) must agree bit for bit: any mismatch is a
porting mistake (a wrong constant, index, routing of a transient / permanent buffer), not arithmetic.  What only the GPU build has - libdevice, fast-math flags, the launch grid -
stays with tests/test_gpu_reblur.py."""
import numpy as np
import pytest


def _frames(oracle, W, H, delta, bounces=4):
    from rtxpt_b200 import scene_builder as sb, scenes
    scene, cam = scenes.cornell_box(W, H, delta_surfaces=delta)
    o = oracle.Oracle(scene)

    def frame(index, camera, prev_camera, plane):
        c = sb.make_constants(W, H, camera, bounce_count=bounces, diffuse_bounce_count=3); c.sampleBaseIndex = index
        o.set_constants(c); o.set_view(sb.world_to_clip(camera))
        rt = sb.make_realtime_constants(W, H, camera, prev_cam=prev_camera, bounce_count=bounces, sub_samples=1)
        r = o.render_realtime(rt); d = o.new_denoiser_targets()
        o.denoiser_prepare_inputs(rt, sb.make_denoiser_constants(camera), r, d, plane, True)
        return d
    return o, cam, frame


def _both(rb, port, sb, cam, prev, f, d, reset=False, use_motion=True):
    prev_ = cam if prev is None else prev
    od, os_, fr = rb.denoise(sb.world_to_view(cam), sb.view_to_clip(cam), f, d["view_z"], d["normal_roughness"], d["diff"], d["spec"], prev_world_to_view=sb.world_to_view(prev_),
                             prev_view_to_clip=sb.view_to_clip(prev_), motion=d["motion"] if use_motion else None, disocclusion_mix=d["disocclusion_mix"], reset=reset)
    pd, ps, pf = port.denoise(sb.make_reblur_frame(cam, prev_, frame_index=f, reset=reset, ignore_motion_vectors=not use_motion), d["view_z"], d["normal_roughness"], d["diff"], d["spec"],
                              motion=d["motion"], disocclusion_mix=d["disocclusion_mix"])
    return (od, os_, fr), (pd, ps, pf)


def _assert_identical(a, b, surf, label):
    for name, x, y in (("diff", a[0], b[0]), ("spec", a[1], b[1])):
        same = (x.view(np.uint16) == y.view(np.uint16)).all(-1)[surf].mean()
        assert same == 1.0, (label, name, same, np.abs(x.astype(np.float32) - y.astype(np.float32))[surf].max())
    assert np.array_equal(a[2][surf], b[2][surf]), (label, "accumulated frames")


@pytest.mark.parametrize("delta,plane", [(False, 0), (True, 0), (True, 1)])
def test_port_equals_oracle_on_a_dolly(oracle, delta, plane):
    """Six frames of a slow dolly (motion vectors, previous matrices, growing history), then an undescribed cut and a reset: all eight passes, both reprojection paths, the
    disocclusion tests and the history bookkeeping, bit for bit."""
    from rtxpt_b200 import scene_builder as sb
    import host_build_lib as emu
    W, H = 80, 64
    o, cam0, frame = _frames(oracle, W, H, delta)
    cam_at = lambda dx: sb.bridge_camera(W, H, pos=(2.78 + dx, 2.73, -8.0), direction=(0, 0, 1), up=(0, 1, 0), fov_y=0.66)
    rb = oracle.Reblur(); port = emu.ReblurPort(); prev = None
    for f in range(6):
        cam = cam_at(0.03 * f)
        d = frame(f, cam, prev, plane)
        a, b = _both(rb, port, sb, cam, prev, f, d)
        surf = d["view_z"] < 1e30
        _assert_identical(a, b, surf, (delta, plane, f))
        prev = cam
    if surf.any(): assert a[2][surf].max() >= 3                                        # the history did grow, so the temporal paths were exercised
    cut = sb.bridge_camera(W, H, pos=(4.2, 1.2, -6.0), direction=(-0.25, 0.1, 1), up=(0, 1, 0), fov_y=0.66)
    d = frame(20, cut, cut, plane)
    a, b = _both(rb, port, sb, cut, cut, 6, d, use_motion=False); _assert_identical(a, b, d["view_z"] < 1e30, (delta, plane, "cut"))
    d = frame(21, cut, cut, plane)
    a, b = _both(rb, port, sb, cut, cut, 7, d, reset=True); _assert_identical(a, b, d["view_z"] < 1e30, (delta, plane, "reset"))
    assert (b[2] == 0).all()
    rb.close(); port.close(); o.close()


def test_port_equals_oracle_with_sky_roughness_and_missing_hits(oracle):
    """Synthetic NRD inputs that the Cornell box lacks: sky tiles and a ragged sky border, a roughness ramp from mirror to diffuse, two material ids, pixels without a hit distance,
    an image size that is not a multiple of the 16x16 tile."""
    from rtxpt_b200 import scene_builder as sb
    import host_build_lib as emu
    from test_oracle_reblur import pack_normal_roughness
    W, H = 90, 52
    cam = sb.bridge_camera(W, H, pos=(0, 0, 0), direction=(0, 0, 1), up=(0, 1, 0), fov_y=0.9)
    rng = np.random.default_rng(11)
    ys, xs = np.mgrid[0:H, 0:W]
    vz = (4.0 + 0.02 * xs + 0.01 * ys).astype(np.float32)
    vz[:, 56:] = 3.4e38; vz[20:30, 50:56] = 3.4e38                                      # sky: whole tiles on the right plus a notch
    n = np.zeros((H, W, 3), np.float32); n[..., 2] = -1; n[ys > 30] = np.float32([0, 0.6, -0.8])
    rough = (xs / (W - 1)).astype(np.float32)
    nr = pack_normal_roughness(n, rough) | (np.where(xs < 20, 1, 0).astype(np.uint32) << 30)
    rb = oracle.Reblur(); port = emu.ReblurPort()
    for f in range(5):
        d = np.zeros((H, W, 4), np.float16); d[..., 0] = rng.gamma(2.0, 0.25, (H, W)); d[..., 1] = rng.normal(0, 0.05, (H, W)); d[..., 2] = rng.normal(0, 0.05, (H, W))
        d[..., 3] = np.where(rng.random((H, W)) < 0.3, 0.0, rng.random((H, W)))
        s = np.zeros((H, W, 4), np.float16); s[..., 0] = rng.gamma(1.0, 0.5, (H, W)); s[..., 1] = rng.normal(0, 0.05, (H, W)); s[..., 3] = np.where(rng.random((H, W)) < 0.3, 0.0, rng.random((H, W)))
        dd = dict(view_z=vz, normal_roughness=nr, diff=d, spec=s, motion=np.zeros((H, W, 4), np.float16), disocclusion_mix=(rng.random((H, W)) < 0.1).astype(np.uint8) * 255)
        a, b = _both(rb, port, sb, cam, cam, f, dd)
        _assert_identical(a, b, vz < 1e30, ("synthetic", f))
    rb.close(); port.close()


def test_spec_hit_t_guide_filter(oracle):
    """DenoisingGuidesBaker::DenoiseSpecHitT: oracle properties (holes filled from neighbours of similar depth only, values capped at 1.5 x + 0.5, tiny values dropped) and the
    product's pixel function compiled for the host equal to the oracle bit for bit, on synthetic guides and on the guide of a rendered frame."""
    from rtxpt_b200 import scene_builder as sb, scenes
    import host_build_lib as emu
    rng = np.random.default_rng(3)
    H, W = 37, 53
    depth = np.where(np.mgrid[0:H, 0:W][1] < 30, 5.0, 50.0).astype(np.float32) * (1 + 0.002 * rng.standard_normal((H, W)).astype(np.float32))
    hit = np.where(rng.random((H, W)) < 0.4, 0.0, rng.gamma(2.0, 3.0, (H, W))).astype(np.float32)
    hit[5, 5] = 1e-3; hit[10, 40] = 1e6; hit[20:24, 10:14] = -1.0
    out = oracle.denoise_spec_hit_t(depth, hit)
    assert np.isfinite(out).all() and (out >= 0).all()
    had = hit >= 5e-2
    assert (out[had] <= (hit[had] * 1.5 + 0.5) * 1.5 + 0.5 + 1e-3).all()                       # two passes of the cap
    assert (out[~had] > 0).mean() > 0.95                                                       # holes get a value
    near = np.zeros((H, W), np.float32); near[:, :30] = 2.0; far = np.zeros((H, W), np.float32); far[:, 30:] = 40.0
    o2 = oracle.denoise_spec_hit_t(depth, near + far)
    assert np.allclose(o2[:, :30], 2.0) and np.allclose(o2[:, 30:], 40.0)                      # nothing crosses the depth edge
    assert np.array_equal(emu.denoise_spec_hit_t(depth, hit), out) and np.array_equal(emu.denoise_spec_hit_t(depth, near + far), o2)
    # a rendered frame's guide
    Wr, Hr = 80, 64
    scene, cam = scenes.cornell_box(Wr, Hr, delta_surfaces=True)
    o = oracle.Oracle(scene); c = sb.make_constants(Wr, Hr, cam, bounce_count=4, diffuse_bounce_count=3); o.set_constants(c); o.set_view(sb.world_to_clip(cam))
    r = o.render_realtime(sb.make_realtime_constants(Wr, Hr, cam, bounce_count=4, sub_samples=1)); o.close()
    a = oracle.denoise_spec_hit_t(r["depth"], r["spec_hit_t"]); b = emu.denoise_spec_hit_t(r["depth"], r["spec_hit_t"])
    assert np.array_equal(a, b) and (a > 0).sum() >= (r["spec_hit_t"] > 0).sum()
