"""CPU: the spatial half of the ReBLUR restatement (oracle/reblur.h): ClassifyTiles, HitDistReconstruction 5x5, PrePass, Blur, PostBlur on NRD-format
inputs.  NRD ships no reference outputs and its MathLib dependency is not vendored with the reference, so parity is unpinned; these tests hold the restatement to the properties
the passes are designed to have (edge stopping by plane distance / normal / roughness, energy preservation, monotone error reduction, hit-distance reconstruction)."""
import numpy as np
import pytest


def pack_normal_roughness(n, roughness):
    """NRD_FrontEnd_PackNormalAndRoughness, R10G10B10A2_UNORM, linear roughness, material id 0 (NRD.hlsli:646-677)."""
    n = np.asarray(n, np.float32); v = n / np.abs(n).sum(-1, keepdims=True)
    wrap = (1.0 - np.abs(v[..., [1, 0]])) * np.where(v[..., :2] >= 0, 1.0, -1.0)
    xy = np.where(v[..., 2:3] >= 0, v[..., :2], wrap) * 0.5 + 0.5
    q = lambda x: (np.clip(x, 0, 1) * 1023.0 + 0.5).astype(np.uint32)
    return q(xy[..., 0]) | (q(xy[..., 1]) << 10) | (q(np.broadcast_to(np.asarray(roughness, np.float32), xy.shape[:-1])) << 20)


def camera(W, H):
    from rtxpt_b200 import scene_builder as sb
    cam = sb.bridge_camera(W, H, pos=(0, 0, 0), direction=(0, 0, 1), up=(0, 1, 0), fov_y=0.9)
    return sb.world_to_view(cam), sb.view_to_clip(cam)


def wall(W, H, z=5.0, normal=(0, 0, -1), roughness=1.0):
    return np.full((H, W), z, np.float32), pack_normal_roughness(np.broadcast_to(np.float32(normal), (H, W, 3)), roughness)


def test_constant_signal_and_sky_tiles(oracle):
    W, H = 96, 64
    wv, vc = camera(W, H)
    vz, nr = wall(W, H)
    vz[:, 64:] = 3.4e38                                              # right third: sky
    d = np.zeros((H, W, 4), np.float16); d[..., 0] = 0.5; d[..., 1] = 0.1; d[..., 2] = -0.05; d[..., 3] = 0.25
    s = d.copy()
    od, os_, _, tiles = oracle.reblur_spatial(wv, vc, 7, vz, nr, d, s)
    assert tiles[:, :4].sum() == 0 and (tiles[:, 4:] == 1).all()                                         # 16x16 tiles entirely beyond the denoising range
    assert np.array_equal(od[:, :64], d[:, :64]) and np.array_equal(os_[:, :64], s[:, :64])              # a constant is a fixed point of every pass (weights normalise)
    assert np.array_equal(od[:, 64:], d[:, 64:])                                                         # sky pixels are never written: the harness passes its input through


def test_edges_stop_the_blur_and_energy_is_kept(oracle):
    """Two walls meeting at a crease with different radiance: nothing leaks across (plane-distance and normal weights), while noise inside each wall goes down."""
    W, H = 128, 64
    wv, vc = camera(W, H)
    rng = np.random.default_rng(2)
    ys, xs = np.mgrid[0:H, 0:W]
    left = xs < W // 2
    # left wall faces the camera at z = 6; right wall is turned by 60 degrees and recedes
    u = (xs + 0.5) / W * 2 - 1
    tanx = np.tan(0.45) * W / H
    nR = np.float32([-np.sin(np.pi / 3), 0, -np.cos(np.pi / 3)])
    vz = np.where(left, 6.0, 6.0 * nR[2] / (nR[0] * u * tanx + nR[2])).astype(np.float32)               # plane through (0, 0, 6) with normal nR
    n = np.where(left[..., None], np.float32([0, 0, -1]), nR)
    nr = pack_normal_roughness(n, 1.0)
    base = np.where(left, 1.0, 0.1).astype(np.float32)
    noisy = base * rng.gamma(4.0, 0.25, (H, W)).astype(np.float32)                                        # mean-preserving multiplicative noise
    d = np.zeros((H, W, 4), np.float16); d[..., 0] = noisy; d[..., 3] = 0.5
    s = d.copy()
    acc = np.full((H, W, 2), 0.0, np.float32)
    err_before = np.abs(noisy - base)
    prev_d = prev_s = None
    for stages in (0b0010, 0b0110, 0b1110):
        od, os_, _, _ = oracle.reblur_spatial(wv, vc, 1, vz, nr, d, s, acc, stages)
        for out in (od, os_):
            y = out[..., 0].astype(np.float32)
            assert abs(y[left].mean() - noisy[left].mean()) < 0.03 and abs(y[~left].mean() - noisy[~left].mean()) < 0.01
            assert np.abs(y - base)[left].mean() < err_before[left].mean() and np.abs(y - base)[~left].mean() < err_before[~left].mean()
            assert y[~left].max() < 0.5 and y[left].min() > 0.15                                          # no leak across the crease in either direction
        if prev_d is not None:
            assert np.abs(od[..., 0].astype(np.float32) - base).mean() <= np.abs(prev_d - base).mean() * 1.02
        prev_d = od[..., 0].astype(np.float32)


def test_roughness_and_hit_distance_weights(oracle):
    """Specular only mixes samples of similar roughness; a mirror-like pixel (roughness 0) keeps its value (lobe too narrow for any neighbour)."""
    W, H = 96, 48
    wv, vc = camera(W, H)
    vz, _ = wall(W, H)
    rough = np.where(np.mgrid[0:H, 0:W][1] < W // 2, 0.0, 0.8).astype(np.float32)
    nr = pack_normal_roughness(np.broadcast_to(np.float32([0, 0, -1]), (H, W, 3)), rough)
    rng = np.random.default_rng(5)
    s = np.zeros((H, W, 4), np.float16); s[..., 0] = rng.random((H, W)).astype(np.float32); s[..., 3] = 0.4
    d = s.copy()
    od, os_, track, _ = oracle.reblur_spatial(wv, vc, 2, vz, nr, d, s, np.zeros((H, W, 2), np.float32), 0b1110)
    y_in, y_out = s[..., 0].astype(np.float32), os_[..., 0].astype(np.float32)
    mirror = rough == 0
    assert np.abs(y_out - y_in)[mirror].max() < 0.05                                                      # smc( 0 ) = 0: minimum radius 0, area factor 0
    assert y_out[~mirror].std() < 0.8 * y_in[~mirror].std()
    assert (track[~mirror] > 0).all()                                                                     # hit distance for tracking = the closest credible hit in the kernel (metres)


def test_hit_distance_reconstruction(oracle):
    """Rays that missed (normalised hit distance 0) take a weighted average of the valid hit distances on the same surface within 5x5; radiance is untouched."""
    W, H = 64, 48
    wv, vc = camera(W, H)
    vz, nr = wall(W, H, roughness=0.5)
    rng = np.random.default_rng(9)
    d = np.zeros((H, W, 4), np.float16); d[..., :3] = rng.random((H, W, 3)); d[..., 3] = 0.3
    holes = rng.random((H, W)) < 0.4
    d[..., 3][holes] = 0
    s = d.copy()
    od, os_, _, _ = oracle.reblur_spatial(wv, vc, 0, vz, nr, d, s, None, 0b0001)
    assert np.array_equal(od[..., :3], d[..., :3]) and np.array_equal(os_[..., :3], s[..., :3])
    assert np.allclose(od[..., 3].astype(np.float32), 0.3, atol=2e-3) and np.allclose(os_[..., 3].astype(np.float32), 0.3, atol=2e-3)
    # a foreground object in front of the wall does not lend its hit distances to the wall behind it
    vz2 = vz.copy(); vz2[:, :32] = 2.0
    d2 = d.copy(); d2[..., 3] = np.where(np.mgrid[0:H, 0:W][1] < 32, 0.9, 0.0).astype(np.float16)
    od2, _, _, _ = oracle.reblur_spatial(wv, vc, 0, vz2, nr, d2, d2.copy(), None, 0b0001)
    assert (od2[..., 3][:, 34:] == 0).all() and np.allclose(od2[..., 3][:, :32].astype(np.float32), 0.9, atol=2e-3)


def _cornell_frames(oracle, W=96, H=96, bounces=4):
    from rtxpt_b200 import scene_builder as sb, scenes
    scene, cam = scenes.cornell_box(W, H)
    o = oracle.Oracle(scene)
    consts = sb.make_constants(W, H, cam, bounce_count=bounces, diffuse_bounce_count=3)

    def frame(index, camera=cam, prev_camera=None):
        c = sb.make_constants(W, H, camera, bounce_count=bounces, diffuse_bounce_count=3); c.sampleBaseIndex = index
        o.set_constants(c); o.set_view(sb.world_to_clip(camera))
        rt = sb.make_realtime_constants(W, H, camera, prev_cam=prev_camera, bounce_count=bounces, sub_samples=1)
        r = o.render_realtime(rt); d = o.new_denoiser_targets()
        o.denoiser_prepare_inputs(rt, sb.make_denoiser_constants(camera), r, d, 0, True)
        return d
    return o, cam, frame


def test_full_chain_converges_on_a_static_view(oracle):
    """Eight passes per frame with history: the accumulated-frame counters grow by one per frame, the error against a converged image of the same demodulated signal keeps
    falling well below what one frame of spatial filtering reaches, and the mean energy stays put."""
    from rtxpt_b200 import scene_builder as sb
    W = H = 96
    o, cam, frame = _cornell_frames(oracle, W, H)
    ref_d = np.zeros((H, W, 4)); ref_s = np.zeros((H, W, 4)); N = 48
    for f in range(N):
        d = frame(1000 + f); ref_d += d["diff"]; ref_s += d["spec"]
    ref_d /= N; ref_s /= N
    rb = oracle.Reblur(); wv, vc = sb.world_to_view(cam), sb.view_to_clip(cam)
    errs = []
    for f in range(16):
        d = frame(f)
        od, os_, frames = rb.denoise(wv, vc, f, d["view_z"], d["normal_roughness"], d["diff"], d["spec"], motion=d["motion"], disocclusion_mix=d["disocclusion_mix"])
        surf = d["view_z"] < 1e30
        assert np.isfinite(od.astype(np.float32)).all() and np.isfinite(os_.astype(np.float32)).all()
        assert abs(frames[surf].mean() - f) < 0.08 * f + 0.01                      # one more frame of history everywhere: nothing moved
        errs.append((np.abs(od[..., 0].astype(np.float32) - ref_d[..., 0])[surf].mean(), np.abs(os_[..., 0].astype(np.float32) - ref_s[..., 0])[surf].mean(),
                     np.abs(d["diff"][..., 0].astype(np.float32) - ref_d[..., 0])[surf].mean(), np.abs(d["spec"][..., 0].astype(np.float32) - ref_s[..., 0])[surf].mean()))
    e = np.float32(errs)
    assert (e[:, 0] < e[:, 2]).all() and (e[:, 1] < e[:, 3]).all()                 # better than the input on every frame
    assert e[15, 0] < 0.55 * e[0, 0] and e[15, 1] < 0.75 * e[0, 1] and e[15, 0] < 0.35 * e[15, 2]
    assert abs(od[..., 0][surf].astype(np.float32).mean() - ref_d[..., 0][surf].mean()) < 0.02 * ref_d[..., 0][surf].mean()
    rb.close(); o.close()


def test_history_follows_a_moving_camera_and_resets_on_a_cut(oracle):
    """A slow dolly: surface-motion reprojection (motion vectors exported with the stable planes, previous view matrices) keeps most of the history; a camera cut
    far away fails the disocclusion tests and the counters restart, as they do when the caller asks for a reset."""
    from rtxpt_b200 import scene_builder as sb
    W = H = 96
    o, cam0, frame = _cornell_frames(oracle, W, H)
    def cam_at(dx, dz=0.0):
        return sb.bridge_camera(W, H, pos=(2.78 + dx, 2.73, -8.0 + dz), direction=(0, 0, 1), up=(0, 1, 0), fov_y=0.66)
    rb = oracle.Reblur(); prev = None
    for f in range(10):
        cam = cam_at(0.02 * f)
        d = frame(f, cam, prev)
        od, os_, frames = rb.denoise(sb.world_to_view(cam), sb.view_to_clip(cam), f, d["view_z"], d["normal_roughness"], d["diff"], d["spec"],
                                     prev_world_to_view=None if prev is None else sb.world_to_view(prev), prev_view_to_clip=None if prev is None else sb.view_to_clip(prev),
                                     motion=d["motion"], disocclusion_mix=d["disocclusion_mix"])
        prev = cam
    surf = d["view_z"] < 1e30
    assert (np.abs(d["motion"][..., 0].astype(np.float32))[surf] > 0).mean() > 0.9                           # the dolly produces screen-space motion
    assert np.median(frames[..., 0][surf]) > 6.0 and (frames[..., 0][surf] > 4.0).mean() > 0.8              # history survives the motion
    # a cut the application does not describe (no motion vectors, previous view = current view): reprojection lands on other surfaces, the plane-distance /
    # normal tests reject them and the counters restart for most of the image
    cut = sb.bridge_camera(W, H, pos=(4.2, 1.2, -6.0), direction=(-0.25, 0.1, 1), up=(0, 1, 0), fov_y=0.66)
    d = frame(20, cut, cut)
    _, _, frames_cut = rb.denoise(sb.world_to_view(cut), sb.view_to_clip(cut), 10, d["view_z"], d["normal_roughness"], d["diff"], d["spec"], motion=None, disocclusion_mix=d["disocclusion_mix"])
    surf = d["view_z"] < 1e30
    assert np.median(frames_cut[..., 0][surf]) < 3.0 and np.median(frames[..., 0]) > 2 * np.median(frames_cut[..., 0][surf])
    d = frame(21, cut, cut)
    _, _, frames_reset = rb.denoise(sb.world_to_view(cut), sb.view_to_clip(cut), 11, d["view_z"], d["normal_roughness"], d["diff"], d["spec"], motion=d["motion"], reset=True)
    assert (frames_reset == 0).all()
    rb.close(); o.close()


def test_whole_denoised_frame_beats_the_noisy_one(oracle):
    """Sample::Denoise end to end in the oracle, as rtxpt_b200_denoise_realtime composes it: realtime frame -> specular hit distance guide filter -> for plane = 2..0 { prepare inputs,
    that plane's ReBLUR instance, final merge }.  After a few frames the merged colour is closer to a converged reference-mode image than the no-denoiser merge, at equal energy."""
    from rtxpt_b200 import scene_builder as sb, scenes
    W, H = 80, 64
    scene, cam = scenes.cornell_box(W, H, delta_surfaces=True)
    o = oracle.Oracle(scene)
    c = sb.make_constants(W, H, cam, bounce_count=5, diffuse_bounce_count=3); o.set_constants(c); o.set_view(sb.world_to_clip(cam))
    acc, n = None, 0
    for base in range(0, 192, 8):
        c.sampleBaseIndex = 5000 + base; o.set_constants(c); acc, n = o.render(0, 8, accum=acc, accum_count=n)[:2]
    ref = acc[..., :3]
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=5, sub_samples=1); k = sb.make_denoiser_constants(cam)
    wv, vc = sb.world_to_view(cam), sb.view_to_clip(cam)
    rbs = [oracle.Reblur() for _ in range(3)]
    for f in range(10):
        c.sampleBaseIndex = f; o.set_constants(c)
        r = o.render_realtime(rt)
        r["spec_hit_t"] = oracle.denoise_spec_hit_t(r["depth"], r["spec_hit_t"])
        d = o.new_denoiser_targets()
        for i, plane in enumerate((2, 1, 0)):
            o.denoiser_prepare_inputs(rt, k, r, d, plane, i == 0)
            od, os_, _ = rbs[plane].denoise(wv, vc, f, d["view_z"], d["normal_roughness"], d["diff"], d["spec"], motion=d["motion"], disocclusion_mix=d["disocclusion_mix"])
            o.denoiser_final_merge(rt, r, d, plane, od, os_)
    den = d["output"][..., :3].astype(np.float32); noisy = r["merged"]
    assert np.isfinite(den).all()
    clip = lambda x: np.minimum(x, 4.0)
    e_den, e_noisy = np.abs(clip(den) - clip(ref)).mean(), np.abs(clip(noisy) - clip(ref)).mean()
    assert e_den < 0.8 * e_noisy, (e_den, e_noisy)                           # the reference image itself still carries noise: the ratio understates the gain
    assert abs(den.mean() - ref.mean()) < 0.12 * ref.mean(), (den.mean(), ref.mean())
    for rb in rbs: rb.close()
    o.close()
