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
