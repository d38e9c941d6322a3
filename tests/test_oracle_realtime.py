"""CPU: realtime mode of the oracle - path-space decomposition into stable planes: the BUILD pass (delta-only exploration),
the FILL pass (noisy radiance deposited per plane) and the no-denoiser merge.  Branch-ID arithmetic, the StablePlane layout and the GenericTS
addressing are pinned against vectors produced by the reference's own C++ halves of StablePlanes.hlsli / Utils.hlsli
(tests/golden/host_golden.json); the passes themselves are checked through their invariants."""
import ctypes as C
import json
import os
import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "host_golden.json")))
INVALID = 0xFFFFFFFF


def test_stable_plane_layout_and_constants_match_reference_headers():
    from rtxpt_b200 import structs as S
    g = GOLD["StablePlane"]
    assert C.sizeof(S.StablePlane) == g["size"] == np.dtype(S.STABLE_PLANE_DTYPE).itemsize == 80
    for field, off in g["offsets"].items():
        assert getattr(S.StablePlane, field).offset == off and np.dtype(S.STABLE_PLANE_DTYPE).fields[field][1] == off, field
    k = GOLD["stable_plane_constants"]
    assert (k["cStablePlaneCount"], k["cStablePlaneInvalidBranchID"]) == (S.STABLE_PLANE_COUNT, S.STABLE_PLANE_INVALID_BRANCH)
    assert (k["cStablePlaneMaxVertexIndex"], k["cStablePlaneEnqueuedBranchID"], k["cStablePlaneJustStartedID"], k["cMaxDeltaLobes"]) == (15, 0xFFFFFFFE, 0, 3)


def test_branch_ids_match_reference(oracle):
    L = oracle.lib()
    for a, b, va, vb, on_path, on_plane, parent in GOLD["branch_ids"]:
        assert L.oracle_branch_vertex_index(a) == va and L.oracle_branch_vertex_index(b) == vb
        assert L.oracle_branch_on_stable_path(a, va, b, vb) == on_path            # StablePlaneIsOnStablePath(plane, vertex): is `b` a prefix of `a`
        assert int(a == b) == on_plane and (a & 3) == parent
        # the id is the camera's 1 followed by two bits per delta lobe taken
        assert L.oracle_branch_advance(a >> 2, a & 3) == a or va == 1
    assert any(r[4] for r in GOLD["branch_ids"]) and not all(r[4] for r in GOLD["branch_ids"])


def test_generic_ts_addressing_matches_reference(oracle, product):
    from rtxpt_b200 import scene_builder as sb
    L = oracle.lib(); P = product.load()
    for f in ("rtxpt_b200_generic_ts_line_stride", "rtxpt_b200_generic_ts_plane_stride", "rtxpt_b200_generic_ts_address"):
        getattr(P, f).restype = C.c_uint32; getattr(P, f).argtypes = [C.c_uint32] * (5 if f.endswith("address") else 2)
    for g in GOLD["generic_ts"]:
        w, h = g["size"]
        assert L.oracle_generic_ts_line_stride(w, h) == P.rtxpt_b200_generic_ts_line_stride(w, h) == g["line"]
        assert L.oracle_generic_ts_plane_stride(w, h) == P.rtxpt_b200_generic_ts_plane_stride(w, h) == g["plane"] and g["count3"] == 3 * g["plane"]
        for x, y, p, addr in g["samples"]:
            assert L.oracle_generic_ts_address(x, y, p, g["line"], g["plane"]) == addr == P.rtxpt_b200_generic_ts_address(x, y, p, g["line"], g["plane"])
            assert int(sb.generic_ts_address(x, y, p, w, h)) == addr
    # a bijection of the padded image onto [0, planeStride)
    ys, xs = np.mgrid[0:40, 0:72]
    a = sb.generic_ts_address(xs, ys, 0, 72, 40)
    assert len(np.unique(a)) == a.size and a.max() == 72 * 40 - 1


def test_ortho_matrix_packing_round_trip(oracle):
    """PackOrthoMatrix / UnpackOrthoMatrix (Utils.hlsli:171-189): two octahedral 15+15-bit rows and the handedness bit."""
    L = oracle.lib(); rng = np.random.default_rng(5)
    for i in range(64):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if i % 2: q[2] = -q[2]                                            # both handednesses
        m = np.ascontiguousarray(q, np.float32); packed = np.zeros(2, np.uint32); back = np.zeros(9, np.float32)
        L.oracle_pack_ortho(m.ctypes.data, packed.ctypes.data); L.oracle_unpack_ortho(packed.ctypes.data, back.ctypes.data)
        assert (packed[1] >> 31) == (1 if np.linalg.det(q) > 0 else 0)
        assert np.abs(back.reshape(3, 3) - q).max() < 2e-4


def _unpack_pairs(words):
    hi = (words >> 16).astype(np.uint16).view(np.float16).astype(np.float32); lo = (words & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
    return hi, lo


@pytest.fixture(scope="module")
def mirror_glass(oracle):
    from rtxpt_b200 import scene_builder as sb, scenes
    W = H = 80
    scene, cam = scenes.cornell_box(W, H, delta_surfaces=True)
    o = oracle.Oracle(scene)
    consts = sb.make_constants(W, H, cam, bounce_count=8, diffuse_bounce_count=3)
    o.set_constants(consts); o.set_view(sb.world_to_clip(cam))
    yield o, consts, cam, W, H
    o.close()


def test_build_pass_structure(mirror_glass):
    """What the BUILD pass must leave behind: plane 0 always valid; the mirror is replaced by what it shows (primary surface replacement, branch
    1 -> reflection lobe 1 = 0b101); glass forks reflection and transmission onto separate planes; dominance follows the material's choice."""
    from rtxpt_b200 import scene_builder as sb
    o, consts, cam, W, H = mirror_glass
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=1)
    r = o.render_realtime(rt)
    hd = r["header"]; ys, xs = np.mgrid[0:H, 0:W]
    p0 = r["planes"][sb.generic_ts_address(xs, ys, 0, W, H)]
    assert (hd[0] != INVALID).all() and (hd[0] != 0xFFFFFFFE).all() and (hd[:3] != 0).all()                   # nothing left enqueued / just started
    ids0 = set(np.unique(hd[0]).tolist()); assert 1 in ids0 and 0b101 in ids0
    vertex = p0["VertexIndexAndRoughness"] >> 16
    assert ((vertex == 1) == (hd[0] == 1)).all() and (vertex[hd[0] == 0b101] == 2).all()                        # vertex index = 1 + number of lobes taken
    split = hd[1] != INVALID
    assert 0.02 < split.mean() < 0.5 and ((hd[2] != INVALID) <= split).all()
    # glass (dominant lobe 0 = transmission): wherever both forks exist, the transmission plane (id ...00) is the dominant one
    both = split & (hd[2] != INVALID)
    dom = hd[3] & 3
    def first_lobe(ids):                                                    # the lobe taken at the first fork: the two bits below the camera's leading 1
        ids = ids.astype(np.uint64); v = (np.floor(np.log2(ids)).astype(np.uint64) // 2) + 1
        return (ids >> (2 * (v - 2))) & 3
    direct = both & (hd[0] == 1)                                            # glass seen directly (not in the mirror): plane 0 stops on it, the two lobes fork
    assert direct.any() and (first_lobe(hd[1][direct]) == 0).all() and (first_lobe(hd[2][direct]) == 1).all() and (dom[both] == 1).all() and (dom[~split] == 0).all()
    # first-hit ray length in the upper 30 bits of layer 3 equals the primary hit distance (fp32 with the 2 low bits cleared)
    prim = o.render(0, 1, want_primary=True)[3]
    hit = prim[..., 0] > 0
    first = (hd[3] & 0xFFFFFFFC).view(np.float32)
    assert np.allclose(first[hit], prim[..., 0][hit], rtol=1e-6) and (first[~hit] >= 9.9e14).all()
    # sky planes: infinite scene length, never dominant-surface guides; surfaces: finite, positive
    sl = p0["SceneLength"]; assert np.isfinite(sl[hit & (vertex == 1)]).all() and (sl[np.isfinite(sl)] > 0).all()
    # guides come from the dominant plane: depth in (0,1) where anything was hit, throughput of the mirror path tinted by the mirror
    assert ((r["depth"] > 0) & (r["depth"] <= 1)).all() and (r["depth"][hit & (hd[0] == 1) & ~split] < 1).all()
    thp_hi, mv_lo = _unpack_pairs(p0["PackedThpAndMVs"])
    assert np.allclose(thp_hi[hd[0] == 1], 1.0) and (thp_hi[hd[0] == 0b101] <= 1.0).all() and np.isclose(thp_hi[hd[0] == 0b101], np.float32([0.95, 0.93, 0.88]), atol=2e-3).all(-1).mean() > 0.5 and (mv_lo[..., :2] == 0).all()     # static camera: zero screen motion
    # stable radiance: the lamp seen directly or through the delta tree, nothing else
    sr = r["stable_radiance"][..., :3].astype(np.float32)
    assert sr.max() >= 17.0 and (sr > 0).any(-1).mean() < 0.2
    # demodulation guides are clamped to [0.04, 65504]
    d_hi, s_lo = _unpack_pairs(p0["DenoiserPackedBSDFEstimate"]); assert d_hi.min() >= 0.04 - 1e-3 and s_lo.min() >= 0.04 - 1e-3


def test_fill_pass_converges_to_reference_mode(mirror_glass):
    """stable radiance + the planes' noisy radiance is an unbiased estimate of the same image reference mode renders."""
    from rtxpt_b200 import scene_builder as sb
    o, consts, cam, W, H = mirror_glass
    frames, spp = 6, 4
    acc = np.zeros((H, W, 3), np.float64)
    for f in range(frames):
        consts.sampleBaseIndex = f * spp; o.set_constants(consts)
        r = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=spp))
        acc += r["merged"]
    acc /= frames
    consts.sampleBaseIndex = 0; o.set_constants(consts)
    ref = o.render(0, 48)[0][..., :3]
    assert abs(acc.mean() - ref.mean()) < 0.03 * ref.mean(), (acc.mean(), ref.mean())
    blocks = lambda a: a.reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3))
    assert np.abs(blocks(acc) - blocks(ref)).mean() < 0.08 * ref.mean()
    # determinism
    r2 = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=spp))
    consts.sampleBaseIndex = (frames - 1) * spp; o.set_constants(consts)
    r3 = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=spp))
    assert np.array_equal(r["merged"], r3["merged"]) and r["planes"].tobytes() == r3["planes"].tobytes() and not np.array_equal(r2["merged"], r3["merged"])
    consts.sampleBaseIndex = 0; o.set_constants(consts)


def test_sub_sample_attenuation_and_plane_count(mirror_glass):
    """The noisy radiance of N sub-samples is attenuated by 1/N as it is deposited (so N samples average); with one active plane nothing forks."""
    from rtxpt_b200 import scene_builder as sb
    o, consts, cam, W, H = mirror_glass
    r1 = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=1))
    r4 = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=4))
    assert np.array_equal(r1["header"], r4["header"]) and np.array_equal(r1["stable_radiance"], r4["stable_radiance"])        # the BUILD pass does not depend on the sample count
    assert abs(r1["merged"].mean() - r4["merged"].mean()) < 0.1 * r1["merged"].mean()
    one = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=1, active_planes=1))
    assert (one["header"][1:3] == INVALID).all() and (one["header"][3] & 3 == 0).all()
    no_psr = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=1, allow_psr=False))
    assert 0b101 not in set(np.unique(no_psr["header"][0]).tolist()) and 0b101 in set(np.unique(no_psr["header"][1:3]).tolist())   # the mirror now forks instead of replacing the primary surface


def test_denoiser_interface_round_trip(mirror_glass):
    """RTXPT's side of the denoiser interface (row a18): prepare inputs -> (identity denoiser) -> final merge over the planes, last plane first, reproduces the
    no-denoiser merge; the prepared inputs obey NRD's input contract (viewZ positive or the sky marker, unit normals, YCoCg radiance, normalised hit distance)."""
    from rtxpt_b200 import scene_builder as sb
    o, consts, cam, W, H = mirror_glass
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=4)
    r = o.render_realtime(rt)
    k = sb.make_denoiser_constants(cam)
    d = o.new_denoiser_targets()
    seen_surface = np.zeros((H, W), bool)
    for i, plane in enumerate((2, 1, 0)):
        o.denoiser_prepare_inputs(rt, k, r, d, plane, i == 0)
        surf = d["view_z"] < 1e30
        valid = r["header"][plane] != INVALID
        ys, xs = np.mgrid[0:H, 0:W]
        finite = np.isfinite(r["planes"][sb.generic_ts_address(xs, ys, plane, W, H)]["SceneLength"])
        assert np.array_equal(surf, valid & finite)                                      # sky planes and unused planes are marked, everything else is a surface
        assert (d["view_z"][surf] > 0).all() and (d["view_z"][~surf] == np.float32(3.402823466e+38)).all()
        nr = d["normal_roughness"][surf]
        ex, ey, rough = (nr & 1023) / 1023.0, ((nr >> 10) & 1023) / 1023.0, ((nr >> 20) & 1023) / 1023.0
        assert (nr >> 30 == 0).all() and (rough >= 0.2 - 1e-3).all()                      # material id 0; kMinRoughness
        px, py = ex * 2 - 1, ey * 2 - 1; n = np.stack([px, py, 1 - np.abs(px) - np.abs(py)], -1); t = np.clip(-n[..., 2], 0, 1)
        n[..., 0] -= t * np.where(n[..., 0] >= 0, 1, -1); n[..., 1] -= t * np.where(n[..., 1] >= 0, 1, -1); n /= np.linalg.norm(n, axis=-1, keepdims=True)
        assert np.isfinite(n).all()
        diff, spec = d["diff"][surf].astype(np.float32), d["spec"][surf].astype(np.float32)
        assert (diff[..., 0] >= 0).all() and (spec[..., 0] >= 0).all() and (diff[..., 3] == 0).all() and ((spec[..., 3] >= 0) & (spec[..., 3] <= 1)).all()      # Y >= 0, hit distance in [0, 1]
        lum_max = 128.0 + 1                                                             # min(255, preExposedGrayLuminance * clampK * 16)
        assert diff[..., 0].max() <= lum_max * 1.5 and spec[..., 0].max() <= lum_max * 1.5
        o.denoiser_final_merge(rt, r, d, plane, d["diff"].copy(), d["spec"].copy())
        seen_surface |= surf
    out = d["output"][..., :3].astype(np.float32)
    assert (d["output"][..., 3] == 1).all()
    assert np.isclose(out, r["merged"], rtol=2e-2, atol=4e-3).all(-1).mean() > 0.999 and abs(out.mean() - r["merged"].mean()) < 1e-3 * r["merged"].mean()
    # motion vectors are the plane's own; the disocclusion relaxation only exists behind a delta bounce (vertex index > 1)
    assert (d["motion"] == 0).all()                                                     # static camera


def _opt_glass_into_decomposition(scene):
    """What a .material.json with PSDExclude = false, PSDDominantDeltaLobe = 0 does for clear glass (roughness below the delta threshold)."""
    from rtxpt_b200 import structs as S
    n = 0
    for i in range(scene.desc.materialCount):
        m = scene.desc.materials[i]
        if m.TransmissionFactor > 0 and m.Roughness * m.Roughness < 0.0064:
            m.Flags = (m.Flags & ~S.MATFLAG_PSDExclude & ~0x0F000000) | (1 << 24); n += 1
    return n


def test_realtime_on_textured_env_lit_scene(oracle):
    """Realtime mode on the city stand-in (textures, environment map + NEE-AT proxies, alpha-tested foliage, thin and solid glass): same expectation as
    reference mode, the specular hit distance is exported where the dominant plane scatters off glossy surfaces, sky planes carry no noisy radiance."""
    from rtxpt_b200 import scene_builder as sb, scenes
    W, H = 160, 90
    scene, cam = scenes.city_block(target_triangles=120000, width=W, height=H, texture_size=128, n_textures=6, n_materials=64)
    assert _opt_glass_into_decomposition(scene) > 0
    o = oracle.Oracle(scene)
    consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=6, env_enabled=True, firefly_threshold=5000.0, nee=True, nee_type=2)
    o.set_view(sb.world_to_clip(cam))
    acc = np.zeros((H, W, 3)); frames = 6
    for f in range(frames):
        consts.sampleBaseIndex = f * 4; o.set_constants(consts)
        r = o.render_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=4)); acc += r["merged"]
    acc /= frames
    consts.sampleBaseIndex = 0; o.set_constants(consts)
    ref = o.render(0, 48)[0][..., :3]
    assert np.isfinite(acc).all() and abs(acc.mean() - ref.mean()) < 0.03 * ref.mean(), (acc.mean(), ref.mean())
    hd = r["header"]
    assert (hd[0] != INVALID).all() and 0 < (hd[1] != INVALID).mean() < 0.2
    # a path that ends at a hit while the distance is still being measured leaves the negative "started" marker behind, as in the reference (the
    # denoiser front end saturates it to 0); most measurements complete
    assert 0.02 < (r["spec_hit_t"] > 0).mean() < 0.9 and (r["spec_hit_t"] < 0).mean() < 0.25 * (r["spec_hit_t"] > 0).mean()
    ys, xs = np.mgrid[0:H, 0:W]
    p0 = r["planes"][sb.generic_ts_address(xs, ys, 0, W, H)]
    sky = ~np.isfinite(p0["SceneLength"])
    assert sky.any() and (p0["PackedNoisyRadianceAndSpecAvg"][sky] == 0).all()
    assert (r["stable_radiance"][..., :3][sky & (hd[1] == INVALID)] > 0).any()                     # the sky seen directly is stable radiance
    o.close()
