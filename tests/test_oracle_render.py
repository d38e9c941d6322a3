"""CPU: the oracle path tracer on the Cornell box (BASELINE.json configs[0], "reference-mode CPU shading (no GPU)")."""
import numpy as np


def test_cornell_c1_cpu(oracle, cornell):
    from rtxpt_b200 import scene_builder as sb
    scene, cam = cornell
    o = oracle.Oracle(scene)
    consts = sb.make_constants(256, 256, cam, bounce_count=2, diffuse_bounce_count=2)
    o.set_constants(consts)
    acc, n, last, prim, st = o.render(0, 1, want_primary=True)
    assert n == 1 and np.isfinite(acc).all() and (acc[..., :3] >= 0).all() and (acc[..., 3] == 1).all()
    assert np.array_equal(acc[..., :3], last)                              # first sample overwrites (blend >= 1)
    # the box is open towards the camera: centre rays hit it, a border of rays passes outside; path length <= bounceCount + 1 scatter segments
    assert (prim[64:192, 64:192, 0] > 0).all() and (prim[..., 0] > 0).mean() > 0.6
    assert 256 * 256 <= st.scatterRays <= 256 * 256 * 3 and st.shadowRays <= 256 * 256 * 2
    # the light (17,12,4) is visible directly
    assert 17.0 <= acc[..., 0].max() < 19.0        # emission 17 (fp16-exact) plus what the lamp surface reflects
    # determinism + running mean: two more samples, replayed
    a2, n2 = o.render(1, 2, accum=acc.copy(), accum_count=1)[:2]
    b2, _ = o.render(1, 2, accum=acc.copy(), accum_count=1)[:2]
    assert n2 == 3 and np.array_equal(a2, b2)
    # red wall on the left of the image? camera looks down +z with +y up: world x=0 (red) appears on the RIGHT
    left, right = acc[96:160, 8:40, :3].mean((0, 1)), acc[96:160, 216:248, :3].mean((0, 1))
    assert right[0] > 2 * right[1] and left[1] > 2 * left[0]
    o.close()


def test_furnace_nee_and_bsdf_sampling_agree(oracle):
    """White diffuse plane under a uniform white sky: next-event estimation + MIS must give the same expectation as pure BSDF sampling, and the
    albedo of the roughness-1 Frostbite lobe (energyFactor 1/1.51, BxDF.hlsli:195-206) plus the F0 = 0.04 specular stays below 1."""
    from rtxpt_b200.scene_builder import SceneBuilder, Material, bridge_camera, make_constants
    from rtxpt_b200.scenes import _quad
    b = SceneBuilder(); m = b.add_material(Material(base_color=(1, 1, 1), roughness=1.0))
    b.add_mesh([_quad((-50, 0, -50), (-50, 0, 50), (50, 0, 50), (50, 0, -50), m)]); b.add_instance(0)
    b.set_env_cube(np.ones((6, 16, 16, 4), np.float32))
    scene = b.build()
    cam = bridge_camera(64, 64, pos=(0, 5, 0), direction=(0, -1, 0.001), up=(0, 0, 1), fov_y=0.5)
    o = oracle.Oracle(scene)
    means = []
    for nee in (False, True):
        o.set_constants(make_constants(64, 64, cam, bounce_count=4, diffuse_bounce_count=4, env_enabled=True, nee=nee))
        acc, n, _, _, st = o.render(0, 32)
        means.append(float(acc[..., :3].mean()))
        assert (st.shadowRays > 0) == nee
    assert 0.6 < means[0] < 1.0
    assert abs(means[0] - means[1]) < 0.01 * means[0], means
    o.close()


def test_analytic_lights_cpu(oracle):
    """Sphere / spot / zero-radius point lights: list layout [env nodes | analytic | triangles], packed records, and their effect."""
    from rtxpt_b200 import scene_builder as sb, scenes
    scene, cam = scenes.cornell_box(96, 96, analytic_lights=True)
    plain, _ = scenes.cornell_box(96, 96)
    o = oracle.Oracle(scene); consts = sb.make_constants(96, 96, cam, bounce_count=2, diffuse_bounce_count=2); o.set_constants(consts)
    infos, counters, proxies = o.lights(); ex = o.lights_ex()
    assert infos.shape[0] == 5368 + 3 + 2 and ex.shape == (3, 4)
    types = (infos[5368:5373, 3] >> 24) & 0xf
    assert list(types) == [0, 0, 4, 1, 1]                                   # sphere, sphere(spot), point, 2 emissive triangles
    assert (infos[5369, 3] >> 28) & 1 == 1 and (infos[5368, 3] >> 28) & 1 == 0  # shaping bit only on the spot
    assert counters[5368] > 0 and counters[5369] > 0 and counters[5370] == 0      # kPoint records carry no weight in the reference either
    # sphere radius packed with the truncating half conversion: 0.22 -> 0x330A (RNE would give 0x330B)
    assert infos[5368, 6] & 0xffff == 0x330A
    # spot axis decodes back to the authored direction within the 16-bit octahedral grid; cone = cos(34 deg), softness = 1 - 18/34
    half = lambda h: np.frombuffer(np.uint16(h).tobytes(), np.float16)[0]
    assert abs(half(ex[1, 2] & 0xffff) - np.cos(np.radians(34.0))) < 1e-3 and abs(half(ex[1, 2] >> 16) - (1 - 18.0 / 34.0)) < 1e-3
    a, _, _, _, st = o.render(0, 8); o.close()
    o2 = oracle.Oracle(plain); o2.set_constants(consts); b, _, _, _, st2 = o2.render(0, 8); o2.close()
    assert np.isfinite(a).all() and st.shadowRays > 0
    # the blue sphere light at (1.2, 3.9, 1.6) sits near the red wall (world x = 0 is the right side of the image): it adds blue there
    assert a[20:60, 60:90, 2].mean() > b[20:60, 60:90, 2].mean() + 0.02
    assert a[..., :3].mean() > b[..., :3].mean() * 1.05
    # the proxy cube of the sphere light shows the analytic sphere's radiance to camera rays: 14 * (0.4, 0.6, 1.0) / (pi 0.22^2) through the 8-bit + log packing
    peak = a[..., 2].max()
    assert abs(peak - 14.0 / (np.pi * 0.22 ** 2)) < 0.02 * peak
