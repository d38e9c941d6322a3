"""Synthetic stand-in scenes.  Every binary asset of the reference is a git-LFS pointer stub and there is no Cornell
box in its tree, so the BASELINE.json configurations are exercised on scenes authored here:
  cornell_box()  — C1: 36 triangles, roughness-1 white/red/green walls, emissive ceiling quad (17,12,4), two boxes
  city_block()   — C2/C3 stand-in for Bistro exterior: streets, tessellated facades, glass, alpha-tested foliage, emissive lamps,
                   ~250 materials drawn from the parameter ranges of Assets/Materials/bistro.*.material.json, procedural textures
                   with full mip chains, procedural sky cube
"""
import math
import numpy as np
from .scene_builder import SceneBuilder, Material, bridge_camera, identity34, translate_scale


def _quad(p0, p1, p2, p3, material, uv_scale=1.0):
    """Two triangles p0,p1,p2 / p0,p2,p3 (counter-clockwise seen from the side the normal points to)."""
    P = np.array([p0, p1, p2, p3], np.float32)
    n = np.cross(P[1] - P[0], P[2] - P[0]); n = n / np.linalg.norm(n)
    return dict(positions=P, indices=np.array([[0, 1, 2], [0, 2, 3]], np.uint32), normals=np.tile(n.astype(np.float32), (4, 1)),
                uvs=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32) * uv_scale, material=material)


def _box(corners_bottom, height, material):
    """Closed box from 4 bottom corners (counter-clockwise seen from above), outward facing."""
    b = [np.array(c, np.float32) for c in corners_bottom]
    t = [c + np.array([0, height, 0], np.float32) for c in b]
    geos = [_quad(t[0], t[3], t[2], t[1], material)]                                # top (normal +y)
    geos.append(_quad(b[0], b[1], b[2], b[3], material))                            # bottom (normal -y)
    for i in range(4):
        j = (i + 1) % 4
        geos.append(_quad(b[j], b[i], t[i], t[j], material))                        # sides
    return geos


def _merge(geos, material):
    pos, idx, nrm, uvs = [], [], [], []
    base = 0
    for g in geos:
        pos.append(g["positions"]); nrm.append(g["normals"]); uvs.append(g["uvs"]); idx.append(g["indices"] + base); base += len(g["positions"])
    return dict(positions=np.concatenate(pos), indices=np.concatenate(idx), normals=np.concatenate(nrm), uvs=np.concatenate(uvs), material=material)


def cornell_box(width=256, height=256, analytic_lights=False, delta_surfaces=False):
    """Returns (scene, camera).  Classic Cornell data in metres (x right, y up, z into the box); the camera looks down +z.
    analytic_lights adds a sphere light, a spot light and a zero-radius point light (the reference never samples the last)."""
    scene = cornell_builder(analytic_lights, delta_surfaces).build()
    cam = bridge_camera(width, height, pos=(2.78, 2.73, -8.0), direction=(0, 0, 1), up=(0, 1, 0), fov_y=0.66)
    return scene, cam


def cornell_builder(analytic_lights=False, delta_surfaces=False):
    """The SceneBuilder behind cornell_box (also written out as glTF by the loader tests).  delta_surfaces: the tall box becomes a perfect mirror and the
    short box clear glass, both enabled for path-space decomposition - what realtime mode splits into stable planes."""
    b = SceneBuilder()
    if analytic_lights:
        b.add_point_light(position=(1.2, 3.9, 1.6), color=(0.4, 0.6, 1.0), intensity=14.0, radius=0.22)
        b.add_spot_light(position=(4.4, 4.6, 1.2), direction=(-0.5, -1.0, 0.55), color=(1.0, 0.85, 0.6), intensity=60.0, radius=0.12, inner_angle=18.0, outer_angle=34.0)
        b.add_point_light(position=(2.7, 2.0, 2.7), color=(1.0, 1.0, 1.0), intensity=5.0, radius=0.0)
    white = b.add_material(Material(base_color=(0.73, 0.73, 0.73), roughness=1.0, metalness=0.0))
    red = b.add_material(Material(base_color=(0.65, 0.05, 0.05), roughness=1.0))
    green = b.add_material(Material(base_color=(0.12, 0.45, 0.15), roughness=1.0))
    light = b.add_material(Material(base_color=(0.78, 0.78, 0.78), roughness=1.0, emissive=(17.0, 12.0, 4.0)))
    S = 5.55
    floor = _quad((0, 0, 0), (0, 0, S), (S, 0, S), (S, 0, 0), white)               # normal +y
    ceiling = _quad((0, S, 0), (S, S, 0), (S, S, S), (0, S, S), white)             # normal -y
    back = _quad((0, 0, S), (0, S, S), (S, S, S), (S, 0, S), white)                # normal -z
    left = _quad((0, 0, 0), (0, S, 0), (0, S, S), (0, 0, S), red)                  # x = 0, normal +x
    right = _quad((S, 0, 0), (S, 0, S), (S, S, S), (S, S, 0), green)               # x = S, normal -x
    lq = _quad((2.13, S - 0.01, 2.27), (3.43, S - 0.01, 2.27), (3.43, S - 0.01, 3.32), (2.13, S - 0.01, 3.32), light)   # normal -y (faces down)
    short = _box([(1.30, 0, 0.65), (0.82, 0, 2.25), (2.40, 0, 2.72), (2.90, 0, 1.14)], 1.65, white)
    tall = _box([(4.23, 0, 2.47), (2.65, 0, 2.96), (3.14, 0, 4.56), (4.72, 0, 4.06)], 3.30, white)
    room = b.add_mesh([_merge([floor, ceiling, back], white), left, right])
    lamp = b.add_mesh([lq])
    if delta_surfaces:
        mirror = b.add_material(Material(base_color=(0.95, 0.93, 0.88), roughness=0.0, metalness=1.0, psd_exclude=False, psd_dominant_delta_lobe=1))
        glass = b.add_material(Material(base_color=(0.9, 0.97, 1.0), roughness=0.0, transmission=1.0, ior=1.5, thin_surface=False, nested_priority=2,
                                        volume_color=(0.8, 0.95, 0.9), volume_distance=1.5, psd_exclude=False, psd_dominant_delta_lobe=0))
        boxes = b.add_mesh([_merge(short, glass), _merge(tall, mirror)])
    else:
        boxes = b.add_mesh([_merge(short, white), _merge(tall, white)])
    for m in (room, lamp, boxes):
        b.add_instance(m, identity34())
    if analytic_lights:
        # proxy geometry of the first sphere light (radius 0.22): a cube inscribed in it; BSDF rays that reach it evaluate the analytic sphere
        proxy = b.add_material(Material(base_color=(0, 0, 0), roughness=1.0, analytic_light_proxy=True))
        h = 0.12
        cube = b.add_mesh([_merge(_box([(-h, -h, -h), (h, -h, -h), (h, -h, h), (-h, -h, h)], 2 * h, proxy), proxy)])
        b.add_instance(cube, translate_scale((1.2, 3.9, 1.6)), proxy_light=0)
    return b


def light_gallery(width=128, height=96, bays=6, lamps_per_bay=2):
    """A corridor of `bays` alcoves seen from the side, each lit by its own small ceiling lamps of a different colour and separated from its neighbours by full-height
    partitions: every screen region is lit by a handful of the scene's emissive triangles and shadowed from all others - the case NEE-AT's per-tile samplers exist for.
    Returns (Scene, CameraData)."""
    b = SceneBuilder()
    white = b.add_material(Material(base_color=(0.7, 0.7, 0.7), roughness=1.0))
    bw, depth, hgt = 2.0, 3.0, 2.6
    L = bays * bw
    geos = [_quad((0, 0, 0), (0, 0, depth), (L, 0, depth), (L, 0, 0), white),                 # floor, +y
            _quad((0, hgt, 0), (L, hgt, 0), (L, hgt, depth), (0, hgt, depth), white),         # ceiling, -y
            _quad((0, 0, depth), (0, hgt, depth), (L, hgt, depth), (L, 0, depth), white)]     # back wall, -z
    for k in range(bays + 1):                                                                 # partitions (two-sided: one quad per side), leaving the front open
        x = k * bw
        geos.append(_quad((x + 0.01, 0, 0.6), (x + 0.01, hgt, 0.6), (x + 0.01, hgt, depth), (x + 0.01, 0, depth), white))
        geos.append(_quad((x - 0.01, 0, 0.6), (x - 0.01, 0, depth), (x - 0.01, hgt, depth), (x - 0.01, hgt, 0.6), white))
    room = b.add_mesh([_merge(geos, white)])
    lamps = []
    rng = np.random.default_rng(7)
    for k in range(bays):
        col = 0.3 + 0.7 * rng.random(3)
        m = b.add_material(Material(base_color=(0.8, 0.8, 0.8), roughness=1.0, emissive=tuple((col * (6.0 + 10.0 * rng.random())).tolist())))
        for j in range(lamps_per_bay):
            cx = k * bw + bw * (j + 1) / (lamps_per_bay + 1); cz = 1.2 + 0.9 * rng.random(); r = 0.12
            lamps.append(_quad((cx - r, hgt - 0.02, cz - r), (cx + r, hgt - 0.02, cz - r), (cx + r, hgt - 0.02, cz + r), (cx - r, hgt - 0.02, cz + r), m))
    lamp_mesh = b.add_mesh(lamps)
    b.add_instance(room, identity34()); b.add_instance(lamp_mesh, identity34())
    cam = bridge_camera(width, height, pos=(L * 0.5, 1.3, -7.5), direction=(0, -0.02, 1), up=(0, 1, 0), fov_y=0.8)
    return b.build(), cam


# ----------------------------------------------------------------------------------------------------------------------
def _value_noise(rng, size, octaves=5):
    img = np.zeros((size, size), np.float32)
    amp = 1.0; tot = 0.0
    for o in range(octaves):
        n = 4 << o
        if n > size:
            break
        g = rng.random((n, n), dtype=np.float32)
        rep = size // n
        up = np.kron(g, np.ones((rep, rep), np.float32))
        # cheap smoothing: box blur by rolling
        up = (up + np.roll(up, rep // 2, 0) + np.roll(up, rep // 2, 1) + np.roll(np.roll(up, rep // 2, 0), rep // 2, 1)) * 0.25
        img += amp * up; tot += amp; amp *= 0.5
    return img / tot


def _facade(origin, du, dv, nu, nv, normal, rng, material, depth=0.15):
    """Tessellated facade: (nu x nv) quads with per-vertex inset displacement (window reveals, ledges)."""
    u = np.linspace(0, 1, nu + 1, dtype=np.float32); v = np.linspace(0, 1, nv + 1, dtype=np.float32)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    disp = (rng.random(uu.shape, dtype=np.float32) < 0.35).astype(np.float32) * (-depth) * rng.random(uu.shape, dtype=np.float32)
    disp[0, :] = 0; disp[-1, :] = 0; disp[:, 0] = 0; disp[:, -1] = 0
    P = origin[None, None, :] + uu[..., None] * du[None, None, :] + vv[..., None] * dv[None, None, :] + disp[..., None] * normal[None, None, :]
    P = P.reshape(-1, 3).astype(np.float32)
    i = np.arange(nu, dtype=np.uint32)[None, :] + (np.arange(nv, dtype=np.uint32) * (nu + 1))[:, None]
    i = i.reshape(-1)
    tris = np.stack([np.stack([i, i + 1, i + nu + 2], 1), np.stack([i, i + nu + 2, i + nu + 1], 1)], 1).reshape(-1, 3).astype(np.uint32)
    # smooth-ish normals from the displaced grid
    fn = np.cross(P[tris[:, 1]] - P[tris[:, 0]], P[tris[:, 2]] - P[tris[:, 0]])
    N = np.zeros_like(P)
    for k in range(3):
        np.add.at(N, tris[:, k], fn)
    ln = np.linalg.norm(N, axis=1, keepdims=True); N = np.where(ln > 0, N / np.maximum(ln, 1e-20), normal[None, :]).astype(np.float32)
    lu = np.linalg.norm(du); lv = np.linalg.norm(dv)
    UV = np.stack([uu.reshape(-1) * lu * 0.25, vv.reshape(-1) * lv * 0.25], 1).astype(np.float32)
    return dict(positions=P, indices=tris, normals=N, uvs=UV, material=material)


def _uv_sphere(center, radius, nu, nv, material):
    th = np.linspace(0, math.pi, nv + 1, dtype=np.float32); ph = np.linspace(0, 2 * math.pi, nu + 1, dtype=np.float32)
    pp, tt = np.meshgrid(ph, th, indexing="xy")
    N = np.stack([np.sin(tt) * np.cos(pp), np.cos(tt), np.sin(tt) * np.sin(pp)], -1).reshape(-1, 3).astype(np.float32)
    P = (np.asarray(center, np.float32)[None, :] + radius * N).astype(np.float32)
    i = (np.arange(nu, dtype=np.uint32)[None, :] + (np.arange(nv, dtype=np.uint32) * (nu + 1))[:, None]).reshape(-1)
    tris = np.stack([np.stack([i, i + nu + 2, i + 1], 1), np.stack([i, i + nu + 1, i + nu + 2], 1)], 1).reshape(-1, 3).astype(np.uint32)
    UV = np.stack([pp.reshape(-1) / (2 * math.pi) * 4, tt.reshape(-1) / math.pi * 2], 1).astype(np.float32)
    return dict(positions=P, indices=tris, normals=N, uvs=UV, material=material)


def sky_cube(size=256, sun_dir=(0.35, 0.55, -0.75), sun_intensity=60.0):
    """Procedural daylight cube (zenith/horizon gradient + a sun disc with halo), RGBA32F, D3D face order."""
    s = np.asarray(sun_dir, np.float32); s = s / np.linalg.norm(s)
    t = (np.arange(size, dtype=np.float32) + 0.5) / size * 2 - 1
    sc, tc = np.meshgrid(t, t, indexing="xy")
    one = np.ones_like(sc)
    dirs = [np.stack([one, -tc, -sc], -1), np.stack([-one, -tc, sc], -1), np.stack([sc, one, tc], -1),
            np.stack([sc, -one, -tc], -1), np.stack([sc, -tc, one], -1), np.stack([-sc, -tc, -one], -1)]
    faces = np.zeros((6, size, size, 4), np.float32)
    for f, d in enumerate(dirs):
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        up = np.clip(d[..., 1], -1, 1)
        horizon = np.array([0.85, 0.9, 1.0], np.float32) * 1.2; zenith = np.array([0.15, 0.35, 0.9], np.float32) * 1.5
        ground = np.array([0.12, 0.11, 0.10], np.float32)
        k = np.clip(up, 0, 1)[..., None] ** 0.5
        col = np.where(up[..., None] >= 0, horizon * (1 - k) + zenith * k, ground)
        c = np.clip((d * s).sum(-1), -1, 1)
        sun = (c > 0.9995).astype(np.float32) * sun_intensity + np.exp((c - 1) * 600.0) * 4.0
        col = col + sun[..., None] * np.array([1.0, 0.93, 0.8], np.float32)
        faces[f, ..., :3] = col; faces[f, ..., 3] = 1
    return faces


def city_block(target_triangles=2_800_000, width=1920, height=1080, seed=1234, texture_size=512, n_textures=24, n_materials=254,
               with_env=True, emissive=True, delta_surfaces=False):
    """Procedural stand-in for Bistro exterior.  Returns (scene, camera).
    delta_surfaces (BASELINE configs[2], realtime mode): shop fronts glazed over four storeys with clear glass and a wet street, both opted into the path-space decomposition
    (PSDExclude off, the way Bistro's .material.json files opt glass and puddles in), so that stable planes 1 and 2 cover a real share of the frame.  Geometry, triangle count
    and every other material are the same as without it."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder()
    # textures: albedo (sRGB), ORM (linear, G = roughness, B = metalness), alpha-noise (for foliage)
    alb, orm = [], []
    for i in range(n_textures):
        n = _value_noise(rng, texture_size)
        tint = 0.35 + 0.6 * rng.random(3, dtype=np.float32)
        img = np.zeros((texture_size, texture_size, 4), np.uint8)
        img[..., :3] = np.clip((0.45 + 0.55 * n[..., None]) * tint * 255, 0, 255).astype(np.uint8); img[..., 3] = 255
        alb.append(b.add_texture(img, srgb=True))
        m = _value_noise(rng, texture_size, 4)
        o = np.zeros((texture_size, texture_size, 4), np.uint8)
        o[..., 0] = 255; o[..., 1] = np.clip((0.5 + 0.5 * m) * 255, 0, 255).astype(np.uint8); o[..., 2] = 255; o[..., 3] = 255
        orm.append(b.add_texture(o, srgb=False))
    leaf = np.zeros((texture_size, texture_size, 4), np.uint8)
    ln = _value_noise(rng, texture_size, 6)
    leaf[..., 0] = 40; leaf[..., 1] = np.clip(90 + 120 * ln, 0, 255).astype(np.uint8); leaf[..., 2] = 30
    leaf[..., 3] = np.where(ln > 0.5, 255, 0).astype(np.uint8)
    leaf_tex = b.add_texture(leaf, srgb=True)
    # materials: parameter ranges follow the reference's Bistro material JSONs (35 transmissive, 22 alpha tested, 21 emissive of 254)
    opaque, glass, foliage, lamps = [], [], [], []
    for i in range(n_materials):
        r = rng.random()
        if i < 35 * n_materials // 254:
            rough = float(0.02 + 0.1 * rng.random())
            glass.append(b.add_material(Material(base_color=tuple(0.85 + 0.15 * rng.random(3)), roughness=min(rough, 0.06) if delta_surfaces else rough, transmission=1.0,
                                                 ior=1.5, thin_surface=bool(rng.random() < 0.7), nested_priority=2,
                                                 volume_color=(0.8, 0.9, 0.85), volume_distance=0.5,
                                                 psd_exclude=not delta_surfaces, psd_dominant_delta_lobe=0 if delta_surfaces else -1)))
        elif i < (35 + 22) * n_materials // 254:
            foliage.append(b.add_material(Material(base_color=(1, 1, 1), roughness=0.8, base_texture=leaf_tex, alpha_test=True, alpha_cutoff=0.5,
                                                   diffuse_transmission=0.0)))
        elif i < (35 + 22 + 21) * n_materials // 254 and emissive:
            col = np.array([1.0, 0.75 + 0.25 * rng.random(), 0.4 + 0.5 * rng.random()]) * (8.0 + 30.0 * rng.random())
            lamps.append(b.add_material(Material(base_color=(0.9, 0.9, 0.9), roughness=0.6, emissive=tuple(col))))
        else:
            t = int(rng.integers(0, n_textures))
            opaque.append(b.add_material(Material(base_color=tuple(0.6 + 0.4 * rng.random(3)), roughness=float(0.25 + 0.75 * rng.random()),
                                                  metalness=float(1.0 if r < 0.08 else 0.0), base_texture=alb[t], orm_texture=orm[t])))
    if not lamps:
        lamps = opaque[:1]
    # layout: G x G blocks of buildings separated by streets
    G = 6; block = 28.0; street = 10.0; pitch = block + street
    n_build = G * G
    # triangle budget: ~88% facades, rest ground/trees/lamps
    per_build = target_triangles * 0.86 / n_build
    quads_per_facade = max(4, int(per_build / 2 / 4))
    geos_by_mesh = []
    extent = G * pitch
    for bx in range(G):
        for bz in range(G):
            x0 = bx * pitch - extent / 2 + street / 2; z0 = bz * pitch - extent / 2 + street / 2
            w = block * (0.7 + 0.3 * rng.random()); d = block * (0.7 + 0.3 * rng.random()); h = 8.0 + 30.0 * rng.random()
            nv = max(2, int(math.sqrt(quads_per_facade * h / max(w, d)))); nu = max(2, quads_per_facade // nv)
            mats = [opaque[int(rng.integers(0, len(opaque)))] for _ in range(4)]
            o = np.array([x0, 0, z0], np.float32)
            f = []
            f.append(_facade(o + np.array([0, 0, 0], np.float32), np.array([w, 0, 0], np.float32), np.array([0, h, 0], np.float32), nu, nv, np.array([0, 0, -1], np.float32), rng, mats[0]))
            f.append(_facade(o + np.array([w, 0, 0], np.float32), np.array([0, 0, d], np.float32), np.array([0, h, 0], np.float32), nu, nv, np.array([1, 0, 0], np.float32), rng, mats[1]))
            f.append(_facade(o + np.array([w, 0, d], np.float32), np.array([-w, 0, 0], np.float32), np.array([0, h, 0], np.float32), nu, nv, np.array([0, 0, 1], np.float32), rng, mats[2]))
            f.append(_facade(o + np.array([0, 0, d], np.float32), np.array([0, 0, -d], np.float32), np.array([0, h, 0], np.float32), nu, nv, np.array([-1, 0, 0], np.float32), rng, mats[3]))
            f.append(_quad(o + np.array([0, h, 0], np.float32), o + np.array([0, h, d], np.float32), o + np.array([w, h, d], np.float32), o + np.array([w, h, 0], np.float32), mats[0], uv_scale=4.0))
            # shop-front glass panes in front of the street-facing facade
            gm = glass[int(rng.integers(0, len(glass)))] if glass else mats[0]
            gh = min(12.0, h - 0.5) if delta_surfaces else 3.0
            f.append(_quad(o + np.array([w * 0.1, 0.3, -0.05], np.float32), o + np.array([w * 0.1, gh, -0.05], np.float32), o + np.array([w * 0.9, gh, -0.05], np.float32), o + np.array([w * 0.9, 0.3, -0.05], np.float32), gm))
            geos_by_mesh.append(f)
    # ground: tessellated, slightly noisy
    ng = max(8, int(math.sqrt(target_triangles * 0.06 / 2)))
    gnd = _facade(np.array([-extent / 2 - 20, 0, -extent / 2 - 20], np.float32), np.array([0, 0, extent + 40], np.float32), np.array([extent + 40, 0, 0], np.float32),
                  ng, ng, np.array([0, 1, 0], np.float32), rng,
                  b.add_material(Material(base_color=(0.25, 0.25, 0.27), roughness=0.05, psd_exclude=False, psd_dominant_delta_lobe=1)) if delta_surfaces else opaque[0], depth=0.02)
    geos_by_mesh.append([gnd])
    # trees (alpha-tested canopies) and street lamps (emissive spheres) along the streets
    trees, lamp_geos = [], []
    n_tree = 4 * G * G; seg = max(6, int(math.sqrt(target_triangles * 0.05 / n_tree / 2)))
    for i in range(n_tree):
        x = (rng.integers(0, G + 1) * pitch - extent / 2) + rng.uniform(-3, 3); z = rng.uniform(-extent / 2, extent / 2)
        trees.append(_uv_sphere((x, 4.0 + rng.random(), z), 2.0 + rng.random(), seg * 2, seg, foliage[int(rng.integers(0, len(foliage)))] if foliage else opaque[0]))
    n_lamp = 6 * G * G; lseg = max(4, int(math.sqrt(max(10_000, target_triangles * 0.004) / n_lamp / 2)))
    for i in range(n_lamp):
        x = rng.uniform(-extent / 2, extent / 2); z = (rng.integers(0, G + 1) * pitch - extent / 2) + rng.uniform(-3.5, 3.5)
        lamp_geos.append(_uv_sphere((x, 4.5, z), 0.18, lseg * 2, lseg, lamps[int(rng.integers(0, len(lamps)))]))
    geos_by_mesh.append(trees); geos_by_mesh.append(lamp_geos)
    for geos in geos_by_mesh:
        b.add_instance(b.add_mesh(geos), identity34())
    if with_env:
        b.set_env_cube(sky_cube(256))
    scene = b.build()
    # street-level camera standing in the street between the 3rd and 4th block columns, looking down the street with a slight yaw
    # (cf. /Cameras/Outside in Assets/bistro-programmer-art.scene.json: eye height 1.8, fovY 1.04)
    cam = bridge_camera(width, height, pos=(-extent / 2 + 3 * pitch + 1.0, 1.8, -extent / 2 + 0.4 * pitch), direction=(-0.22, 0.10, 1.0), up=(0, 1, 0), fov_y=1.04)
    return scene, cam
