"""Generates tests/golden/hit_golden.npz from the UNMODIFIED Rtxpt/Shaders/PathTracer/PathTracer.hlsli (HandleHit with GenerateScatterRay, HandleRussianRoulette, AccumulatePathRadiance),
PathTracerNEE.hlsli (HandleNEE: candidate loop, weighted reservoir, ProcessLightSample), PathTracerNestedDielectrics.hlsli, PathState.hlsli, PathPayload.hlsli and the sample generators,
compiled in place as C++ through oracle/ref_hlsl_shim.h behind the stub bridge oracle/ref_bridge_stub.h (oracle/_ref/ref_kat_bsdf, mode "hit"; PATH_TRACER_MODE_REFERENCE).
Run in the build container only:
    make -C oracle ref && python tests/golden/make_hit_golden.py
  hit_in [M,920], hit_out [M,64]: layouts in oracle/ref_kat_bsdf_main.cpp ("hit" mode).  One record = one path vertex: the incoming 80-byte path payload, the ray, the surface
  Bridge::loadSurface would return, the medium table, a 16-light NEE-AT scenario (12 emissive triangles + 4 sphere / spot lights whose records come from the same binary's "lights" /
  "spheres" modes); out: the outgoing payload, the last shadow ray and its answer, the exports, the pixel's feedback reservoir."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_bsdf_golden import run  # noqa: E402
from bsdf_records import make_records  # noqa: E402
from make_sampler_golden import pack_tile  # noqa: E402


def f16(x): return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)
def h16(x): return np.asarray(x, np.float32).astype(np.float16).view(np.uint16).astype(np.uint32)
def bits(x): return np.asarray(x, np.uint32).view(np.float32)

PF = dict(active=1 << 0, hit=1 << 1, transmission=1 << 2, specular=1 << 3, delta=1 << 4, inside=1 << 5, terminateNext=1 << 6, deltaTransmissionPath=1 << 11, deltaOnlyPath=1 << 12,
          onPlane=1 << 16, onBranch=1 << 17, baseScatterDiff=1 << 18, specHitTQueued=1 << 19, onDominant=1 << 20)


def rotations(rng, n):
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True); w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1)
    return R.astype(np.float32)


def envquad_inputs(rng, n, rot=None):
    """the "envquads" mode's input: node x, y, dimension (4 .. 256), weight, radiance, id, rotation (zeros = identity), random 2, viewer 3"""
    u = np.zeros((n, 24), np.float32); dim = 2 ** rng.integers(2, 9, n); u[:, 2] = dim; u[:, 0] = rng.integers(0, dim); u[:, 1] = rng.integers(0, dim)
    u[:, 3] = rng.gamma(2.0, 2.0, n); u[:, 4:7] = rng.gamma(2.0, 1.5, (n, 3)); u[:, 7] = rng.integers(0, 1 << 20, n)
    if rot is not None: u[:, 8:17] = rot
    u[:, 17:19] = rng.random((n, 2)); u[:, 19:22] = (rng.random((n, 3)) - 0.5) * 40
    return u


def light_records(rng, n, around):
    """n x 16 lights x 12 words: 4 environment quads, 8 emissive triangles and 4 sphere / spot lights placed around the shaded points"""
    rec = np.zeros((n, 16, 12), np.uint32)
    rec[:, :4, :] = run("envquads", envquad_inputs(rng, n * 4), 24)[:, :12].view(np.uint32).reshape(n, 4, 12)
    return rec, 4


def light_records_rest(rng, n, around, rec):
    u = np.zeros((n * 12, 24), np.float32); c = np.repeat(around, 12, axis=0)
    u[:, 0:3] = c + rng.normal(size=(n * 12, 3)).astype(np.float32) * np.float32(6); u[:, 3:6] = (rng.random((n * 12, 3)) - 0.5) * np.float32(3); u[:, 6:9] = (rng.random((n * 12, 3)) - 0.5) * np.float32(3)
    u[:, 9:12] = rng.gamma(2.0, 4.0, (n * 12, 3)).astype(np.float32)
    rec[:, 4:12, :8] = run("lights", u, 24)[:, :8].view(np.uint32).reshape(n, 12, 8)[:, :8]
    v = np.zeros((n * 4, 24), np.float32); c = np.repeat(around, 4, axis=0)
    v[:, 0:3] = c + rng.normal(size=(n * 4, 3)).astype(np.float32) * np.float32(5); v[:, 3] = np.float32(0.05) + rng.random(n * 4).astype(np.float32) * np.float32(0.5)
    v[:, 4:7] = rng.gamma(2.0, 30.0, (n * 4, 3)).astype(np.float32); v[:, 7] = rng.random(n * 4); v[:, 8] = rng.random(n * 4)
    v[:, 9:12] = rng.normal(size=(n * 4, 3)).astype(np.float32); v[:, 12] = np.cos(np.radians(rng.uniform(20, 80, n * 4))).astype(np.float32); v[:, 13] = np.float32(0.05) + rng.random(n * 4).astype(np.float32) * np.float32(0.3)
    rec[:, 12:, :] = run("spheres", v, 24)[:, :12].view(np.uint32).reshape(n, 4, 12)
    return rec


def make(rng, n, slots_pool, fill=False, build=False):
    r = np.zeros((n, 1024), np.float32)
    b = make_records(rng, n); V, N, T, B = b[:, 0:3], b[:, 3:6], b[:, 6:9], b[:, 9:12]
    ray_dir = -V; t = np.exp(rng.uniform(-2, 3, n)).astype(np.float32); origin = ((rng.random((n, 3)) - 0.5) * 40).astype(np.float32); pos = origin + ray_dir * t[:, None]
    r[:, 20:23], r[:, 23:26], r[:, 26] = origin, ray_dir, t
    r[:, 28:31] = pos
    fn = N + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.15); fn /= np.linalg.norm(fn, axis=1, keepdims=True); fn *= np.sign((fn * V).sum(1, keepdims=True)); r[:, 31:34] = fn     # facing the viewer
    r[:, 34:37], r[:, 37:40], r[:, 40:43] = N, T, B
    vn = N + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.05); r[:, 43:46] = vn / np.linalg.norm(vn, axis=1, keepdims=True)
    thin = b[:, 32]; front = (rng.random(n) < 0.7).astype(np.float32)
    r[:, 46] = front; r[:, 47] = rng.integers(0, 16, n) * (thin == 0); r[:, 48] = 255; r[:, 49] = thin; r[:, 50] = rng.random(n) < 0.05; r[:, 51] = rng.integers(0, 10, n)    # material ids 8, 9 lie past the table
    r[:, 52] = 1.0; r[:, 53] = f16(np.where(rng.random(n) < 0.3, rng.random(n) * 0.2, 0))
    em = rng.random(n) < 0.3; r[em, 54:57] = f16(rng.gamma(2.0, 2.0, (int(em.sum()), 3)))
    r[:, 57] = rng.random(n) < 0.1; r[:, 58] = rng.integers(0, 3, n)
    r[:, 60:74] = b[:, 18:32]
    if build:
        pure = rng.random(n) < 0.4       # mirrors and clear glass: no non-delta lobe at all, the case in which a path keeps walking the delta tree (plane 0: primary surface replacement)
        r[pure, 60:63] = 0; r[pure, 63] = np.float16(0.02); r[pure, 71] = 0; glass = pure & (rng.random(n) < 0.5); r[glass, 72] = 1.0; r[glass, 49] = rng.random(int(glass.sum())) < 0.3
    if fill or build: r[rng.random(n) < (0.6 if build else 0.45), 63] = np.float16(0.02)        # delta lobes (mirrors, clear glass): what the stable planes follow
    ior = f16(np.where(rng.random(n) < 0.7, 1.5, 1.0 + rng.random(n) * 1.2)); r[:, 74] = ior
    r[:, 75] = np.where(em & (rng.random(n) < 0.8), rng.integers(4, 12, n), -1); r[:, 76] = np.where(rng.random(n) < 0.15, rng.integers(12, 16, n), -1)
    r[:, 77:80] = pos + rng.normal(size=(n, 3)).astype(np.float32) * np.float32(0.01)
    # constants
    r[:, 80] = rng.integers(1, 9, n); r[:, 81] = rng.integers(0, 5, n); r[:, 82] = rng.integers(0, 4096, n); r[:, 83] = rng.integers(1, 9, n); r[:, 84] = np.where(rng.random(n) < 0.75, 1, 2)      # NEEFullSamples: 1 is RTXPT's default (and what the CUDA tier supports)
    r[:, 85] = np.float32(2.0) + rng.random(n).astype(np.float32) * 8;      # never 0: the binary is built with RTXPT_FIREFLY_FILTER = 1, and the application passes threshold 0 only together with the macro off
    r[:, 86] = rng.choice(np.float32([0.0, 0.35, 0.65, 1.0]), n); r[:, 87] = rng.choice(np.float32([1.0, 0.5, 0.25]), n)
    r[:, 88:90] = rng.integers(0, 8, (n, 2)); r[:, 91] = 0.3; r[:, 92] = rng.random(n) < 0.85; r[:, 93] = rng.choice(np.float32([0.0, 1.0, 2.5]), n)
    # materials
    r[:, 96:104] = f16(np.where(rng.random((n, 8)) < 0.6, 1.5, 1.0 + rng.random((n, 8)) * 1.2)); r[:, 104:128] = rng.random((n, 24)) ** 0.3; r[:, 104:107] = 0.0   # material 0: black absorber (the 1e-7 clamp)
    r[:, 128:136] = np.exp(rng.uniform(-3, 3, (n, 8))); r[:, 129] = 0.0                                                                                                   # material 1: zero distance (the 1e-30 clamp)
    # light scenario
    lights, _ = light_records(rng, n, pos); lights = light_records_rest(rng, n, pos, lights); r[:, 728:920] = lights.reshape(n, 192).view(np.float32)
    r[:, 950:959] = rotations(rng, n); r[:, 959] = np.float32(0.5) + rng.random(n).astype(np.float32) * 2
    r[:, 27] = rng.random(n) < 0.25                                                          # a quarter of the rays leave the scene: HandleMiss
    r[r[:, 27] == 1, 26] = 1e15                                                              # ... with rayTCurrent = kMaxRayTravel, as nextHit passes it (PathTracerSample.hlsl:127)
    for i in range(n):
        counters = rng.integers(0, 9, 16) * (rng.random(16) < 0.8)
        if counters.sum() == 0: counters[rng.integers(0, 16)] = 3
        while counters.sum() > 64: counters[np.argmax(counters)] -= 1
        idx = np.repeat(np.arange(16), counters); r[i, 90] = len(idx); r[i, 136:152] = counters; r[i, 152:152 + len(idx)] = idx
        tiles = np.stack([pack_tile(rng.choice(rng.choice(16, rng.integers(1, 9), replace=False), 128)) for _ in range(4)]); r[i, 216:728] = tiles.reshape(-1).view(np.float32)
    # the incoming path
    p = np.zeros((n, 20), np.uint32)
    p[:, 0:3] = origin.view(np.uint32); px = rng.integers(0, 8, (n, 2)).astype(np.uint32); p[:, 3] = (px[:, 0] << 16) | px[:, 1]
    p[:, 4:7] = ray_dir.view(np.uint32); p[:, 7] = np.exp(rng.uniform(-3, 4, n)).astype(np.float32).view(np.uint32)
    thp = rng.random((n, 3)).astype(np.float32) ** np.float32(0.5) * np.float32(1.1); thp[rng.random(n) < 0.05] *= np.float32(0.01)
    p[:, 8] = h16(thp[:, 0]) | (h16(thp[:, 1]) << 16); p[:, 9] = h16(thp[:, 2])
    L = rng.gamma(1.0, 0.5, (n, 4)).astype(np.float32) * (rng.random((n, 1)) < 0.6); p[:, 10] = h16(L[:, 0]) | (h16(L[:, 1]) << 16); p[:, 11] = h16(L[:, 2]) | (h16(L[:, 3]) << 16)
    inside = rng.random(n) < 0.35; pick = rng.integers(0, len(slots_pool), n); p[inside, 12] = slots_pool[pick[inside], 0]; p[inside, 13] = slots_pool[pick[inside], 1]
    p[:, 14] = rng.integers(0, 4, n).astype(np.uint32) | (rng.integers(0, 6, n).astype(np.uint32) << 8) | (rng.integers(0, 5, n).astype(np.uint32) << 16)
    p[:, 15] = rng.integers(1, 64, n)
    p[:, 16] = (h16(np.exp(rng.uniform(-7, 0, n))) << 16) | h16(np.exp(rng.uniform(-8, -1, n)))
    pdf = f16(np.exp(rng.uniform(-3, 5, n))) * (rng.random(n) < 0.85); p[:, 17] = (h16(f16(rng.random(n) * 0.99 + 0.01)) << 16) | h16(pdf)
    mis = ((rng.random(n) < 0.8).astype(np.uint32) << 15) | ((rng.random(n) < 0.5).astype(np.uint32) << 13) | (r[:, 83].astype(np.uint32) << 6) | r[:, 84].astype(np.uint32)
    p[:, 18] = (mis << 16) | h16(np.where(rng.random(n) < 0.7, 1.0, 1.0 + rng.random(n) * 2))
    flags = np.full(n, PF["active"] | PF["hit"], np.uint32)
    for name, prob in (("transmission", 0.2), ("specular", 0.4), ("delta", 0.2), ("inside", 0.3), ("terminateNext", 0.15), ("deltaOnlyPath", 0.4), ("deltaTransmissionPath", 0.1), ("baseScatterDiff", 0.4)):
        flags |= (rng.random(n) < prob).astype(np.uint32) * np.uint32(PF[name])
    vertex = rng.integers(0, 7, n).astype(np.uint32)
    if fill:
        # FILL pass: the path tracks the delta tree of the BUILD pass.  The pixel's header holds three plane branch ids - often the id the path will have after this scatter (stable
        # branch id << 2 | delta lobe), so that landing on a plane (commit of the noisy radiance, plane switch, dominant flag) happens - and the planes hold earlier radiance
        for name, prob in (("onPlane", 0.5), ("onBranch", 0.6), ("specHitTQueued", 0.3), ("onDominant", 0.5)): flags |= (rng.random(n) < prob).astype(np.uint32) * np.uint32(PF[name])
        flags |= rng.integers(0, 3, n).astype(np.uint32) << 14                                    # stable plane index
        branch = np.where(rng.random(n) < 0.5, 1, rng.integers(1, 1 << 10, n)).astype(np.uint32); p[:, 15] = branch
        hdr = np.full((n, 4), 0xFFFFFFFF, np.uint32)
        for k in range(3):
            kind = rng.random(n); lobe = rng.integers(0, 3, n).astype(np.uint32)
            hdr[:, k] = np.where(kind < 0.45, (branch << 2) | lobe, np.where(kind < 0.7, ((branch << 2) | lobe) << 2 | rng.integers(0, 3, n).astype(np.uint32), np.where(kind < 0.85, rng.integers(1, 1 << 12, n).astype(np.uint32), 0xFFFFFFFF)))
        hdr[:, 3] = (np.exp(rng.uniform(-1, 4, n)).astype(np.float32).view(np.uint32) & np.uint32(0xFFFFFFFC)) | rng.integers(0, 3, n).astype(np.uint32)
        r[:, 920:924] = hdr.view(np.float32)
        rad = rng.gamma(1.0, 0.5, (n, 3, 4)).astype(np.float32) * (rng.random((n, 3, 1)) < 0.6); rad[rng.random((n, 3)) < 0.15, 0:2] = 0     # r = g = 0 with b > 0: the commit's "both words non-zero" test
        r[:, 924:930] = np.stack([h16(rad[..., 0]) | (h16(rad[..., 1]) << 16), h16(rad[..., 2]) | (h16(rad[..., 3]) << 16)], axis=-1).reshape(n, 6).view(np.float32)
        r[:, 930] = np.where(rng.random(n) < 0.5, -np.exp(rng.uniform(-3, 4, n)), np.where(rng.random(n) < 0.5, 0, np.exp(rng.uniform(-3, 3, n))))
    if build:
        # BUILD pass: the delta tree is explored plane by plane.  Words 10-11 of the payload hold the packed image transform (two 30-bit octahedral rows + handedness), word 17 the scene
        # length at which motion vectors were blocked (mostly 0 = not blocked), word 18 nothing; the header says which planes are free (invalid id), enqueued or taken
        p[:, 10] = rng.integers(0, 1 << 30, n); p[:, 11] = rng.integers(0, 1 << 30, n).astype(np.uint32) | (rng.integers(0, 2, n).astype(np.uint32) << 31)
        p[:, 17] = np.where(rng.random(n) < 0.75, 0, np.exp(rng.uniform(-2, 3, n))).astype(np.float32).view(np.uint32); p[:, 18] = 0
        cur = rng.integers(0, 3, n).astype(np.uint32); flags |= cur << 14; flags |= (rng.random(n) < 0.6).astype(np.uint32) * np.uint32(PF["onDominant"])
        p[:, 15] = np.where(rng.random(n) < 0.4, 1, rng.integers(1, 1 << 10, n)).astype(np.uint32)
        kind = rng.random((n, 3)); hdr = np.where(kind < 0.6, 0xFFFFFFFF, np.where(kind < 0.8, 0xFFFFFFFE, rng.integers(1, 1 << 12, (n, 3)))).astype(np.uint32); hdr[np.arange(n), cur] = 0
        r[:, 920:923] = hdr.view(np.float32); r[:, 923] = ((np.exp(rng.uniform(-1, 4, n)).astype(np.float32).view(np.uint32) & np.uint32(0xFFFFFFFC)) | rng.integers(0, 3, n).astype(np.uint32)).view(np.float32)
        r[:, 931:934] = (rng.random((n, 3)) - 0.5) * 20; d = rng.normal(size=(n, 3)); r[:, 934:937] = d / np.linalg.norm(d, axis=1, keepdims=True); r[:, 937:943] = rng.normal(size=(n, 6)) * 0.01
        r[:, 943] = rng.integers(1, 10, n); r[:, 944] = rng.random(n) < 0.7; r[:, 945] = 0.95; r[:, 946:949] = f16(rng.gamma(1.0, 0.5, (n, 3)) * (rng.random((n, 1)) < 0.5))
    p[:, 19] = (flags << 10) | vertex
    r[:, 0:20] = p.view(np.float32)
    return r


if __name__ == "__main__":
    rng = np.random.default_rng(777)
    g = np.load(os.path.join(ROOT, "tests", "golden", "interior_golden.npz")); slots = g["interior_out"].reshape(-1, 12, 6)[:, :, 0:2].reshape(-1, 2).view(np.uint32); slots = slots[(slots != 0).any(1)]
    e = envquad_inputs(np.random.default_rng(776), 2000, rotations(np.random.default_rng(775), 2000)); e[:200, 8:17] = 0
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "envquad_lights_golden.npz"), envquads_in=e, envquads_out=run("envquads", e, 24),
                        source=np.array("Rtxpt/Shaders/PathTracer/Lighting/PolymorphicLight.hlsli (EnvironmentQuadLight) at reference commit f08d1c7, compiled as C++ by oracle/Makefile target _ref/ref_kat_bsdf"))
    def with_ops(u, n_init, seed):
        """op 2 (EmptyPathInitialize) on the first n_init records: only the pixel, the cone spread angle and the bounce limits (0 ends the path at the first vertex) matter"""
        r2 = np.random.default_rng(seed); u[:n_init, 27] = 2; u[:n_init, 1020] = np.exp(r2.uniform(-9, -5, n_init)); u[:n_init // 4, 80] = 0
        return u

    u = with_ops(make(rng, 1200, slots), 60, 1)
    out = run("hit", u, 128)
    src = "Rtxpt/Shaders/PathTracer/{PathTracer,PathTracerNEE,PathTracerNestedDielectrics,PathTracerStablePlanes,StablePlanes,PathState,PathPayload}.hlsli + Utils/SampleGenerators.hlsli + Rtxpt/Shaders/PathTracerSample.hlsl:33-113 at reference commit f08d1c7, compiled as C++ by oracle/Makefile targets _ref/ref_kat_bsdf (PATH_TRACER_MODE 0), _ref/ref_kat_pt_build (1) and _ref/ref_kat_pt_fill (2) behind oracle/ref_bridge_stub.h"
    ub = with_ops(make(np.random.default_rng(779), 1200, slots, build=True), 60, 2)
    outb = run("hit", ub, 128, exe="ref_kat_pt_build")
    # op 4 (postProcessHit): what the BUILD records above left behind - the pixel's header with planes enqueued for exploration, the three planes, the path as it ended
    U, O = ub.view(np.uint32), outb.view(np.uint32); enq = np.where((O[:, 47:50] == 0xFFFFFFFE).any(1) & (ub[:, 27] <= 1))[0]; r4 = np.random.default_rng(3); pick = r4.choice(enq, 200)
    u4 = ub[pick].copy(); u4[:, 27] = 4; u4[:, 0:20] = outb[pick, 0:20]; u4[:, 920:924] = outb[pick, 47:51]; u4[:, 960:1020] = outb[pick, 56:116]
    flip = r4.random(200) < 0.25; u4.view(np.uint32)[flip, 19] |= np.uint32(1 << 10)                    # a path that is still active explores nothing
    ub = np.concatenate([ub, u4]); outb = np.concatenate([outb, run("hit", u4, 128, exe="ref_kat_pt_build")])
    uf = with_ops(make(np.random.default_rng(778), 1200, slots, fill=True), 60, 3)
    # op 3 (FirstHitFromVBuffer): the FILL pass restarts from plane 0 as the BUILD records above stored it - surfaces and sky (scene length +inf: the miss is handled inline)
    base = np.where((O[:1200, 47] != 0xFFFFFFFF) & (O[:1200, 47] != 0xFFFFFFFE) & (O[:1200, 47] != 0) & (ub[:1200, 27] <= 1))[0]; r3 = np.random.default_rng(4); pick = r3.choice(base, 240)
    u3 = uf[:240].copy(); u3[:, 27] = 3; u3[:, 920:924] = outb[pick, 47:51]; u3[:, 960:1020] = outb[pick, 56:116]; u3[:, 1020] = np.exp(r3.uniform(-9, -5, 240))
    uf = np.concatenate([uf[240:], u3, uf[:240]])[:1200 + 240]
    outf = run("hit", uf, 128, exe="ref_kat_pt_fill")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hit_golden.npz"), hit_in=u, hit_out=out, fill_in=uf, fill_out=outf, build_in=ub, build_out=outb, source=np.array(src))
    print(u.shape, out.shape, uf.shape, outf.shape, ub.shape, outb.shape, os.path.getsize(os.path.join(ROOT, "tests", "golden", "hit_golden.npz")))
