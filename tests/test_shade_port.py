"""The CUDA source of the shading kernel held to the reference's own shader code on the CPU.

tests/emu/shade_host_emu.cu (test infrastructure, never linked into the product) compiles rtxpt_b200/csrc/shade.cuh - shadeHit / shadeMiss with bsdf.cuh, the NEE-AT sampler of
neeat.cuh and the sample generators of device_math.cuh, the very functions k_shade wraps - for the host, and runs them on the records of tests/golden/hit_golden.npz: path vertices whose
expected outcome was produced by the UNMODIFIED Rtxpt/Shaders/PathTracer/PathTracer.hlsli (HandleHit / HandleMiss with PathTracerNEE.hlsli, PathTracerNestedDielectrics.hlsli, the
sample generators ...) compiled in place behind a stub bridge (tests/golden/make_hit_golden.py).  The surface a hit loads and the environment cube are what that bridge supplies (hooks
under PT_HOST_EMU in shade.cuh); the shadow kernel's half of ProcessLightSample (radiance into L, feedback reservoir, Russian-roulette outcome of a visible sample) is applied by the
emulation as kernels.cu does.  What only the GPU build has - libdevice's transcendental functions, fast-math in the default build, the wavefront's queues - stays with the -m gpu tests."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s", "_build/libshade_emu.so"], check=True)
    L = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libshade_emu.so"))
    for f in (L.shade_emu_reference_vertex, L.shade_emu_build_vertex, L.shade_emu_fill_vertex): f.argtypes = [C.c_void_p, C.c_void_p]; f.restype = C.c_int
    return L


def _run(fn, u, ref):
    out = np.zeros_like(ref); ran = np.zeros(len(u), bool)
    for i in range(len(u)): ran[i] = fn(u[i].ctypes.data, out[i].ctypes.data) == 0
    return out, ran


PAYLOAD = list(range(20)); SHADOW = list(range(20, 29)); FEEDBACK = [39, 40]


def test_shade_kernel_source_matches_reference_path_tracer_golden():
    """Reference mode with NEE-AT feedback (k_shade< .., NEEAT >): every hit with NEEFullSamples = 1 (what this tier supports; rtxpt_b200_set_constants refuses more) and every miss of
    the golden - the outgoing 80-byte path state (all words but stableBranchID, which carries the sample index in reference mode), the shadow ray and its answer, the pixel's feedback
    reservoir - bit for bit."""
    L = _lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "hit_golden.npz"))
    u, ref = np.ascontiguousarray(g["hit_in"]), g["hit_out"]
    out, ran = _run(L.shade_emu_reference_vertex, u, ref)
    hits = ran & (u[:, 27] == 0) & (u[:, 84] == 1); misses = ran & (u[:, 27] == 1)
    assert hits.sum() > 500 and misses.sum() > 200
    cols = [c for c in PAYLOAD if c != 15] + SHADOW + FEEDBACK
    same = ref.view(np.uint32)[:, cols] == out.view(np.uint32)[:, cols]
    assert same[hits].all(), np.argwhere(~same[hits])[:8]
    assert same[misses].all(), np.argwhere(~same[misses])[:8]
    # the hits took every turn: shadow rays seen and blocked, feedback written, false hits rejected, paths ended by the bounce limit and by roulette
    assert 0.25 < ref[hits, 28].mean() < 0.7 and (ref[hits, 39] > 0).mean() > 0.2 and (ref[hits, 20] > 0).mean() > 0.5


def test_realtime_shade_kernel_source_matches_reference_path_tracer_golden():
    """The realtime passes (k_rt_shade< BUILD >, k_rt_shade< FILL, .., NEEAT >: realtime.cuh's stablePlanesHandleHit / HandleMiss / OnScatter, splitDeltaPath, storeStablePlane,
    storeExplorationStart, commitDenoiserRadiance, the specular hit distance) against the same shader code compiled as the BUILD and FILL shaders: the outgoing path state (all 20
    words), the pixel's three 80-byte stable planes, header, stable radiance, hit distance, shadow ray and feedback reservoir - bit for bit, hits and misses."""
    L = _lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "hit_golden.npz"))
    for key, fn, cols, single in (("build", L.shade_emu_build_vertex, PAYLOAD + [29, 31] + list(range(47, 51)) + list(range(52, 56)) + list(range(56, 116)), False),
                                  ("fill", L.shade_emu_fill_vertex, PAYLOAD + SHADOW + [37] + FEEDBACK + list(range(41, 51)), True)):
        u, ref = np.ascontiguousarray(g[key + "_in"]), g[key + "_out"]
        out, ran = _run(fn, u, ref)
        hits = ran & (u[:, 27] == 0) & ((u[:, 84] == 1) | (not single)); misses = ran & (u[:, 27] == 1)      # BUILD does no NEE: every record applies
        assert hits.sum() > 600 and misses.sum() > 200
        same = ref.view(np.uint32)[:, cols] == out.view(np.uint32)[:, cols]
        assert same[hits].all(), (key, np.argwhere(~same[hits])[:8])
        assert same[misses].all(), (key, np.argwhere(~same[misses])[:8])


def test_bsdf_source_matches_reference_header_golden():
    """bsdf.cuh's StandardBSDF (eval, evalPdf, sample, getLobes) compiled for the host against tests/golden/bsdf_golden.npz - the vectors produced by the reference's own material headers.
    Bit for bit on every record with all lobes active (what the bridge hands the path tracer; the golden's single-lobe records exercise a MaterialHeader mask the CUDA path has no use for).
    The same comparison runs on the device in tests/test_gpu_parity.py, there with libdevice's functions and a tolerance."""
    L = _lib(); L.shade_emu_bsdf.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]; L.shade_emu_bsdf.restype = None
    g = np.load(os.path.join(ROOT, "tests", "golden", "bsdf_golden.npz"))
    u, ref = np.ascontiguousarray(g["bsdf_in"]), np.ascontiguousarray(g["bsdf_out"][:, :16])
    out = np.zeros_like(ref); L.shade_emu_bsdf(u.ctypes.data, len(u), out.ctypes.data)
    all_lobes = u[:, 33] == 255
    same = (out.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(out) & np.isnan(ref))
    assert all_lobes.sum() > 3800 and same[all_lobes].all(), np.argwhere(~same[all_lobes])[:8]


def test_shade_kernel_source_agrees_with_oracle_on_recombined_vertices(oracle):
    """Beyond the golden's own records: 60 000 vertices per pass made by recombining the golden's parts (a path with its ray and surface; constants; medium table; light scenario;
    plane header; environment - each from a different record, counters / roulette words swapped in) - new situations the reference binary never saw, on which the oracle (which
    reproduces the reference on the golden) and the host build of the CUDA source must still agree on every word."""
    L = _lib(); L.shade_emu_vertices.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32]; L.shade_emu_vertices.restype = None
    O = oracle.lib(); O.oracle_hit_funcs.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]; O.oracle_hit_funcs.restype = None
    g = np.load(os.path.join(ROOT, "tests", "golden", "hit_golden.npz"))
    n = 60000
    for key, mode, cols in (("hit", 0, [c for c in PAYLOAD if c != 15] + SHADOW + FEEDBACK), ("build", 1, PAYLOAD + [29, 31] + list(range(47, 51)) + list(range(52, 56)) + list(range(56, 116))),
                            ("fill", 2, PAYLOAD + SHADOW + [37] + FEEDBACK + list(range(41, 51)))):
        rng = np.random.default_rng(100 + mode); u = g[key + "_in"]; u = u[u[:, 27] <= 1]; pick = lambda: rng.integers(0, len(u), n)
        r = u[pick()].copy()
        for lo, hi in ((80, 96), (96, 136), (136, 920), (920, 950), (950, 960)): r[:, lo:hi] = u[pick(), lo:hi]
        for w in (14, 17, 18): q = pick(); sel = rng.random(n) < 0.5; r[sel, w] = u[q[sel], w]
        r = np.ascontiguousarray(r)
        a = np.zeros((n, 128), np.float32); st = np.zeros(n, np.int32); L.shade_emu_vertices(r.ctypes.data, n, a.ctypes.data, st.ctypes.data, mode)
        b = np.zeros((n, 128), np.float32); O.oracle_hit_funcs(r.ctypes.data, n, b.ctypes.data, mode)
        run = (st == 0) & ((r[:, 84] == 1) | (mode == 1) | (r[:, 27] == 1))
        same = a.view(np.uint32)[:, cols] == b.view(np.uint32)[:, cols]
        assert run.sum() > 0.7 * n and same[run].all(), (key, np.argwhere(~same[run])[:8])
