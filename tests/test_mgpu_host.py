"""The C++ multi-GPU host (include/rtxpt_b200_mgpu.h, rtxpt_b200/csrc/multigpu_host.cpp): tile partition over N contexts + one ncclAllGather per frame.
CPU: the library loads, exports every symbol its header declares and refuses to run without a device.  GPU: with one device the host reproduces the single-context frame bit for
bit; with two or more (gpurun --gpus 2) every GPU ends up with the single-context frame, bit for bit, after the all-gather."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib(product):
    import torch  # noqa: F401  - torch bundles a NEWER libnccl.so.2 than the system one the C++ host links; in a process that uses both, torch's must be the one that gets loaded
    product.load()                                           # librtxpt_b200.so first: the multi-GPU host links against it
    L = C.CDLL(os.path.join(os.path.dirname(product.LIB_PATH), "librtxpt_b200_mgpu.so"))
    L.rtxpt_b200_mgpu_last_error.restype = C.c_char_p
    L.rtxpt_b200_mgpu_context.restype = C.c_void_p; L.rtxpt_b200_mgpu_context.argtypes = [C.c_void_p, C.c_uint32]
    L.rtxpt_b200_mgpu_local_count.restype = C.c_uint32; L.rtxpt_b200_mgpu_local_count.argtypes = [C.c_void_p]
    for name, args in (("create", [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]), ("destroy", [C.c_void_p]), ("upload_scene", [C.c_void_p, C.c_void_p]),
                       ("set_constants", [C.c_void_p, C.c_void_p]), ("render_frame", [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int]), ("synchronize", [C.c_void_p]),
                       ("last_frame_ms", [C.c_void_p, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float)]), ("unique_id", [C.c_void_p]),
                       ("create_rank", [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)])):
        fn = getattr(L, "rtxpt_b200_mgpu_" + name); fn.argtypes = args; fn.restype = C.c_int
    return L


def test_mgpu_library_exports_every_declared_symbol(product):
    src = open(os.path.join(ROOT, "include", "rtxpt_b200_mgpu.h")).read()
    names = sorted(set(re.findall(r"RTXPT_API\s+[\w\s\*]+?\b(rtxpt_b200_mgpu_\w+)\s*\(", src)))
    assert len(names) == 15, names
    L = _lib(product)
    for n in names: assert hasattr(L, n), n


def test_mgpu_refuses_without_a_device(product):
    import torch
    if torch.cuda.is_available(): pytest.skip("a CUDA device is present")
    from rtxpt_b200 import structs as S
    L = _lib(product); cfg = S.Config(); cfg.maxSubSamplesPerLaunch = 4; cfg.tileSize = 64; h = C.c_void_p()
    assert L.rtxpt_b200_mgpu_create(C.byref(cfg), 1, None, C.byref(h)) != 0 and b"no CUDA device" in L.rtxpt_b200_mgpu_last_error()


def _render_with_host(L, product, scene, consts, gpus, frames=2, spp=4):
    from rtxpt_b200 import structs as S
    cfg = S.Config(); cfg.maxSubSamplesPerLaunch = spp; cfg.tileSize = 32; h = C.c_void_p()
    assert L.rtxpt_b200_mgpu_create(C.byref(cfg), gpus, None, C.byref(h)) == 0, L.rtxpt_b200_mgpu_last_error()
    assert L.rtxpt_b200_mgpu_upload_scene(h, C.byref(scene.desc)) == 0, L.rtxpt_b200_mgpu_last_error()
    for f in range(frames):
        consts.sampleBaseIndex = f * spp
        assert L.rtxpt_b200_mgpu_set_constants(h, C.byref(consts)) == 0 and L.rtxpt_b200_mgpu_render_frame(h, 0, spp, 1, 1) == 0, L.rtxpt_b200_mgpu_last_error()
    assert L.rtxpt_b200_mgpu_synchronize(h) == 0
    W, H = consts.imageWidth, consts.imageHeight; imgs = []; PL = product.load()
    for i in range(L.rtxpt_b200_mgpu_local_count(h)):
        img = np.empty((H, W, 4), np.float32)
        assert PL.rtxpt_b200_readback(L.rtxpt_b200_mgpu_context(h, i), S.BUFFER_ACCUMULATED_F32, img.ctypes.data, img.nbytes) == 0
        t, x = C.c_float(), C.c_float(); assert L.rtxpt_b200_mgpu_last_frame_ms(h, i, C.byref(t), C.byref(x)) == 0 and t.value > 0
        imgs.append(img)
    assert L.rtxpt_b200_mgpu_destroy(h) == 0
    return imgs


def _single(product, scene, consts, frames=2, spp=4):
    c = product.Context(max_sub_samples_per_launch=spp); c.upload_scene(scene)
    for f in range(frames):
        consts.sampleBaseIndex = f * spp; c.set_constants(consts); c.path_trace(0, spp, True)
    c.synchronize(); img = c.readback_accumulated(); c.close()
    return img


@pytest.mark.gpu
def test_mgpu_one_device_equals_the_single_context(product, small_city):
    from rtxpt_b200 import scene_builder as sb
    scene, cam = small_city; W, H = cam.ViewportSize[0], cam.ViewportSize[1]
    consts = sb.make_constants(W, H, cam, env_enabled=True, firefly_threshold=5000.0)
    imgs = _render_with_host(_lib(product), product, scene, consts, 1)
    assert np.array_equal(imgs[0], _single(product, scene, consts))


@pytest.mark.gpu
def test_mgpu_all_devices_hold_the_single_context_frame(product, small_city):
    import torch
    n = torch.cuda.device_count()
    if n < 2: pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    from rtxpt_b200 import scene_builder as sb
    scene, cam = small_city; W, H = cam.ViewportSize[0], cam.ViewportSize[1]
    consts = sb.make_constants(W, H, cam, env_enabled=True, firefly_threshold=5000.0)
    ref = _single(product, scene, consts)
    for gpus in sorted({2, n}):
        for img in _render_with_host(_lib(product), product, scene, consts, gpus): assert np.array_equal(img, ref)


@pytest.mark.gpu
def test_cpp_multigpu_example_writes_the_frame(product, tmp_path):
    import subprocess
    import gltf_export
    from rtxpt_b200 import scenes
    exe = os.path.join(os.path.dirname(product.LIB_PATH), "multigpu_gltf")
    path = gltf_export.export(scenes.cornell_builder(), str(tmp_path / "cornell.gltf"), camera=dict(position=(2.78, 2.73, -8.0), direction=(0, 0, 1), up=(0, 1, 0), yfov=0.66, znear=0.1, zfar=1e7))
    out = tmp_path / "mg.pfm"
    r = subprocess.run([exe, str(path), str(out), "1", "96", "96", "8", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "gpu 0: last frame trace" in r.stderr and out.stat().st_size > 96 * 96 * 12


def _realtime_with_host(L, product, scene, cam, consts, gpus, frames=2):
    """The C++ host's realtime frame (rtxpt_b200_mgpu_render_realtime_frame) on `gpus` devices; returns every local context's LDR image and output colour."""
    from rtxpt_b200 import scene_builder as sb, structs as S
    for name, args in (("set_view", [C.c_void_p, C.c_void_p]), ("set_realtime", [C.c_void_p, C.c_void_p]), ("render_realtime_frame", [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int])):
        fn = getattr(L, "rtxpt_b200_mgpu_" + name); fn.argtypes = args; fn.restype = C.c_int
    W, H = consts.imageWidth, consts.imageHeight
    cfg = S.Config(); cfg.maxSubSamplesPerLaunch = 1; cfg.tileSize = 32; h = C.c_void_p()
    assert L.rtxpt_b200_mgpu_create(C.byref(cfg), gpus, None, C.byref(h)) == 0, L.rtxpt_b200_mgpu_last_error()
    assert L.rtxpt_b200_mgpu_upload_scene(h, C.byref(scene.desc)) == 0, L.rtxpt_b200_mgpu_last_error()
    view = S.ViewConstants(); view.matWorldToClip[:] = [float(x) for x in np.asarray(sb.world_to_clip(cam), np.float32).reshape(16)]
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=2); k = sb.make_denoiser_constants(cam); tm = S.make_tone_mapping_params(op=5, auto_exposure=True)
    assert L.rtxpt_b200_mgpu_set_constants(h, C.byref(consts)) == 0 and L.rtxpt_b200_mgpu_set_view(h, C.byref(view)) == 0 and L.rtxpt_b200_mgpu_set_realtime(h, C.byref(rt)) == 0, L.rtxpt_b200_mgpu_last_error()
    for f in range(frames):
        consts.sampleBaseIndex = 2 * f; frame = sb.make_reblur_frame(cam, cam, frame_index=f)
        assert L.rtxpt_b200_mgpu_set_constants(h, C.byref(consts)) == 0 and L.rtxpt_b200_mgpu_render_realtime_frame(h, C.byref(k), C.byref(frame), C.byref(tm), 0) == 0, L.rtxpt_b200_mgpu_last_error()
    assert L.rtxpt_b200_mgpu_synchronize(h) == 0
    PL = product.load(); out = []
    for i in range(L.rtxpt_b200_mgpu_local_count(h)):
        ldr = np.empty((H, W, 4), np.uint8); col = np.empty((H, W, 4), np.float16); ctx = L.rtxpt_b200_mgpu_context(h, i)
        assert PL.rtxpt_b200_readback(ctx, S.BUFFER_LDR_COLOR_RGBA8, ldr.ctypes.data, ldr.nbytes) == 0 and PL.rtxpt_b200_readback(ctx, S.BUFFER_OUTPUT_COLOR_F16, col.ctypes.data, col.nbytes) == 0
        out.append((ldr, col))
    assert L.rtxpt_b200_mgpu_destroy(h) == 0
    return out


@pytest.mark.gpu
def test_mgpu_realtime_frame_equals_the_single_context(product):
    """BASELINE configs[2]'s frame through the C++ host: one device reproduces rtxpt_b200_denoise_realtime's frame; with two or more devices every device ends up with that frame."""
    import torch
    from rtxpt_b200 import scene_builder as sb, scenes, structs as S
    W, H = 160, 128
    scene, cam = scenes.cornell_box(W, H, delta_surfaces=True)
    consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=3)
    c = product.Context(max_sub_samples_per_launch=1, tile_size=32); c.upload_scene(scene); c.set_constants(consts); c.set_view(sb.world_to_clip(cam))
    c.set_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=2)); k = sb.make_denoiser_constants(cam); tm = S.make_tone_mapping_params(op=5, auto_exposure=True)
    for f in range(2):
        consts.sampleBaseIndex = 2 * f; c.set_constants(consts)
        c.path_trace_realtime(False); c.denoise_realtime(k, sb.make_reblur_frame(cam, cam, frame_index=f)); c.tone_map(tm)
    c.synchronize(); ref_ldr, ref_col = c.readback_ldr(), c.readback_output_color(); c.close()
    for gpus in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        for ldr, col in _realtime_with_host(_lib(product), product, scene, cam, consts, gpus):
            assert ldr.tobytes() == ref_ldr.tobytes() and col.tobytes() == ref_col.tobytes(), gpus
