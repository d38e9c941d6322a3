"""Times realtime mode (stable-plane BUILD pass + FILL pass + no-denoiser merge) on bench.py's city workload and prints ONE JSON object.
Run by bench.py in a child process after the headline measurement (a failure here must not cost the headline line), or by hand on a GPU box:
    python scripts/bench_realtime.py [--frames 10] [--sub-samples 1]
Clear glass (roughness below the delta threshold) is opted into the path-space decomposition the way a .material.json with PSDExclude = false does, so the BUILD pass
has real forks to explore; everything else keeps RTXPT's default (excluded)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10); ap.add_argument("--warmup", type=int, default=3); ap.add_argument("--sub-samples", type=int, default=1)
    ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080); ap.add_argument("--triangles", type=int, default=2_800_000)
    args = ap.parse_args()
    real_stdout = os.dup(1); os.dup2(2, 1)
    from rtxpt_b200 import lib, scenes, scene_builder as sb, structs as S
    W, H = args.width, args.height
    scene, cam = scenes.city_block(target_triangles=args.triangles, width=W, height=H)
    opted = 0
    for i in range(scene.desc.materialCount):
        m = scene.desc.materials[i]
        if m.TransmissionFactor > 0 and m.Roughness * m.Roughness < 0.0064:
            m.Flags = (m.Flags & ~S.MATFLAG_PSDExclude & ~0x0F000000) | (1 << 24); opted += 1          # PSDExclude off, dominant delta lobe = transmission
    consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=6, env_enabled=True, firefly_threshold=5000.0, nee=True, nee_type=2)
    ctx = lib.Context(max_sub_samples_per_launch=1)
    ctx.upload_scene(scene); ctx.set_constants(consts); ctx.set_view(sb.world_to_clip(cam))
    rt = sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=args.sub_samples)
    ctx.set_realtime(rt)
    ms = []
    for f in range(args.warmup + args.frames):
        consts.sampleBaseIndex = f * args.sub_samples; ctx.set_constants(consts)
        ctx.path_trace_realtime(True); ctx.synchronize()
        st = ctx.stats()
        if f >= args.warmup: ms.append(float(st.msTotal))
    r = ctx.readback_realtime()
    hd = r["header"]
    out = {"ms_per_frame": float(np.median(ms)), "ms_min": float(np.min(ms)), "ms_max": float(np.max(ms)), "frames": args.frames, "warmup": args.warmup,
           "image": [W, H], "sub_samples": args.sub_samples, "active_planes": 3, "max_vertex_depth": int(rt.maxStablePlaneVertexDepth), "kernel_launches_per_frame": int(st.kernelLaunches),
           "materials_opted_into_decomposition": opted, "fill_pass_scatter_rays": int(st.scatterRays), "fill_pass_shadow_rays": int(st.shadowRays),
           "pixels_with_plane": [float((hd[p] != 0xFFFFFFFF).mean()) for p in range(3)], "pixels_with_non_primary_dominant_plane": float(((hd[3] & 3) != 0).mean()),
           "mean_radiance": float(r["merged"].mean()), "timing": "CUDA events around the whole rtxpt_b200_path_trace_realtime call (BUILD + FILL x sub_samples + merge), median over frames",
           "workload": "bench.py city workload, clear glass opted into path-space decomposition"}
    # realtime mode WITH the denoiser (BASELINE config 3's shape: stable planes + ReBLUR per plane + final merge).  The ReBLUR kernels had not run on a GPU when this was written:
    # any failure is reported in place and leaves the numbers above intact.
    try:
        k = sb.make_denoiser_constants(cam); trace_ms, dn_ms = [], []
        for f in range(args.warmup + args.frames):
            consts.sampleBaseIndex = 1000 + f * args.sub_samples; ctx.set_constants(consts)
            ctx.path_trace_realtime(False); ctx.synchronize(); t = float(ctx.stats().msTotal)
            ctx.denoise_realtime(k, sb.make_reblur_frame(cam, cam, frame_index=f)); d = ctx.last_denoise_ms()
            if f >= args.warmup: trace_ms.append(t); dn_ms.append(d)
        img = ctx.readback_output_color()[..., :3].astype(np.float32)
        px = W * H
        out["denoised"] = {"trace_ms": float(np.median(trace_ms)), "denoise_ms": float(np.median(dn_ms)), "denoise_ms_min": float(np.min(dn_ms)), "planes_denoised": 3,
                           "reblur_passes_per_plane": 8, "finite": bool(np.isfinite(img).all()), "mean_radiance": float(img.mean()),
                           "algorithmic_bytes_per_plane": int(px * (4 + 4 + 8 + 1 + 8 + 8 + 8 + 8 + 42 * 2)),
                           "note": "denoise_ms = CUDA events around rtxpt_b200_denoise_realtime (3 x { prepare inputs, 8 ReBLUR passes, final merge }); first GPU execution of these kernels"}
    except Exception as e:  # noqa: BLE001
        out["denoised"] = {"error": repr(e)[:300]}
    # realtime mode with NEE-AT temporal feedback (per-tile light samplers): first GPU execution of those kernels too; same containment as above
    try:
        consts.NEEATFeedback = 1; fb_ms, upd = [], []
        for f in range(args.warmup + args.frames + 8):                     # + 8: the loop needs a few frames before the tile samplers carry real feedback
            consts.sampleBaseIndex = 2000 + f * args.sub_samples; ctx.set_constants(consts)
            t0 = time.perf_counter(); ctx.neeat_update_begin(); ctx.synchronize(); t1 = time.perf_counter()
            ctx.path_trace_realtime(True); ctx.synchronize()
            if f >= args.warmup + 8: fb_ms.append(float(ctx.stats().msTotal)); upd.append((t1 - t0) * 1e3)
        ctl = ctx.neeat_raw(8, np.uint32, 8)
        img = ctx.readback_output_color()[..., :3].astype(np.float32)
        out["neeat_feedback"] = {"ms_per_frame": float(np.median(fb_ms)), "update_begin_ms_host_clock": float(np.median(upd)), "finite": bool(np.isfinite(img).all()), "mean_radiance": float(img.mean()),
                                 "pixels_with_feedback": float(ctl[7]) / float(W * H), "sampling_proxies": int(ctl[4]),
                                 "note": "ms_per_frame = CUDA events around rtxpt_b200_path_trace_realtime (BUILD + update_end's 4 passes + FILL with local candidates + merge); update_begin timed by the host clock around a synchronize"}
        consts.NEEATFeedback = 0; ctx.set_constants(consts)
    except Exception as e:  # noqa: BLE001
        out["neeat_feedback"] = {"error": repr(e)[:300]}
    # the callers either side of the path (§8f rows 3-4), first GPU executions as well: tone mapping, environment bake, BVH refit (last: a wrong refit would spoil what follows)
    try:
        tm = S.make_tone_mapping_params(op=5, auto_exposure=True); ctx.tone_map(tm); ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(10): ctx.tone_map(tm)
        ctx.synchronize(); tone_ms = (time.perf_counter() - t0) * 100.0
        ldr = ctx.readback_ldr()
        eq = np.random.default_rng(1).gamma(2.0, 0.5, (256, 512, 4)).astype(np.float32)
        t0 = time.perf_counter(); mips = ctx.bake_env_map(512, eq, lights=[((1.0, 0.9, 0.8), 10.0, (0.3, -0.8, 0.52), 0.05)]); bake_ms = (time.perf_counter() - t0) * 1e3
        n_inst = scene.desc.instanceCount
        xf = np.stack([np.array(scene.desc.instances[i].transform[:], np.float32).reshape(3, 4) for i in range(n_inst)])
        ctx.update_instance_transforms(xf); ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): ctx.update_instance_transforms(xf)
        ctx.synchronize(); refit_ms = (time.perf_counter() - t0) * 200.0
        consts.sampleBaseIndex = 0; ctx.set_constants(consts); ctx.path_trace_realtime(True); ctx.synchronize()
        after = float(ctx.readback_output_color()[..., :3].astype(np.float32).mean())
        out["callers"] = {"tone_map_ms_host_clock": tone_ms, "tone_map_avg_luminance": ctx.tone_map_average_luminance(), "ldr_mean": float(ldr[..., :3].mean()),
                          "env_bake_512_ms_host_clock_incl_copies": bake_ms, "env_bake_finite": bool(np.isfinite(mips[0]).all()),
                          "bvh_refit_ms_host_clock": refit_ms, "instances": int(n_inst), "bvh_intact_after_identity_refit": bool(abs(after - out["mean_radiance"]) < 0.25 * abs(out["mean_radiance"]) + 1e-6)}
    except Exception as e:  # noqa: BLE001
        out["callers"] = {"error": repr(e)[:300]}
    os.write(real_stdout, (json.dumps(out) + "\n").encode())         # before teardown: a device fault in the untested stage must not cost the line
    try: ctx.close()
    except Exception: pass  # noqa: BLE001
    return 0


if __name__ == "__main__":
    sys.exit(main())
