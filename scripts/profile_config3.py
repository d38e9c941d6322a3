"""Workload for the ncu launch list of BASELINE configs[2]'s frame (profiles/): city with delta surfaces, 1920x1080, NEE-AT feedback, 4 sub-samples, ReBLUR x 3 planes, tone map.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_config3_launches.csv python scripts/profile_config3.py
FRAMES (default 6) frames; the last one is the one to read (the caches are warm by then; 32 frames are what the bench uses, a launch list does not need them)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtxpt_b200 import lib, scenes, scene_builder as sb, structs as S

W, H, SPP = 1920, 1080, 4
scene, cam = scenes.city_block(width=W, height=H, delta_surfaces=True)
consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=6, env_enabled=True, firefly_threshold=5000.0, nee=True, nee_type=2)
consts.NEEATFeedback = int(os.environ.get("NEEAT", "1"))
ctx = lib.Context(max_sub_samples_per_launch=1)
ctx.upload_scene(scene); ctx.set_constants(consts); ctx.set_view(sb.world_to_clip(cam))
ctx.set_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=SPP))
k = sb.make_denoiser_constants(cam); tm = S.make_tone_mapping_params(op=5, auto_exposure=True)
for f in range(int(os.environ.get("FRAMES", "6"))):
    consts.sampleBaseIndex = f * SPP; ctx.set_constants(consts)
    if consts.NEEATFeedback: ctx.neeat_update_begin()
    ctx.path_trace_realtime(False); ctx.synchronize(); t = ctx.stats().msTotal
    ctx.denoise_spec_hit_t(); ctx.denoise_realtime(k, sb.make_reblur_frame(cam, cam, frame_index=f, frame_time_ms=16.0)); d = ctx.last_denoise_ms()
    ctx.tone_map(tm); ctx.synchronize()
    print("frame %d: trace %.3f ms, denoise %.3f ms" % (f, t, d))
ctx.close()
