"""Workload for `ncu` captures of the ReBLUR passes (profiles/): BASELINE configs[2]'s frame on the city with delta surfaces at 1920x1080 - realtime trace (1 sub-sample, enough to
feed the denoiser) then rtxpt_b200_denoise_realtime, 8 frames so that the history is warm.  24 ReBLUR launches per frame (3 planes x 8 passes; plane 0, the full frame, is the last 8).
    ncu --set full --clock-control none --import-source on -k regex:k_rb_ -s 168 -c 24 -o gpurun_out/r2_reblur python scripts/profile_reblur.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtxpt_b200 import lib, scenes, scene_builder as sb

W, H = 1920, 1080
scene, cam = scenes.city_block(width=W, height=H, delta_surfaces=True)
consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=6, env_enabled=True, firefly_threshold=5000.0, nee=True, nee_type=2)
ctx = lib.Context(max_sub_samples_per_launch=1)
ctx.upload_scene(scene); ctx.set_constants(consts); ctx.set_view(sb.world_to_clip(cam))
ctx.set_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=1))
k = sb.make_denoiser_constants(cam)
ms = []
for f in range(int(os.environ.get("FRAMES", "8"))):
    consts.sampleBaseIndex = f; ctx.set_constants(consts)
    ctx.path_trace_realtime(False)
    ctx.denoise_realtime(k, sb.make_reblur_frame(cam, cam, frame_index=f, frame_time_ms=16.0)); ms.append(ctx.last_denoise_ms())
print("denoise ms per frame:", ["%.3f" % m for m in ms])
ctx.close()
