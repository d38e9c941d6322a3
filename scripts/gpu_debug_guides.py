import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, oracle_lib
from rtxpt_b200 import lib, scenes, scene_builder as sb, structs as S
scene, cam = scenes.city_block(target_triangles=60000, width=320, height=180, seed=3, texture_size=64, n_textures=8, n_materials=40)
W, H = 320, 180
for bounces in (0, 1, 3):
    consts = sb.make_constants(W, H, cam, bounce_count=bounces, diffuse_bounce_count=bounces, env_enabled=True)
    m = sb.world_to_clip(cam)
    c = lib.Context(max_sub_samples_per_launch=1, strict=True, flags=S.CFG_EXPORT_GUIDES)
    c.upload_scene(scene); c.set_constants(consts); c.set_view(m); c.path_trace(0, 1, True); c.synchronize()
    depth, mv, thp = c.readback_guides(); c.close()
    o = oracle_lib.Oracle(scene); o.set_constants(consts); o.set_view(m); od, ot = o.render_guides(0); o.close()
    bad = depth != od
    print("bounces", bounces, "depth mismatch", bad.mean(), "thp mismatch", (thp != ot).mean(), "both", (bad & (thp != ot)).mean())
    if bad.any():
        d = np.abs(depth[bad] - od[bad]); print("  abs diff percentiles", np.percentile(d, [10, 50, 90, 99, 100]), " oracle==1:", (od[bad] == 1).mean(), " gpu==1:", (depth[bad] == 1).mean(), "gpu==0", (depth[bad] == 0).mean())
        ys, xs = np.nonzero(bad); print("  sample", [(int(x), int(y), float(depth[y, x]), float(od[y, x])) for x, y in list(zip(xs, ys))[:6]])
