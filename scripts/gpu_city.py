"""City stand-in through product vs oracle at reduced size: lights, random rays (closest + any-hit incl. alpha test), images."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rtxpt_b200 import lib, scenes, scene_builder as sb, structs as S
from rtxpt_b200.imageio import per_pixel_l2
import oracle_lib as ol

W, H = 480, 270
scene, cam = scenes.city_block(target_triangles=300000, width=W, height=H, texture_size=256, n_textures=8, n_materials=96)
consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=6, env_enabled=True, firefly_threshold=5000.0)
ctx = lib.Context(max_sub_samples_per_launch=4, flags=S.CFG_COUNT_TRAVERSAL_STEPS | S.CFG_TIME_KERNELS)
t = time.time(); ctx.upload_scene(scene); print("upload %.3fs tris %d" % (time.time() - t, scene.triangle_count))
t = time.time(); ctx.set_constants(consts); print("set_constants %.3fs" % (time.time() - t))
t = time.time(); o = ol.Oracle(scene); o.set_constants(consts); print("oracle setup %.3fs" % (time.time() - t))
li_p, ct_p, px_p = ctx.lights(); li_o, ct_o, px_o = o.lights()
print("lights equal:", np.array_equal(li_p, li_o), np.array_equal(ct_p, ct_o), np.array_equal(px_p, px_o), li_p.shape, px_p.shape)
if not np.array_equal(li_p, li_o):
    bad = np.nonzero((li_p != li_o).any(1))[0]; print("  differing lights:", len(bad), bad[:10])
rng = np.random.default_rng(3)
n = 400000
org = rng.uniform([-100, 0.2, -100], [100, 30, 100], (n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, np.zeros((n, 1), np.float32), d, np.full((n, 1), 1e15, np.float32)], 1).astype(np.float32)
for any_hit in (False, True):
    hp = ctx.trace_rays(rays, any_hit); ho = o.trace_rays(rays, any_hit)
    if any_hit:
        same = (hp["t"] >= 0) == (ho["t"] >= 0)
    else:
        same = (hp["t"].view(np.uint32) == ho["t"].view(np.uint32)) & (hp["prim"] == ho["prim"]) & (hp["inst"] == ho["inst"]) & (hp["geom"] == ho["geom"]) & (hp["u"].view(np.uint32) == ho["u"].view(np.uint32)) & (hp["v"].view(np.uint32) == ho["v"].view(np.uint32))
    print("anyHit=%d agree: %d / %d ; hit fraction %.3f" % (any_hit, same.sum(), n, (ho["t"] >= 0).mean()))
    if not same.all():
        for b in np.nonzero(~same)[0][:5]: print("  ray", b, hp[b], ho[b])
ctx.path_trace(0, 1); img = ctx.readback_accumulated(); st = ctx.stats()
acc, _, last, _, ost = o.render(0, 1)
diff = np.abs(img[..., :3] - acc[..., :3]); rel = diff / (np.abs(acc[..., :3]) + 1e-2)
print("1spp: product mean", img[..., :3].mean((0, 1)), "oracle mean", acc[..., :3].mean((0, 1)))
print("1spp: pixels exactly equal %.4f ; rel<1e-2 %.4f ; rel<5e-2 %.4f ; L2 %.3e" % ((diff.max(-1) == 0).mean(), (rel.max(-1) < 1e-2).mean(), (rel.max(-1) < 5e-2).mean(), per_pixel_l2(img, acc)))
print("rays product %d+%d oracle %d+%d" % (st.scatterRays, st.shadowRays, ost.scatterRays, ost.shadowRays))
ctx.reset_accumulation(); ctx.path_trace(0, 32); img = ctx.readback_accumulated(); st = ctx.stats()
acc, _, _, _, ost = o.render(0, 32)
print("32spp: L2 %.3e  mean abs diff %.3e  means %s %s" % (per_pixel_l2(img, acc), np.abs(img[..., :3] - acc[..., :3]).mean(), img[..., :3].mean((0, 1)), acc[..., :3].mean((0, 1))))
print("32spp GPU %.2f ms (%.1f Mrays/s) closest %.2f shadow %.2f shade %.2f other %.2f | oracle %.2fs (%.2f Mrays/s, %d threads)" % (st.msTotal, (st.scatterRays + st.shadowRays) / st.msTotal / 1e3,
      st.msTraceClosest, st.msTraceShadow, st.msShade, st.msOther, ost.seconds, (ost.scatterRays + ost.shadowRays) / ost.seconds / 1e6, ost.threads))
print("nodes/ray %.1f tris/ray %.1f shadow nodes/ray %.1f" % (st.traversalNodeVisits / max(1, st.scatterRays), st.traversalTriTests / max(1, st.scatterRays), st.shadowNodeVisits / max(1, st.shadowRays)))
os.makedirs("gpurun_out", exist_ok=True)
from rtxpt_b200.imageio import write_png
write_png("gpurun_out/city_product.png", img); write_png("gpurun_out/city_oracle.png", acc)
