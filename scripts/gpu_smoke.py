"""Quick GPU bring-up: Cornell box through the product vs the oracle (primary hits, lights, 1-spp image)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from rtxpt_b200 import lib, scenes, scene_builder as sb, structs as S
import oracle_lib as ol

W = H = int(os.environ.get("SMOKE_RES", "256"))
scene, cam = scenes.cornell_box(W, H)
consts = sb.make_constants(W, H, cam, bounce_count=2, diffuse_bounce_count=2)
ctx = lib.Context(max_sub_samples_per_launch=4, flags=S.CFG_COUNT_TRAVERSAL_STEPS)
t = time.time(); ctx.upload_scene(scene); print("upload %.3fs" % (time.time() - t))
ctx.set_constants(consts)
o = ol.Oracle(scene); o.set_constants(consts)

# lights
li_p, ct_p, px_p = ctx.lights(); li_o, ct_o, px_o = o.lights()
print("lights equal:", np.array_equal(li_p, li_o), np.array_equal(ct_p, ct_o), np.array_equal(px_p, px_o), li_p.shape, px_p.shape)

# random rays
rng = np.random.default_rng(1)
n = 200000
org = rng.uniform([0.2, 0.2, -3], [5.3, 5.3, 5.3], (n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.concatenate([org, np.zeros((n, 1), np.float32), d, np.full((n, 1), 1e15, np.float32)], 1).astype(np.float32)
hp = ctx.trace_rays(rays); ho = o.trace_rays(rays)
same = (hp["t"].view(np.uint32) == ho["t"].view(np.uint32)) & (hp["prim"] == ho["prim"]) & (hp["inst"] == ho["inst"]) & (hp["u"].view(np.uint32) == ho["u"].view(np.uint32))
print("closest-hit bit-exact: %d / %d ; hit fraction %.3f" % (same.sum(), n, (ho["t"] >= 0).mean()))
if not same.all():
    bad = np.nonzero(~same)[0][:5]
    for b in bad: print("  ray", b, hp[b], ho[b])

# image, 1 spp
ctx.path_trace(0, 1); img = ctx.readback_accumulated(); st = ctx.stats()
acc, _, last, _, ost = o.render(0, 1)
diff = np.abs(img[..., :3] - acc[..., :3])
rel = diff / (np.abs(acc[..., :3]) + 1e-3)
print("image: product mean", img[..., :3].mean((0, 1)), "oracle mean", acc[..., :3].mean((0, 1)))
print("pixels exactly equal: %.4f ; rel err < 1e-2: %.4f ; max abs diff %.4f" % ((diff.max(-1) == 0).mean(), (rel.max(-1) < 1e-2).mean(), diff.max()))
print("rays: product scatter %d shadow %d | oracle scatter %d shadow %d" % (st.scatterRays, st.shadowRays, ost.scatterRays, ost.shadowRays))
print("ms %.3f launches %d nodes/ray %.1f tris/ray %.1f" % (st.msTotal, st.kernelLaunches, st.traversalNodeVisits / max(1, st.scatterRays + st.shadowRays), st.traversalTriTests / max(1, st.scatterRays + st.shadowRays)))
# 16 spp timing
ctx.reset_accumulation(); ctx.path_trace(0, 16); ctx.synchronize(); st = ctx.stats()
print("16spp: %.3f ms, %.1f Mrays/s" % (st.msTotal, (st.scatterRays + st.shadowRays) / st.msTotal / 1e3))
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/cornell_product.npy", ctx.readback_accumulated())
