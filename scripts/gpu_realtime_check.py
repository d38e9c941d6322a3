"""One-shot diagnosis of realtime mode on the GPU against the oracle (test infrastructure): prints how many header words / plane fields / merged pixels
agree for the strict and the default build.  Meant for a single short gpurun call; the assertions live in tests/test_gpu_realtime.py."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from rtxpt_b200 import lib as P, scene_builder as sb, scenes

W = H = 96
scene, cam = scenes.cornell_box(W, H, delta_surfaces=True)
consts = sb.make_constants(W, H, cam, bounce_count=8, diffuse_bounce_count=3)
o = O.Oracle(scene); o.set_constants(consts); o.set_view(sb.world_to_clip(cam))
rt = sb.make_realtime_constants(W, H, cam, bounce_count=8, sub_samples=2)
t0 = time.time(); r = o.render_realtime(rt); print("oracle %.2fs" % (time.time() - t0), flush=True)
ys, xs = np.mgrid[0:H, 0:W]
for strict in (True, False):
    try:
        c = P.Context(max_sub_samples_per_launch=1, strict=strict); c.upload_scene(scene); c.set_constants(consts); c.set_view(sb.world_to_clip(cam))
        c.set_realtime(rt); t0 = time.time(); c.path_trace_realtime(True); c.synchronize(); print("strict" if strict else "fast", "gpu %.3fs" % (time.time() - t0), flush=True)
        g = c.readback_realtime()
        for layer in range(4): print("  header layer", layer, "equal", float((g["header"][layer] == r["header"][layer]).mean()), "gpu ids", np.unique(g["header"][layer])[:8].tolist(), flush=True)
        same = (g["header"] == r["header"]).all(0)
        for plane in range(3):
            v = same & (r["header"][plane] != 0xFFFFFFFF)
            if not v.any(): continue
            a = g["planes"][sb.generic_ts_address(xs[v], ys[v], plane, W, H)]; b = r["planes"][sb.generic_ts_address(xs[v], ys[v], plane, W, H)]
            for f in a.dtype.names:
                print("  plane", plane, f, "bit-equal", float((a[f] == b[f]).mean()) if a[f].dtype.kind == "u" else float(np.isclose(a[f], b[f], rtol=1e-6, atol=1e-6, equal_nan=True).mean()), flush=True)
        for k in ("stable_radiance", "depth", "motion", "throughput", "spec_hit_t"):
            print("  ", k, "equal", float((g[k] == r[k]).mean()), "close", float(np.isclose(g[k].astype(np.float64), r[k].astype(np.float64), rtol=2e-3, atol=1e-3).mean()), flush=True)
        d = np.abs(g["merged"] - r["merged"])
        print("  merged equal", float((d == 0).all(-1).mean()), "mean gpu/oracle", float(g["merged"].mean()), float(r["merged"].mean()), "max diff", float(d.max()), flush=True)
        c.path_trace_realtime(True); c.synchronize(); g2 = c.readback_realtime()
        print("  deterministic", all(g[k].tobytes() == g2[k].tobytes() for k in g), flush=True)
        c.close()
    except Exception as e:
        print("FAILED", strict, repr(e), flush=True)
