"""Multi-GPU tile split on real GPUs (run under torchrun, one rank per GPU): every rank traces its tiles, one NCCL all-gather exchanges the
radiance, and rank 0 checks the assembled frame bit-for-bit against the same frame traced by a single-GPU context."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from rtxpt_b200 import lib, scenes, scene_builder as sb


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h, spp = 640, 360, 4
    scene, cam = scenes.city_block(target_triangles=200000, width=w, height=h, seed=5)
    consts = sb.make_constants(w, h, cam, bounce_count=4, diffuse_bounce_count=4, env_enabled=True, firefly_threshold=5000.0, nee=True, nee_type=2)
    ctx = lib.Context(max_sub_samples_per_launch=spp, device=local, tile_rank=rank, tile_world=world, tile_size=64)
    ctx.upload_scene(scene); ctx.set_constants(consts)
    owned, padded = ctx.tile_layout()
    send = torch.empty((padded, 4), dtype=torch.float32, device="cuda")
    gathered = torch.empty((world * padded, 4), dtype=torch.float32, device="cuda")
    tstream = torch.cuda.Stream(); torch.cuda.set_stream(tstream)        # one non-default stream for the kernels, the tile copies and NCCL (0 would mean "the context's own stream")
    stream = tstream.cuda_stream
    for frame in range(2):                      # two accumulated frames
        consts.sampleBaseIndex = frame * spp; ctx.set_constants(consts)
        ctx.path_trace(0, spp, True, stream)
        ctx.pack_owned(send.data_ptr(), stream)
        dist.all_gather_into_tensor(gathered, send)
        ctx.unpack_all(gathered.data_ptr(), stream)
    torch.cuda.synchronize()
    img = ctx.readback_accumulated()
    ok = True
    if rank == 0:
        one = lib.Context(max_sub_samples_per_launch=spp, device=local)
        one.upload_scene(scene)
        for frame in range(2):
            consts.sampleBaseIndex = frame * spp; one.set_constants(consts)
            one.path_trace(0, spp, True)
        one.synchronize()
        ref = one.readback_accumulated()
        same = np.array_equal(img, ref)
        print(f"[multi-gpu] world={world} owned={owned} padded={padded} frame bit-identical to single-GPU frame: {same}; max abs diff {np.abs(img - ref).max():.3g}", flush=True)
        ok = same
        one.close()
    # every rank must hold the same assembled frame
    t = torch.from_numpy(img).cuda(); ref0 = t.clone(); dist.broadcast(ref0, 0)
    eq = torch.tensor([int(torch.equal(t, ref0))], device="cuda"); dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    if rank == 0: print(f"[multi-gpu] all ranks hold the same frame: {bool(eq.item())}", flush=True)
    ctx.close(); dist.destroy_process_group()
    return 0 if (ok and bool(eq.item())) else 1


if __name__ == "__main__":
    sys.exit(main())
