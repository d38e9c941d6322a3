"""BASELINE configs[2] on N GPUs: realtime mode, 4 sub-samples, NEE-AT feedback (>= 10 k emissive triangles, 32 warm-up frames), ReBLUR on 3 stable planes, tone map -
one rank per GPU under torchrun, the frame recipe of rtxpt_b200/realtime_mgpu.py.  Prints ONE JSON line on rank 0: frame time (CUDA events on the launching stream, max over
ranks), the split trace / exchange + denoise, bytes all-gathered per frame.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/bench_config3_mgpu.py [--frames 10]
    python scripts/bench_config3_mgpu.py            (N = 1: the same code path with a one-rank group)"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=10); ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--width", type=int, default=1920); ap.add_argument("--height", type=int, default=1080); ap.add_argument("--triangles", type=int, default=2_800_000)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    real_stdout = os.dup(1); os.dup2(2, 1)
    import torch
    from rtxpt_b200 import lib, scenes, scene_builder as sb, structs as S, realtime_mgpu as M
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, SPP = args.width, args.height, 4
    scene, cam = scenes.city_block(target_triangles=args.triangles, width=W, height=H, delta_surfaces=True)
    consts = sb.make_constants(W, H, cam, bounce_count=6, diffuse_bounce_count=6, env_enabled=True, firefly_threshold=5000.0, nee=True, nee_type=2); consts.NEEATFeedback = 1
    ctx = lib.Context(max_sub_samples_per_launch=1, device=local, tile_rank=rank, tile_world=world, tile_size=64)
    ctx.upload_scene(scene); ctx.set_constants(consts); ctx.set_view(sb.world_to_clip(cam)); ctx.set_realtime(sb.make_realtime_constants(W, H, cam, bounce_count=6, sub_samples=SPP))
    k = sb.make_denoiser_constants(cam); tm = S.make_tone_mapping_params(op=5, auto_exposure=True)
    tstream = torch.cuda.Stream(); torch.cuda.set_stream(tstream)
    group = M.DistGroup(ctx, tstream) if world > 1 else M.LocalGroup([ctx])
    if world == 1: group.stream = tstream.cuda_stream
    def barrier():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]; frame_ms, trace_ms, moved = [], [], 0
    for f in range(args.warmup + args.frames):
        consts.sampleBaseIndex = f * SPP; ctx.set_constants(consts)
        frame = sb.make_reblur_frame(cam, cam, frame_index=f, frame_time_ms=16.0)
        barrier(); ev[0].record()
        ctx.neeat_update_begin(tstream.cuda_stream); ctx.path_trace_realtime(False, tstream.cuda_stream); ev[1].record()
        moved = group.exchange(M.GUIDES); ctx.denoise_spec_hit_t(tstream.cuda_stream); first = True
        for plane in (2, 1, 0):
            ctx.denoiser_prepare_inputs(plane, first, k, tstream.cuda_stream); moved += group.exchange(M.NRD_INPUTS)
            ctx.reblur_denoise(plane, frame, tstream.cuda_stream); ctx.denoiser_final_merge(plane, stream=tstream.cuda_stream, identity=False); first = False
        moved += group.exchange(M.OUTPUT); ctx.tone_map(tm, stream=tstream.cuda_stream); ev[2].record()
        barrier()
        if f >= args.warmup: frame_ms.append(ev[0].elapsed_time(ev[2])); trace_ms.append(ev[0].elapsed_time(ev[1]))
    t = torch.tensor([float(np.median(frame_ms)), float(np.median(trace_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    else: tmax = tmin = t
    img = ctx.readback_output_color()[..., :3].astype(np.float32); ldr = ctx.readback_ldr()
    if rank == 0:
        out = {"workload": "BASELINE configs[2]: city with delta surfaces %dx%d, 4 sub-samples, NEE-AT feedback, ReBLUR x 3 planes, tone map" % (W, H), "n_gpus": world, "frames": args.frames, "warmup_frames": args.warmup,
               "frame_ms": float(tmax[0]), "trace_ms_max_rank": float(tmax[1]), "trace_ms_min_rank": float(tmin[1]), "exchange_denoise_tonemap_ms": float(tmax[0] - tmax[1]),
               "all_gathered_bytes_per_frame": int(moved), "finite": bool(np.isfinite(img).all()), "mean_radiance": float(img.mean()), "ldr_mean": float(ldr[..., :3].mean()),
               "partition": "1 GPU" if world == 1 else "interleaved 64x64 screen tiles over %d GPUs; all-gathers per frame: guides (depth, spec hit distance, plane neighbour guides), NRD inputs x 3 planes, output colour; ReBLUR replicated; NEE-AT adapts per rank" % world,
               "timing": "CUDA events on the launching stream around one frame, median over the timed frames, max over ranks"}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    ctx.close()
    if world > 1: dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
