"""CPU: the N>1 path's host logic with world_size 2 over gloo — tile partition, pack -> all_gather -> unpack, and the bench launcher contract."""
import json
import os
import subprocess
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_partition_covers_every_pixel_once():
    from rtxpt_b200 import tiles
    for (w, h, t, world) in [(1920, 1080, 64, 8), (320, 180, 32, 2), (257, 131, 64, 3), (64, 64, 64, 4)]:
        tables, padded = tiles.gather_layout(w, h, t, world)
        allp = np.concatenate(tables)
        assert len(allp) == w * h and len(np.unique(allp)) == w * h
        assert max(len(x) for x in tables) == padded
        if world <= (w + t - 1) // t * ((h + t - 1) // t):
            assert min(len(x) for x in tables) > 0
        # 32 consecutive slots form an 8x4 pixel block inside full tiles (coherent primary-ray warps)
        blk = tables[0][:32]
        assert (blk >> 16).max() - (blk >> 16).min() == 7 and (blk & 0xFFFF).max() - (blk & 0xFFFF).min() == 3


def _worker(rank, world, port, w, h, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from rtxpt_b200 import tiles
    yy, xx = np.mgrid[0:h, 0:w]
    truth = np.stack([xx * 1.0, yy * 2.0, xx * yy * 0.001, np.ones_like(xx, dtype=float)], -1).astype(np.float32)
    tables, padded = tiles.gather_layout(w, h, tile, world)
    mine = np.zeros_like(truth)
    x, y = tables[rank] >> 16, tables[rank] & 0xFFFF
    mine[y, x] = truth[y, x]                                    # this rank only rendered its own tiles
    send = torch.from_numpy(tiles.pack_owned(mine, tables[rank], padded))
    gathered = torch.empty((world * padded, 4), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, send)
    full = tiles.unpack_all(gathered.numpy(), tables, padded, mine.copy())
    ok = np.array_equal(full, truth)
    # whole-job throughput bookkeeping: sum of per-rank ray counts, max of per-rank times
    t = torch.tensor([float(rank + 1), 10.0 * (rank + 1)], dtype=torch.float64)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    q.put((rank, ok, float(tmax[0]), float(tsum[1])))
    dist.destroy_process_group()


def test_gather_reassembles_frame_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29613, 200, 120, 32, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(60)
    assert all(r[1] for r in res)
    assert all(r[2] == 2.0 and r[3] == 30.0 for r in res)


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _exchange_worker(rank, world, port, w, h, tile, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from rtxpt_b200 import tiles
    rng = np.random.default_rng(11)                                  # every rank builds the same "truth"; it only owns its tiles of it
    truth = [rng.random((h, w)).astype(np.float32), rng.integers(0, 255, (h, w), dtype=np.uint8), rng.random((h, w, 4)).astype(np.float16), rng.integers(0, 2 ** 32 - 1, (h, w), dtype=np.uint32)]
    tables, padded = tiles.gather_layout(w, h, tile, world)
    x, y = tables[rank] >> 16, tables[rank] & 0xFFFF
    mine = [np.zeros_like(a) for a in truth]
    for m, a in zip(mine, truth): m[y, x] = a[y, x]
    send = torch.from_numpy(tiles.exchange_pack(mine, tables[rank], padded))
    gathered = torch.empty(world * send.numel(), dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, send)
    tiles.exchange_unpack(gathered.numpy(), tables, padded, mine, skip_rank=rank)
    q.put((rank, all(np.array_equal(m, a) for m, a in zip(mine, truth)), send.numel()))
    dist.destroy_process_group()


def test_realtime_frame_exchange_layout_world2_gloo():
    """The generic per-pixel image exchange of the realtime frame (rtxpt_b200_exchange_pack / _unpack; tiles.exchange_* is its host mirror): images of 4, 1, 8 and 4 bytes per
    pixel, packed tile-wise per rank, one all-gather, scattered back - every rank ends up with the whole of every image."""
    from rtxpt_b200 import tiles
    ctx = mp.get_context("spawn"); q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, 29617, 200, 120, 32, q)) for r in range(2)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(60)
    assert all(r[1] for r in res)
    _, padded = tiles.gather_layout(200, 120, 32, 2)
    offs, total = tiles.exchange_layout([4, 1, 8, 4], padded)
    assert res[0][2] == total and all(o % 16 == 0 for o in offs)
