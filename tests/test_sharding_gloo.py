"""N>1 path on CPU: world_size-2 gloo processes shard a grid by point range,
each shades its strip (with the CPU oracle standing in for the device, since
there is no GPU here) and rank 0 gathers the strips; the result must equal the
single-process result.  Covers openshadinglanguage_b200/sharding.py."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers
from openshadinglanguage_b200.sharding import point_range


def test_point_ranges_partition_exactly():
    for n in (0, 1, 255, 256, 257, 1000, 4096 * 4096, 12345677):
        for world in (1, 2, 3, 4, 8):
            ranges = [point_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b and c <= d
            for b, e in ranges:
                if e > b:   # non-empty strips start on a CTA tile boundary
                    assert b % 256 == 0 and (e % 256 == 0 or e == n)


def _worker(rank, world, port, res, q):
    sys.path.insert(0, helpers.ROOT)
    sys.path.insert(0, os.path.join(helpers.ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import oracle
    from openshadinglanguage_b200.sharding import gather_strips, point_range as pr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = res * res
    b, e = pr(n, rank, world)
    layers, outputs, _ = helpers.image_case_group("noise")
    g = oracle.OracleGroup(layers, outputs=outputs)
    var, uni = oracle.testshade_globals(res, res)
    # this rank's strip of the SoA globals
    sl = {k: (np.asarray(v).reshape(-1, n)[:, b:e].copy() if np.asarray(v).size != n else np.asarray(v)[b:e].copy())
          for k, v in var.items()}
    out = np.zeros((e - b, 3), np.float32)
    g.run(e - b, sl, uni, out)
    full = gather_strips(torch.from_numpy(out), n, rank, world)
    if rank == 0:
        q.put(full.numpy())
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_process():
    import torch.multiprocessing as mp
    from oracle import oracle
    res = 96
    n = res * res
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, res, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    layers, outputs, _ = helpers.image_case_group("noise")
    g = oracle.OracleGroup(layers, outputs=outputs)
    var, uni = oracle.testshade_globals(res, res)
    want = np.zeros((n, 3), np.float32)
    g.run(n, var, uni, want)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_rank_tiles_partition_the_image():
    from openshadinglanguage_b200.sharding import rank_tiles, tile_pixel_index
    for (xres, yres) in ((128, 128), (160, 120), (100, 70)):
        for world in (1, 2, 3, 8):
            seen = np.zeros(xres * yres, np.int32)
            for r in range(world):
                idx = tile_pixel_index(rank_tiles(xres, yres, r, world), xres).numpy()
                seen[idx] += 1
            assert (seen == 1).all()          # every pixel in exactly one work set


def _tile_worker(rank, world, port, xres, yres, q):
    sys.path.insert(0, helpers.ROOT)
    import torch
    import torch.distributed as dist
    from openshadinglanguage_b200.sharding import gather_tiles, rank_tiles, tile_pixel_index
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(xres * yres * 3, dtype=torch.float32).reshape(xres * yres, 3)   # the "frame"
    strip = full[tile_pixel_index(rank_tiles(xres, yres, rank, world), xres)]           # this rank's render
    img = gather_tiles(strip, xres, yres, rank, world)
    if rank == 0:
        q.put(img.numpy())
    dist.destroy_process_group()


def test_two_rank_tile_gather_reassembles_the_frame():
    import torch.multiprocessing as mp
    xres, yres = 160, 120          # ragged: 64 divides neither
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tile_worker, args=(r, 2, port, xres, yres, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(xres * yres * 3, dtype=np.float32).reshape(yres, xres, 3)
    assert np.array_equal(got, want)
