"""CPU tests of the multi-GPU host logic with world_size 2 over gloo: stream sharding (weak and strong), chunk-range
sharding of one long stream, and the max/sum reductions bench.py uses.  The data path itself has no collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from sea_codec_b200 import dist
from util import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sea_path, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import sea_oracle as O
    from sea_codec_b200 import dist as D

    info = D.init("gloo")
    assert (info.rank, info.world) == (rank, world)
    # reductions used for timing (max over ranks) and counters (sum over ranks)
    assert D.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    assert D.sum_over_ranks(rank + 1) == world * (world + 1) / 2
    D.barrier()
    # chunk-range sharding of ONE stream (config 3 style): each rank decodes its own chunk range, no exchange
    sea = open(sea_path, "rb").read()
    chunk_size = sea[6] | (sea[7] << 8)
    n_chunks = (len(sea) - 22 + chunk_size - 1) // chunk_size
    b, e = D.shard_chunks(n_chunks, rank, world)
    part = O.sea_decode(D.chunk_range_as_file(sea, b, e)).samples
    q.put((rank, b, e, part))
    total = D.sum_over_ranks(part.size)
    q.put((rank, "total", total, None))
    D.barrier()
    D.shutdown()


def test_shard_helpers():
    for n in (0, 1, 7, 4096, 1024):
        for world in (1, 2, 3, 8):
            ranges = [dist.shard_streams(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert [dist.weak_streams(4096, r) for r in range(3)] == [(0, 4096), (4096, 8192), (8192, 12288)]


def test_two_ranks_gloo(oracle, tmp_path):
    from sea_codec_b200 import synth

    pcm = synth.gen_stream(9, 5120 * 7 + 333, 3, 48000)
    sea = oracle.sea_encode(pcm, 48000, 3, oracle.make_settings(4.0))
    whole = oracle.sea_decode(sea).samples
    path = tmp_path / "x.sea"
    path.write_bytes(sea)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(path), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2 * world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = sorted([r for r in results if r[1] != "total"], key=lambda r: r[0])
    assert parts[0][1] == 0 and parts[0][2] == parts[1][1]
    assert np.array_equal(np.concatenate([p[3] for p in parts]), whole)
    for r in results:
        if r[1] == "total":
            assert r[2] == whole.size
