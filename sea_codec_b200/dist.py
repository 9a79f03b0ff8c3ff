"""Multi-GPU plumbing.  The SEA hot path shards by independent streams: every rank encodes/decodes its own
streams, nothing is exchanged on the data path, and only per-rank counters and timings are reduced on the host side
(torch.distributed: NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Tuple


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def rank_info() -> RankInfo:
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str | None = None) -> RankInfo:
    """Joins the process group when launched under torchrun (WORLD_SIZE > 1); no-op for a single process."""
    info = rank_info()
    if info.world > 1:
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(info.local_rank)
                kw["device_id"] = torch.device("cuda", info.local_rank)  # binds the communicator to this rank's GPU up front
            dist.init_process_group(backend=backend, rank=info.rank, world_size=info.world, **kw)
    return info


def _reduce(value: float, op_name: str) -> float:
    info = rank_info()
    if info.world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def max_over_ranks(value: float) -> float:
    return _reduce(value, "MAX")


def sum_over_ranks(value: float) -> float:
    return _reduce(value, "SUM")


def barrier() -> None:
    if rank_info().world > 1:
        import torch.distributed as dist

        dist.barrier()


def shutdown() -> None:
    if rank_info().world > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


def shard_streams(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Strong-scaling split of n_total independent streams: contiguous [begin, end) per rank, sizes differ by <= 1."""
    base, extra = divmod(n_total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def weak_streams(n_per_rank: int, rank: int) -> Tuple[int, int]:
    """Weak-scaling assignment: every rank owns n_per_rank streams with globally unique ids."""
    return rank * n_per_rank, (rank + 1) * n_per_rank


def shard_chunks(n_chunks: int, rank: int, world: int) -> Tuple[int, int]:
    """Decode of ONE long stream shards by chunk range: chunk k starts at byte 22 + k*chunk_size (file.rs:185) and writes
    PCM at frame k*frames_per_chunk, so ranks need no communication either."""
    return shard_streams(n_chunks, rank, world)


def chunk_range_as_file(sea: bytes, begin: int, end: int) -> bytes:
    """A rank's chunk range [begin, end) of a .sea file, re-wrapped with a header so sea_decode can run on it."""
    import struct

    channels = sea[5]
    chunk_size, fpc = struct.unpack_from("<HH", sea, 6)
    rate, total = struct.unpack_from("<II", sea, 10)
    body = sea[22 + begin * chunk_size: 22 + end * chunk_size]
    frames = max(0, min(total, end * fpc) - begin * fpc) if total else 0
    return b"seac" + struct.pack("<BBHHIII", 1, channels, chunk_size, fpc, rate, frames, 0) + body
