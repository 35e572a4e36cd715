"""Tile-parallel sharding across GPUs.

Every tile (its own h0, omega and output) is a pure function of (h0, omega, t), so the path shards
with no exchange step: global tile i goes to rank floor(i * G / T) in contiguous blocks, each rank
runs the same two-kernel frame over its local tiles, and no collective touches the data path. The
only communication is control plane: the launch fan-out (rank 0 broadcasts the frame parameter block), the
barrier, the max of the elapsed time and the gathered checksums."""
from __future__ import annotations


def tiles_of_rank(rank: int, world: int, n_tiles: int) -> list[int]:
    """Global tile indices owned by `rank`: contiguous blocks, sizes differing by at most one."""
    if not (0 <= rank < world) or n_tiles < 0:
        raise ValueError("bad rank/world/n_tiles")
    return list(range((rank * n_tiles) // world, ((rank + 1) * n_tiles) // world))


def rank_of_tile(tile: int, world: int, n_tiles: int) -> int:
    for r in range(world):
        if tile in tiles_of_rank(r, world, n_tiles):
            return r
    raise ValueError("tile out of range")


def checksum(out) -> float:
    """Order-independent per-tile checksum used to verify a sharded run against a single-GPU run."""
    import numpy as np
    a = np.asarray(out, dtype=np.float64)
    return float(np.abs(a).sum())


def fan_out(first_frame: int, n_frames: int, dt: float, device=None, src: int = 0) -> tuple[int, int, float]:
    """Launch fan-out: rank `src` broadcasts the frame parameter block -- (first frame id, number of
    frames, time step) -- and every rank enqueues exactly those frames on its own tiles. One broadcast per block of
    frames, not per frame: the frames of a block need no further communication. Identity without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(first_frame), int(n_frames), float(dt)
    block = torch.tensor([float(first_frame), float(n_frames), float(dt)], dtype=torch.float64, device=device or "cpu")
    dist.broadcast(block, src=src)
    first, count, step = block.tolist()
    return int(first), int(count), float(step)
