"""Reader/writer for the reference's on-disk inputs.

``data/omega.bin`` is a bincode-1.3.1 ``Vec<f32>`` and ``data/spectrum.bin`` a
``Vec<[f32; 2]>`` (decoded at /root/reference/src/render.rs:769-771 and :808-810;
bincode pinned at Cargo.lock:92-93): a little-endian u64 element count followed by
the raw little-endian f32 payload. Fixed-size arrays carry no extra prefix.
"""
from __future__ import annotations

import struct

import numpy as np


def read_vec_f32(path: str) -> np.ndarray:
    """``Vec<f32>`` -> float32[count]."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 8:
        raise ValueError(f"{path}: too short for a bincode length prefix")
    (count,) = struct.unpack_from("<Q", raw, 0)
    if len(raw) != 8 + 4 * count:
        raise ValueError(f"{path}: {len(raw)} bytes, expected {8 + 4 * count} for Vec<f32>[{count}]")
    return np.frombuffer(raw, "<f4", count, 8).copy()


def read_vec_f32x2(path: str) -> np.ndarray:
    """``Vec<[f32; 2]>`` -> float32[count, 2]."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) < 8:
        raise ValueError(f"{path}: too short for a bincode length prefix")
    (count,) = struct.unpack_from("<Q", raw, 0)
    if len(raw) != 8 + 8 * count:
        raise ValueError(f"{path}: {len(raw)} bytes, expected {8 + 8 * count} for Vec<[f32;2]>[{count}]")
    return np.frombuffer(raw, "<f4", 2 * count, 8).reshape(count, 2).copy()


def write_vec_f32(path: str, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a, "<f4").reshape(-1)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", a.size))
        f.write(a.tobytes())


def write_vec_f32x2(path: str, a: np.ndarray) -> None:
    a = np.ascontiguousarray(a, "<f4").reshape(-1, 2)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", a.shape[0]))
        f.write(a.tobytes())


def resolution_of(count: int) -> int:
    """Grid resolution N for a file holding N*N elements (N a power of two)."""
    n = int(round(count ** 0.5))
    if n * n != count or n < 2 or n & (n - 1):
        raise ValueError(f"{count} elements is not a power-of-two square grid")
    return n
