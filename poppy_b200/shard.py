"""Phase sharding for multi-GPU runs: frame phases are independent in direct mode, so the path is partitioned by
giving each rank a contiguous phase range; there is no data-path collective."""
from __future__ import annotations

import numpy as np


def phase_range(rank: int, world: int, n_total: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: lo = floor(rank * n / world)."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad rank/world/n_total")
    return (rank * n_total) // world, ((rank + 1) * n_total) // world


def phase_schedule(n_total: int) -> np.ndarray:
    """s_k = k / (n - 1), k = 0..n-1; a single phase renders s = 0.5."""
    if n_total == 1:
        return np.array([0.5], np.float32)
    return (np.arange(n_total, dtype=np.float64) / (n_total - 1)).astype(np.float32)


def combine_checksums(sums) -> int:
    """Order-sensitive fold of per-rank frame checksums (rank order = phase order)."""
    acc = 0xCBF29CE484222325
    for s in sums:
        acc = ((acc ^ (int(s) & 0xFFFFFFFFFFFFFFFF)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return acc
