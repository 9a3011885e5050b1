"""Batch-mode partitioning of independent proofs across ranks/GPUs: contiguous by proof index, no
data-path collective; the per-proof RNG seed depends only on the global proof index, so results are
placement-independent."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of proof indices owned by `rank` out of `world` (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def seeds_for_range(base, lo, hi):
    """seed of proof i = LE64(base + i) || 24 zero bytes  (uint8[hi-lo, 32])."""
    idx = (np.arange(lo, hi, dtype=np.uint64) + np.uint64(base))
    out = np.zeros((hi - lo, 32), np.uint8)
    out[:, :8] = idx.view(np.uint8).reshape(-1, 8)
    return out
