"""Batch sharding of the hot path across ranks: samples are
independent, so rank r takes a contiguous slice and the only exchange is a sum
of loss numerators / counts (and the 428 gradient-penalty column sums)."""
import numpy as np


def shard_range(batch, world, rank):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most 1."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_partials(kp_abs_sum, kp_count, mesh_sum=0.0, gp_col_sums=None):
    """The fp32 vector a rank contributes to the all-reduce: [kp numerator,
    kp count, mesh sum, 428 GP column sums].  Counts are exact in fp32 below
    2^24 (global batch * K * 2 << 2^24 for every BASELINE config)."""
    v = np.zeros(3 + 428, dtype=np.float32)
    v[0], v[1], v[2] = kp_abs_sum, kp_count, mesh_sum
    if gp_col_sums is not None:
        v[3:] = np.asarray(gp_col_sums, dtype=np.float32)
    return v


def finish_losses(reduced, m_total=None):
    """Global losses from the all-reduced vector: numerators and counts are
    reduced separately and divided once (averaging per-shard losses is NOT the
    reference's loss: shards have different num_present)."""
    kp = float(reduced[0] / reduced[1]) if reduced[1] > 0 else 0.0
    out = {"kp_loss": kp, "kp_count": int(reduced[1]), "mesh_loss": float(reduced[2])}
    if m_total:
        cols = np.asarray(reduced[3:], dtype=np.float64) / m_total
        segs = [(0, 169), (169, 211), (211, 221), (221, 428)]
        out["gradient_penalty"] = float(sum((1.0 - np.sqrt(np.sum(cols[a:b] ** 2))) ** 2 for a, b in segs))
    return out
