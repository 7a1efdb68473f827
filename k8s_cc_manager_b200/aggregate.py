"""Cross-rank aggregation for multi-process runs (bench.py under torchrun).

The scrub path shards by GPU with NO data-path collective; the only
things ranks exchange are scalars after the timed region: the slowest rank's device
time (MAX — never a mean, never wall clock) and the bytes / kernel launches of all
ranks (SUM).  Backend-agnostic so the same code is exercised over gloo on CPU.
"""
from __future__ import annotations


def _reduce(dist, device, value: float, op: str) -> float:
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


def aggregate_job(dist, device, *, region_bytes: int, steps: int, elapsed_ms: float, launches: int) -> dict:
    """Whole-job numbers from per-rank measurements.

    value = (bytes zeroed + bytes read back by ALL ranks over `steps` steps) / (slowest
    rank's CUDA-event time).  Weak scaling: each rank's region is its own GPU's HBM.
    """
    ms = _reduce(dist, device, elapsed_ms, "MAX")
    total_bytes = _reduce(dist, device, float(region_bytes), "SUM")
    total_launches = int(_reduce(dist, device, float(launches), "SUM"))
    world = dist.get_world_size() if dist is not None else 1
    return {
        "world": world,
        "ms": ms,
        "total_region_bytes": int(total_bytes),
        "launches": total_launches,
        "value_gbs": 2.0 * total_bytes * steps / (ms * 1e-3) / 1e9,
    }
