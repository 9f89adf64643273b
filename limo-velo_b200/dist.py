"""Multi-GPU plumbing of the benchmark: independent sequences, one per rank, no data-path collective.

The hot path does not shard inside a sequence (sweep n+1 needs the state and the map of sweep n:
src/main.cpp:84-105), so N GPUs run N independent sequences (BASELINE.json configs[4]).  The only
exchange is the final gather of throughput counters, done with torch.distributed (NCCL on GPUs, gloo
in the CPU tests).
"""
import torch
import torch.distributed as dist


def sequence_for_rank(rank, world_size, n_sequences=None):
    """Round-robin assignment of independent sequences to ranks (one each by default)."""
    if n_sequences is None:
        n_sequences = world_size
    return [s for s in range(n_sequences) if s % world_size == rank]


def reduce_counters(step_ms, points, matched, e2e_s, e2e_points, launches, device="cpu"):
    """Whole-job numbers from per-rank counters: times are the MAX over ranks, work is the SUM."""
    t = torch.tensor([step_ms, float(points), float(matched), e2e_s, float(e2e_points), float(launches)],
                     dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = t, t
    return {"step_ms": float(mx[0]), "e2e_s": float(mx[3]), "points": float(sm[1]), "matched": float(sm[2]),
            "e2e_points": float(sm[4]), "launches": float(sm[5])}


def throughput(points, ms):
    return points / (ms * 1e-3) if ms > 0 else 0.0
