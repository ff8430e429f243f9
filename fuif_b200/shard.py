"""Multi-GPU plumbing of the decode path: files are independent units (own header, own MANIAC trees, own range
coders -- SURVEY.md 8e), so a batch shards one-image-per-GPU with NO data-path collective.  torch.distributed is
used only to agree on the timing (max over ranks) and to count what was decoded."""
from __future__ import annotations


def units_for_rank(n_units_per_rank: int, rank: int) -> list:
    """Weak scaling: every rank owns n_units_per_rank consecutive global unit ids (= image seeds)."""
    return [rank * n_units_per_rank + i for i in range(n_units_per_rank)]


def replica_units(n_units_per_rank: int) -> list:
    """Weak scaling with replicas: every rank decodes its own copy of the same n_units_per_rank images, so the per-GPU work is
    identical and the max over ranks measures the machine, not which image happened to land where."""
    return list(range(n_units_per_rank))


def split_batch(n_units: int, rank: int, world: int) -> list:
    """Strong scaling of a fixed batch: round-robin, the layout BASELINE config 4 (64 images over 8 GPUs) uses."""
    return list(range(rank, n_units, world))


def reduce_step_time(ms_local: float, units_local: int, device=None):
    """Returns (max step time over ranks in ms, total units over ranks). Works on gloo (CPU) and nccl."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return ms_local, units_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), int(u.item())
