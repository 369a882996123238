"""Sharding of independent notes across GPUs (one process per GPU, no data-path collective).

The reference proves batches of notes in parallel on CPU threads, one fresh RNG per note
(`/root/reference/src/utils/params_builder.rs:195-233`); here note i of a batch goes to rank
i mod world_size.  `torch.distributed` is used only for the barrier and for reducing timings
(max over ranks) and counters (sum over ranks)."""
from __future__ import annotations


def notes_for_rank(num_notes: int, world: int, rank: int) -> list[int]:
    """Global indices of the notes rank `rank` proves (round-robin deal)."""
    assert 0 <= rank < world
    return list(range(rank, num_notes, world))


def notes_per_rank(num_notes: int, world: int) -> list[int]:
    return [len(range(r, num_notes, world)) for r in range(world)]


def reduce_timing(ms: float, count: int, device=None, backend_tensor="cpu"):
    """(max over ranks of elapsed ms, sum over ranks of count).  No-op without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return ms, count
    dev = device if device is not None else backend_tensor
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor([count], dtype=torch.int64, device=dev)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), int(c.item())
