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


# ---- one large MSM split by point range -----------------------------------------------------
def point_range(n_points: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) slice of the bases (and scalars) that rank `rank` owns."""
    lo = n_points * rank // world
    hi = n_points * (rank + 1) // world
    return lo, hi


def split_msm(ctx, srs_local, d_scalars_local, mont: bool = False):
    """sum_i s_i P_i with bases / scalars sharded by point range across the ranks of the default
    process group.  Each rank runs the MSM of its slice on its own GPU (`srs_local` holds only the
    slice's bases, `d_scalars_local` is a CUDA int64 tensor of shape (n_local, 4)), the 64-byte
    affine partial results are exchanged with ONE NCCL all-gather over NVLink, and every rank folds
    them with EC additions on its GPU.  Returns a CUDA int64 tensor (8,) = x || y of the result."""
    import torch
    import torch.distributed as dist
    from ctypes import c_void_p
    from . import _lib
    lib = ctx.lib
    n_local = int(d_scalars_local.shape[0])
    part = torch.zeros(8, dtype=torch.int64, device=d_scalars_local.device)
    _lib.check(lib.capgpu_msm_g1_dev(ctx.h, srs_local.h, 0, c_void_p(d_scalars_local.data_ptr()), n_local, 1, int(mont),
                                     c_void_p(part.data_ptr())), ctx.h)
    ctx.sync()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return part
    world = dist.get_world_size()
    gathered = torch.empty((world, 8), dtype=torch.int64, device=part.device)
    dist.all_gather_into_tensor(gathered, part)
    torch.cuda.current_stream().synchronize()
    out = torch.zeros(8, dtype=torch.int64, device=part.device)
    _lib.check(lib.capgpu_g1_sum_dev(ctx.h, c_void_p(gathered.data_ptr()), world, c_void_p(out.data_ptr())), ctx.h)
    ctx.sync()
    return out
