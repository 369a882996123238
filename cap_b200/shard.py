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


# ---- one large MSM split across the GPUs of a box -----------------------------------------------
def point_range(n_points: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) slice of the bases (and scalars) that rank `rank` owns in a point-range split."""
    lo = n_points * rank // world
    hi = n_points * (rank + 1) // world
    return lo, hi


def bucket_parts(world: int) -> int:
    """Bucket-range slices for `world` ranks: the largest power of two <= world (the bucket count
    2^(c-1) is a power of two; ranks beyond it hold an empty slice and contribute infinity)."""
    p = 1
    while p * 2 <= world:
        p *= 2
    return p


class SplitMsm:
    """sum_i s_i P_i for ONE scalar vector, split across the ranks of the default process group by
    BUCKET RANGE (BASELINE north_star: "a single large MSM can be split ... with partial sums reduced
    over NVLink").  Every rank holds the whole commit key `srs` and the whole scalar vector; rank r
    sorts, accumulates and reduces only the buckets of slice r (capgpu_msm_g1_dev_part), so all three
    phases shrink with the number of GPUs.  The slice results (128-byte XYZZ sums) are exchanged with one NCCL
    all-gather and folded with EC additions and ONE conversion to affine (capgpu_g1_sum_xyzz_dev).  Everything -- the MSM kernels, the
    all-gather and the fold -- is enqueued on the context's stream (it is made torch's current stream
    for the collective), with no host synchronisation in between; the caller synchronises once."""

    def __init__(self, ctx, srs, exchange: str = "auto"):
        """exchange: "peer" = slice results stored straight into every GPU's symmetric-memory buffer by the last
        reduction kernel, flags released with system scope, the fold waits on the flags (no collective call);
        "nccl" = one all-gather on the context stream; "auto" = peer when symmetric memory can be set up."""
        import torch
        import torch.distributed as dist
        self.ctx, self.srs = ctx, srs
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.parts = bucket_parts(self.world)
        self.stream = torch.cuda.ExternalStream(ctx.stream)
        dev = torch.device("cuda", ctx.device)
        # keys of >= 2^12 points have >= 2^14 buckets: slices then hand over their XYZZ sum (128 B) and only the
        # fold converts to affine (one inversion on the critical path instead of two)
        self.xyzz = srs.size >= (1 << 12) and self.world > 1
        words = 16 if self.xyzz else 8
        self.part = torch.zeros(words, dtype=torch.int64, device=dev)
        self.gathered = torch.zeros((self.world, words), dtype=torch.int64, device=dev)
        self.out = torch.zeros(8, dtype=torch.int64, device=dev)
        self.peer = None
        self.epoch = 0
        if self.xyzz and exchange in ("auto", "peer") and self.parts == self.world and self.world <= 16:
            try:
                self._setup_peer(torch, dist, dev)
            except Exception:  # symmetric memory unavailable on this box: the collective path stays
                if exchange == "peer":
                    raise
                self.peer = None

    def _setup_peer(self, torch, dist, dev):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        # two alternating buffers of [parts slots of 16 x int64 | parts flags at 8-byte stride]
        self.slot_words = 16 * self.parts
        self.buf_words = self.slot_words + 16 * ((self.parts + 15) // 16)  # flags padded: every half stays 128-byte aligned
        buf = symm_mem.empty(2 * self.buf_words, dtype=torch.int64, device=dev)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        torch.cuda.synchronize()
        dist.barrier()
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.peer = {"buf": buf, "hdl": hdl, "ptrs": ptrs}
        self.peer_args = []
        for half in range(2):
            base = half * self.buf_words * 8
            slots = (ctypes.c_void_p * self.world)(*[p + base + self.rank * 128 for p in ptrs])
            flags = (ctypes.c_void_p * self.world)(*[p + base + self.slot_words * 8 + self.rank * 8 for p in ptrs])
            self.peer_args.append((slots, flags, buf.data_ptr() + base, buf.data_ptr() + base + self.slot_words * 8))

    def __call__(self, d_scalars, mont: bool = False):
        """d_scalars: CUDA int64 tensor (n, 4) on the context's GPU, identical on every rank.  Returns
        a CUDA int64 tensor (8,) = x || y of the result (valid after ctx.sync()).  With the peer exchange the
        ranks must not run more than one call ahead of each other (two alternating buffers)."""
        import torch
        import torch.distributed as dist
        from ctypes import c_void_p
        from . import _lib
        lib, ctx = self.ctx.lib, self.ctx
        n = int(d_scalars.shape[0])
        if self.peer is not None:
            self.epoch += 1
            slots, flags, my_slots, my_flags = self.peer_args[self.epoch & 1]
            _lib.check(lib.capgpu_msm_g1_dev_part_peer(ctx.h, self.srs.h, 0, c_void_p(d_scalars.data_ptr()), n, int(mont), self.rank, self.parts,
                                                       slots, flags, self.world, self.epoch), ctx.h)
            _lib.check(lib.capgpu_g1_sum_xyzz_wait_dev(ctx.h, c_void_p(my_slots), c_void_p(my_flags), 8, self.world, self.epoch,
                                                       c_void_p(self.out.data_ptr())), ctx.h)
            return self.out
        slice_fn = lib.capgpu_msm_g1_dev_part_xyzz if self.xyzz else lib.capgpu_msm_g1_dev_part
        if self.rank < self.parts:
            _lib.check(slice_fn(ctx.h, self.srs.h, 0, c_void_p(d_scalars.data_ptr()), n, int(mont), self.rank, self.parts,
                                c_void_p(self.part.data_ptr())), ctx.h)
        if self.world == 1:
            return self.part
        with torch.cuda.stream(self.stream):
            if self.rank >= self.parts:
                self.part.zero_()
            dist.all_gather_into_tensor(self.gathered, self.part)
        fold = lib.capgpu_g1_sum_xyzz_dev if self.xyzz else lib.capgpu_g1_sum_dev
        _lib.check(fold(ctx.h, c_void_p(self.gathered.data_ptr()), self.world, c_void_p(self.out.data_ptr())), ctx.h)
        return self.out
