"""Host logic of the N > 1 path on CPU: world_size-2 gloo process group, round-robin note
sharding, max-over-ranks timing and sum-over-ranks counters (what bench.py does over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cap_b200 import shard


def test_round_robin_partition():
    for world in (1, 2, 4, 8):
        for num in (0, 1, 7, 64, 1024):
            parts = [shard.notes_for_rank(num, world, r) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(num))
            assert [len(p) for p in parts] == shard.notes_per_rank(num, world)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.notes_for_rank(11, world, rank)
    dist.barrier()
    ms, total = shard.reduce_timing(10.0 * (rank + 1), len(mine))
    # every rank sees the same reduced values
    gathered = [None] * world
    dist.all_gather_object(gathered, (ms, total, mine))
    if rank == 0:
        q.put(gathered)
    dist.destroy_process_group()


def test_two_rank_gloo_timing_reduction():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered = q.get()
    assert [g[0] for g in gathered] == [20.0, 20.0]  # max over ranks
    assert [g[1] for g in gathered] == [11, 11]      # all notes accounted for
    assert sorted(gathered[0][2] + gathered[1][2]) == list(range(11))


def test_no_group_is_identity():
    assert shard.reduce_timing(3.5, 7) == (3.5, 7)


def test_point_range_partition():
    for world in (1, 2, 3, 8):
        for n in (1, 17, 1 << 17):
            rs = [shard.point_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))


def test_bucket_parts():
    assert [shard.bucket_parts(w) for w in (1, 2, 3, 4, 5, 8)] == [1, 2, 2, 4, 4, 8]


def _split_worker(rank, world, port, q):
    """Split of one MSM: partial sums per rank (oracle arithmetic stands in for the GPU here),
    one all-gather of the 64-byte results, EC fold on every rank."""
    import random
    from oracle import bn254 as B
    from oracle import msm as omsm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, tau = 37, 123456789
    rng = random.Random(1)
    scalars = [rng.randrange(B.R) for _ in range(n)]
    bases = B.srs_powers(tau, n)
    lo, hi = shard.point_range(n, world, rank)
    part = omsm.msm_naive(bases[lo:hi], scalars[lo:hi])
    gathered = [None] * world
    dist.all_gather_object(gathered, part)
    total = None
    for p in gathered:
        total = B.g1_add(total, p)
    if rank == 0:
        q.put(total == omsm.kzg_commit_tau(scalars, tau))
    dist.destroy_process_group()


def test_split_msm_fold_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True
