"""torchrun script (N >= 2 GPUs): ONE MSM split by bucket range across the ranks (cap_b200.shard.
SplitMsm: slice MSM kernels, NCCL all-gather of the 64-byte slice results and the EC fold all
enqueued on the context stream), checked against p(tau) * G from the oracle and timed beside the
single-GPU MSM.  Used by tests/test_gpu_primitives.py::test_split_msm_across_gpus and as evidence
(gpurun_out/split_msm_<N>gpu.json):
    torchrun --nproc-per-node N tests/gpu_scripts/split_msm.py [--points P] [--reps R]"""
import argparse
import json
import os
import statistics
import sys
from ctypes import c_void_p

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cap_b200 import _lib, device, field, shard  # noqa: E402
from oracle import msm as omsm  # noqa: E402  (checker)

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1 << 17)
ap.add_argument("--reps", type=int, default=12)
args = ap.parse_args()
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = device.Context(local)
N = args.points
srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=N)  # the whole commit key on every GPU
sc = np.random.default_rng(5).integers(0, 1 << 62, size=(N, 4), dtype=np.uint64)
sc[:, 3] &= (1 << 60) - 1
d_all = torch.from_numpy(sc.view(np.int64)).cuda()
split = shard.SplitMsm(ctx, srs)                        # peer-memory exchange when symmetric memory is available
split_nccl = shard.SplitMsm(ctx, srs, exchange="nccl")  # one NCCL all-gather on the context stream

want = omsm.kzg_commit_tau(field.fr_from_raw_array(sc), TAU)
ok = True
for s_ in (split, split_nccl, split):
    res = s_(d_all)
    ctx.sync()
    dist.barrier()
    ok = ok and field.g1_from_mont_array(res.cpu().numpy().view(np.uint64))[0] == want
oks = [None] * world
dist.all_gather_object(oks, bool(ok))


def timed(fn, reps):
    """CUDA events on the context stream, max over ranks, median of reps (2 warm-ups dropped)."""
    stream = torch.cuda.ExternalStream(ctx.stream)
    ts = []
    for _ in range(reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return statistics.median(ts[2:])


split_ms = timed(lambda: split(d_all), args.reps)
nccl_ms = timed(lambda: split_nccl(d_all), args.reps)
out1 = torch.zeros(8, dtype=torch.int64, device="cuda")
single_ms = timed(lambda: _lib.check(ctx.lib.capgpu_msm_g1_dev(ctx.h, srs.h, 0, c_void_p(d_all.data_ptr()), N, 1, 0, c_void_p(out1.data_ptr())), ctx.h),
                  args.reps)
if rank == 0:
    line = {"n_gpus": world, "points": N, "split": "bucket range", "bucket_parts": split.parts, "exchange": "peer memory (symmetric memory, stores + flags fused into the kernels)" if split.peer is not None else "nccl all-gather",
            "split_msm_ms": split_ms, "split_msm_ms_nccl_all_gather": nccl_ms,
            "single_gpu_msm_ms": single_ms, "speedup": single_ms / split_ms, "correct": all(oks)}
    print(json.dumps(line), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(line, open(f"gpurun_out/split_msm_{world}gpu.json", "w"))
dist.barrier()
dist.destroy_process_group()
