"""Dev / evidence script (torchrun, N >= 2 GPUs): one 2^17-point MSM split by point range across
the ranks, partial results all-gathered over NCCL/NVLink and folded with EC additions; checked
against p(tau) * G from the oracle and timed beside the single-GPU MSM."""
import json
import os
import statistics
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from cap_b200 import device, field, shard  # noqa: E402
from oracle import msm as omsm  # noqa: E402

TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = device.Context(local)
N = 1 << 17
full = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=N)
lo, hi = shard.point_range(N, world, rank)
local_srs = device.Srs(ctx, points_xy=full.export()[lo:hi])
sc = np.random.default_rng(5).integers(0, 1 << 62, size=(N, 4), dtype=np.uint64)
sc[:, 3] &= (1 << 60) - 1
d_all = torch.from_numpy(sc.view(np.int64)).cuda()
d_loc = d_all[lo:hi].contiguous()

res = shard.split_msm(ctx, local_srs, d_loc)
got = field.g1_from_mont_array(res.cpu().numpy().view(np.uint64))[0]
ok = True
if rank == 0:
    ok = got == omsm.kzg_commit_tau(field.fr_from_raw_array(sc), TAU)


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return statistics.median(ts[2:])


split_ms = timed(lambda: shard.split_msm(ctx, local_srs, d_loc))
out1 = torch.zeros(8, dtype=torch.int64, device="cuda")
from ctypes import c_void_p  # noqa: E402
from cap_b200 import _lib  # noqa: E402


def single():
    _lib.check(ctx.lib.capgpu_msm_g1_dev(ctx.h, full.h, 0, c_void_p(d_all.data_ptr()), N, 1, 0, c_void_p(out1.data_ptr())), ctx.h)
    ctx.sync()


single_ms = timed(single)
if rank == 0:
    line = {"n_gpus": world, "points": N, "split_msm_ms": split_ms, "single_gpu_msm_ms": single_ms, "correct": bool(ok)}
    print(json.dumps(line), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(line, open(f"gpurun_out/split_msm_{world}gpu.json", "w"))
dist.barrier()
dist.destroy_process_group()
