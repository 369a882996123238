"""Standalone kernel sweep (BASELINE config 4): G1 MSM 2^12..2^17 points and scalar-field NTT
2^12..2^18, device-resident and timed with CUDA events on the ctx stream, beside the C restatement
of arkworks' CPU algorithms on the host cores.  Writes gpurun_out/sweep.json."""
import json
import os
import statistics
import sys
import time
from ctypes import c_void_p

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from cap_b200 import _lib, device, field  # noqa: E402
from oracle import cpu  # noqa: E402  (checker / CPU baseline)

TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
ctx = device.Context(0)
lib = ctx.lib
stream = torch.cuda.ExternalStream(ctx.stream)
g = torch.Generator(device="cuda").manual_seed(3)
threads = os.cpu_count() or 1
calib = ctx.calibrate()
out = {"calibration": calib, "cpu_threads": threads, "msm": [], "ntt": []}


def timeit(fn, reps=10):
    fn()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn()
            e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def cpu_time(fn, reps=2):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts) * 1e3


for log_n in range(12, 18):
    n = 1 << log_n
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    sc[:, 3] &= (1 << 60) - 1
    res = torch.zeros(8, dtype=torch.int64, device="cuda")
    ms = timeit(lambda: _lib.check(lib.capgpu_msm_g1_dev(ctx.h, srs.h, 0, c_void_p(sc.data_ptr()), n, 1, 0, c_void_p(res.data_ptr())), ctx.h))
    host_sc = sc.cpu().numpy().view(np.uint64)
    srs_xy = srs.export()
    cpu_res = cpu.msm(srs_xy, host_sc, mont=False, nthreads=threads)
    same = bool(np.array_equal(cpu_res, res.cpu().numpy().view(np.uint64)))
    cms = cpu_time(lambda: cpu.msm(srs_xy, host_sc, mont=False, nthreads=threads))
    out["msm"].append({"log_n": log_n, "gpu_ms": ms, "cpu_ms": cms, "speedup": cms / ms, "bit_exact_vs_cpu": same})
    print("msm", out["msm"][-1], flush=True)
    srs.close()

for log_n in range(12, 19):
    n = 1 << log_n
    a = torch.randint(0, 1 << 60, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    b = torch.empty_like(a)
    ms = timeit(lambda: _lib.check(lib.capgpu_ntt_dev(ctx.h, c_void_p(a.data_ptr()), n, c_void_p(b.data_ptr()), log_n, 1, 0, 0), ctx.h))
    host = a.cpu().numpy().view(np.uint64)
    cpu_res = cpu.ntt(host, log_n, False, False, nthreads=threads)
    same = bool(np.array_equal(cpu_res, b.cpu().numpy().view(np.uint64)))
    cms = cpu_time(lambda: cpu.ntt(host, log_n, False, False, nthreads=threads))
    gbs = 2 * 32 * n * (2 if log_n > 10 else 1) / (ms * 1e-3) * 1e-9
    out["ntt"].append({"log_n": log_n, "gpu_ms": ms, "cpu_ms": cms, "speedup": cms / ms, "bit_exact_vs_cpu": same,
                       "algorithmic_gbs": gbs, "gbutterflies_per_s": (n / 2) * log_n / (ms * 1e-3) * 1e-9})
    print("ntt", out["ntt"][-1], flush=True)

# batch-verification shape (benches/batch_verification.rs): the aggregated commitment sum of 1024
# proofs of one note type is ~18 + 13 * 1024 + 1 ad-hoc bases; host buffers in, affine point out
from cap_b200.device import msm_adhoc  # noqa: E402
n = 18 + 13 * 1024 + 1
srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
pts = srs.export()
srs.close()
host_sc = np.random.default_rng(5).integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
host_sc[:, 3] &= (1 << 60) - 1
got = msm_adhoc(ctx, pts, host_sc, mont=False)
gms = cpu_time(lambda: msm_adhoc(ctx, pts, host_sc, mont=False), reps=5)
cpu_res = cpu.msm(pts, host_sc, mont=False, nthreads=threads)
cms = cpu_time(lambda: cpu.msm(pts, host_sc, mont=False, nthreads=threads))
out["adhoc_msm_batch_verify_1024"] = {"points": n, "gpu_ms_host_to_host": gms, "cpu_ms": cms, "speedup": cms / gms,
                                      "bit_exact_vs_cpu": bool(np.array_equal(cpu_res, got))}
print("adhoc", out["adhoc_msm_batch_verify_1024"], flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
