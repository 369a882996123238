"""Dev script (gpurun): device-resident lone / batched MSM timings with the per-stage profile, every result
checked through p(tau) G.  Sizes from argv (log_n:batch pairs), default 12..17 lone + 15x5 + 15x40."""
import json
import os
import statistics
import sys
from ctypes import byref, c_double, c_uint64, c_void_p

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from cap_b200 import _lib, device, field  # noqa: E402
from oracle import msm as omsm  # noqa: E402

TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
wb = int(os.environ.get("CAPGPU_WINDOW_BITS", "0"))
ctx = device.Context(0)
lib = ctx.lib
stream = torch.cuda.ExternalStream(ctx.stream)
g = torch.Generator(device="cuda").manual_seed(1)
res = {"env": {k: v for k, v in os.environ.items() if k.startswith("CAPGPU_")}}


def timeit(fn, reps=20):
    for _ in range(3):
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
    return statistics.median(ts), min(ts)


def profile(fn):
    lib.capgpu_profile_enable(ctx.h, 1)
    fn()
    ctx.sync()
    out = {}
    for pid, name in enumerate(["accumulate", "ntt", "quotient", "sort", "reduce", "gp"]):
        ms, cnt, units = c_double(), c_uint64(), c_double()
        lib.capgpu_profile_read(ctx.h, pid, byref(ms), byref(cnt), byref(units))
        if cnt.value:
            out[name] = round(ms.value, 4)
    lib.capgpu_profile_enable(ctx.h, 0)
    return out


cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]] or [(12, 1), (13, 1), (14, 1), (15, 1), (16, 1), (17, 1), (15, 5), (15, 40)]
for log_n, batch in cases:
    n = (1 << log_n) + (3 if batch > 1 else 0)
    srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n, window_bits=wb)
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (batch, n, 4), dtype=torch.int64, device="cuda", generator=g)
    sc[..., 3] &= (1 << 60) - 1  # uniform 252-bit scalars
    out = torch.zeros((batch, 8), dtype=torch.int64, device="cuda")
    fn = lambda: _lib.check(lib.capgpu_msm_g1_dev(ctx.h, srs.h, 0, c_void_p(sc.data_ptr()), n, batch, 0, c_void_p(out.data_ptr())), ctx.h)
    ms, best = timeit(fn)
    prof = profile(fn)
    ok = True
    for v in sorted({0, batch - 1}):
        host = sc[v].cpu().numpy().view("uint64")
        got = field.g1_from_mont_array(out[v].cpu().numpy().view("uint64"))[0]
        ok = ok and got == omsm.kzg_commit_tau(field.fr_from_raw_array(host), TAU)
    res[f"msm_2^{log_n}_x{batch}"] = {"ms": round(ms, 4), "min": round(best, 4), "ok": ok, **prof}
    print(f"msm_2^{log_n}_x{batch}", res[f"msm_2^{log_n}_x{batch}"], flush=True)
    srs.close()
print(json.dumps(res))
