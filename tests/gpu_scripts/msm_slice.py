"""Dev script: one bucket-range slice (part 3 of 8) of a 2^17-point MSM on one GPU, timed and under ncu."""
import os, sys, statistics
from ctypes import c_void_p
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cap_b200 import _lib, device, field
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
ctx = device.Context(0)
lib = ctx.lib
n = 1 << 17
srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
g = torch.Generator(device="cuda").manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[..., 3] &= (1 << 60) - 1
out = torch.zeros(16, dtype=torch.int64, device="cuda")
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fn = lambda: _lib.check(lib.capgpu_msm_g1_dev_part_xyzz(ctx.h, srs.h, 0, c_void_p(sc.data_ptr()), n, 0, 3 % parts, parts, c_void_p(out.data_ptr())), ctx.h)
stream = torch.cuda.ExternalStream(ctx.stream)
for _ in range(3):
    fn()
ctx.sync()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream); fn(); e1.record(stream)
    e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("slice of", parts, "ms", statistics.median(ts))
