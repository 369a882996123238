"""Profiling target: one batched MSM (default 5 x (2^15+3) uniform scalars; PROF_LOGN / PROF_BATCH override) inside a profiler window."""
import os, sys
from ctypes import c_void_p
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from cap_b200 import _lib, device, field
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
ctx = device.Context(0)
log_n, batch = int(os.environ.get("PROF_LOGN", "15")), int(os.environ.get("PROF_BATCH", "5"))
n = (1 << log_n) + (3 if log_n == 15 else 0)
srs = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
g = torch.Generator(device="cuda").manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (batch, n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[..., 3] &= (1 << 60) - 1
out = torch.zeros((batch, 8), dtype=torch.int64, device="cuda")
fn = lambda: _lib.check(ctx.lib.capgpu_msm_g1_dev(ctx.h, srs.h, 0, c_void_p(sc.data_ptr()), n, batch, 0, c_void_p(out.data_ptr())), ctx.h)
fn(); ctx.sync(); torch.cuda.synchronize()
torch.cuda.profiler.start()
fn(); ctx.sync()
torch.cuda.profiler.stop()
