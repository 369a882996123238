"""Workload for compute-sanitizer: an n = 2^9 lockstep group of 3 proofs (every
prover kernel, the data-dependent atomicAdd scatters, the flat accumulation with its chunk-boundary
writes in latency mode) and a 2^13-point MSM with skewed scalars (heavy buckets, empty buckets, a
bucket-range slice), each checked against the single-proof / whole-MSM result.
    compute-sanitizer --tool memcheck  python tests/gpu_scripts/sanitize.py
    compute-sanitizer --tool racecheck python tests/gpu_scripts/sanitize.py"""
import os
import sys
from ctypes import c_void_p

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from cap_b200 import _lib, device, field, plonk, synth  # noqa: E402

TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % field.R
ctx = device.Context(0)
ctx.set_group(3)
lib = ctx.lib

# ---- prover, n = 2^9, group of 3 + the same proofs one at a time in latency mode ---------------------
circ = synth.make_circuit(9, num_inputs=5, seed=3, zero_inputs=0.3, bool_inputs=0.3)
srs = plonk.PlonkKzgSnark.universal_setup(ctx, circ.n + 2, TAU)
pk = plonk.PlonkKzgSnark.preprocess(ctx, srs, circ)
circs = [circ, circ.with_witness(1), circ.with_witness(2)]
wires = [plonk.wire_values(c) for c in circs]
pubs = [field.fr_to_mont_array(plonk.public_input(c)) for c in circs]
rng = np.random.default_rng(1)
bls = rng.integers(0, 1 << 62, size=(3, 17, 4), dtype=np.uint64)
bls[..., 3] &= (1 << 60) - 1
proofs, status = plonk.prove_batch_raw([ctx], pk, [w.ctypes.data for w in wires], pubs, list(bls), [b"s0", b"", b"s2"])
assert status == [0, 0, 0]
ctx.set_latency_mode(True)
for i, msg in enumerate([b"s0", b"", b"s2"]):
    single = plonk.PlonkKzgSnark.prove_raw(ctx, pk, wires[i], pubs[i], bls[i], msg)
    assert bytes(single) == bytes(proofs[i]), i
ctx.set_latency_mode(False)
print("prover: group of 3 == single proofs (throughput and latency schedules)", flush=True)

# ---- 2^13-point MSM, skewed scalars ------------------------------------------------------------------
n = 1 << 13
srs13 = device.Srs(ctx, tau_mont=field.fr_to_mont_array([TAU])[0], size=n)
sc = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
sc[:, 3] &= (1 << 60) - 1
sc[: n // 2, 1:] = 0
sc[: n // 2, 0] = rng.integers(0, 3, size=n // 2, dtype=np.uint64)  # half the scalars in {0, 1, 2}: one heavy bucket, many zeros
sc[n // 2: n // 2 + 64] = sc[n - 1]                                 # a repeated scalar
whole = srs13.msm(sc, mont=False)
batch = srs13.msm(np.stack([sc, sc[::-1].copy()]), mont=False)
assert np.array_equal(batch[0], whole)
d = torch.from_numpy(sc.view(np.int64)).cuda()
outs = torch.zeros((4, 8), dtype=torch.int64, device="cuda")
for p in range(4):
    _lib.check(lib.capgpu_msm_g1_dev_part(ctx.h, srs13.h, 0, c_void_p(d.data_ptr()), n, 0, p, 4, c_void_p(outs[p].data_ptr())), ctx.h)
total = torch.zeros(8, dtype=torch.int64, device="cuda")
_lib.check(lib.capgpu_g1_sum_dev(ctx.h, c_void_p(outs.data_ptr()), 4, c_void_p(total.data_ptr())), ctx.h)
ctx.sync()
assert np.array_equal(total.cpu().numpy().view(np.uint64), whole)
print("msm: skewed 2^13 whole == batched == 4 bucket-range slices", flush=True)

# ---- NTT round trip 2^13 (two-pass register-radix kernels) -------------------------------------------
a = rng.integers(0, 1 << 62, size=(2, n, 4), dtype=np.uint64)
a[..., 3] &= (1 << 60) - 1
f = ctx.ntt(a, 13, inverse=False, coset=True)
b = ctx.ntt(f, 13, inverse=True, coset=True)
assert np.array_equal(a, b)
print("ntt: coset round trip 2^13", flush=True)
print("SANITIZE_WORKLOAD_OK", flush=True)
