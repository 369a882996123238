"""The other pairing curves of the reference's `CapConfig` (BLS12-381 / BLS12-377, /root/reference/src/config.rs:
86-114): oracle constants, the 12-limb field / G1 code of cap_b200/csrc/fpn.cuh compiled for the CPU against
exact integers, and (gpu) the same through the C ABI: `capgpu_curve_fq_op`, `capgpu_curve_msm_g1`."""
import ctypes
import random

import numpy as np
import pytest

from cap_b200 import curves as C
from oracle import curves as O

ALL = [C.BLS12_381, C.BLS12_377]


@pytest.mark.parametrize("cv", ALL, ids=lambda c: c.name)
def test_oracle_constants(cv):
    G = (cv.gx, cv.gy)
    assert O.on_curve(G, cv.q, cv.b)
    assert O.mul(cv.r, G, cv.q) is None and O.mul(cv.r - 1, G, cv.q) == O.neg(G, cv.q)
    assert O.add(G, G, cv.q) == O.mul(2, G, cv.q) and O.on_curve(O.mul(12345, G, cv.q), cv.q, cv.b)
    assert cv.q.bit_length() in (381, 377) and pow(3, cv.q - 1, cv.q) == 1 and pow(3, cv.r - 1, cv.r) == 1


def _fq_edge(cv):
    q = cv.q
    return [0, 1, 2, q - 1, q - 2, (q - 1) // 2, (q + 1) // 2, 1 << 380 if q.bit_length() == 381 else 1 << 376, (1 << 384) % q, (1 << 32) - 1,
            (1 << 352), int("ffffffff" * 11, 16) % q, int("80000000" * 11, 16) % q, (1 << 62) - 1, (1 << 62) + 1]


def _emu_op(emu, cv, op, a, b=None):
    ab = (ctypes.c_uint32 * 12).from_buffer_copy(a.to_bytes(48, "little"))
    bb = (ctypes.c_uint32 * 12).from_buffer_copy(b.to_bytes(48, "little")) if b is not None else None
    out = (ctypes.c_uint32 * 12)()
    emu.emu_fqn_op(cv.id, op, ab, bb, out)
    return int.from_bytes(bytes(out), "little")


@pytest.mark.parametrize("cv", ALL, ids=lambda c: c.name)
def test_field_code_on_cpu(emu, cv):
    """fpn.cuh (CIOS product, add / sub, batched binary-GCD inversion over 12 limbs) against Python integers."""
    rng = random.Random(11)
    q = cv.q
    ri = pow(C.R384, -1, q)
    edge = _fq_edge(cv)
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(q), rng.randrange(q)) for _ in range(1500)]
    for a, b in pairs:
        assert _emu_op(emu, cv, 0, a, b) == a * b * ri % q
        assert _emu_op(emu, cv, 3, a, b) == (a + b) % q
        assert _emu_op(emu, cv, 4, a, b) == (a - b) % q
    r2 = C.R384 * C.R384 % q
    vals = [v for v in edge if v] + [rng.randrange(1, 1 << k) % q or 1 for k in range(1, 381, 3)] + [rng.randrange(1, q) for _ in range(400)]
    for x in vals:
        want = pow(x, -1, q) * r2 % q  # x is read as a Montgomery residue y R: the result is R / y
        assert _emu_op(emu, cv, 2, x) == want, hex(x)
        assert _emu_op(emu, cv, 1, x) == x * x * ri % q
        assert _emu_op(emu, cv, 6, x) == (-x) % q
    for x in vals[:6]:
        assert _emu_op(emu, cv, 5, x) == pow(x, -1, q) * r2 % q
    assert _emu_op(emu, cv, 2, 0) == 0 and _emu_op(emu, cv, 6, 0) == 0


@pytest.mark.parametrize("cv", ALL, ids=lambda c: c.name)
def test_group_code_on_cpu(emu, cv):
    """XYZZ mixed / full additions, doublings and the affine conversion of fpn.cuh, incl. the exact special cases."""
    rng = random.Random(5)
    q = cv.q
    G = (cv.gx, cv.gy)
    pts = [O.mul(rng.randrange(1, cv.r), G, q) for _ in range(6)]

    def chain(pa, negs, pb, dbl):
        a = C.g1_to_mont_array(cv, pa) if pa else np.zeros((1, 12), dtype=np.uint64)
        b = C.g1_to_mont_array(cv, pb) if pb else np.zeros((1, 12), dtype=np.uint64)
        ng = (ctypes.c_int * max(1, len(pa)))(*negs)
        out = np.zeros(12, dtype=np.uint64)
        emu.emu_g1n_chain(cv.id, a.ctypes.data_as(ctypes.c_void_p), ng, len(pa), b.ctypes.data_as(ctypes.c_void_p), len(pb), dbl,
                          out.ctypes.data_as(ctypes.c_void_p))
        return C.g1_from_mont_array(cv, out)[0]

    def want(pa, negs, pb, dbl):
        acc = None
        for p, n in zip(pa, negs):
            acc = O.add(acc, O.neg(p, q) if n else p, q)
        for p in pb:
            acc = O.add(acc, p, q)
        return O.mul(1 << dbl, acc, q)

    cases = [
        (pts[:4], [0, 1, 0, 0], pts[4:], 3),
        ([pts[0], pts[0]], [0, 0], [], 0),                 # mixed doubling
        ([pts[0], pts[0]], [0, 1], [pts[1]], 1),           # cancellation, then + P
        ([pts[0], pts[1]], [0, 0], [pts[0], pts[1]], 0),   # full-addition doubling
        ([pts[0], pts[1]], [0, 0], [O.neg(pts[1], q), O.neg(pts[0], q)], 2),  # full-addition cancellation -> infinity
        ([], [], [pts[2]], 5),
        ([pts[3]], [1], [], 0),
    ]
    for pa, negs, pb, dbl in cases:
        assert chain(pa, negs, pb, dbl) == want(pa, negs, pb, dbl)


@pytest.mark.gpu
@pytest.mark.parametrize("cv", ALL, ids=lambda c: c.name)
def test_gpu_field_ops(ctx, cv):
    rng = random.Random(21)
    q = cv.q
    ri = pow(C.R384, -1, q)
    r2 = C.R384 * C.R384 % q
    edge = _fq_edge(cv)
    pairs = [(a, b) for a in edge for b in edge] + [(rng.randrange(q), rng.randrange(q)) for _ in range(2000)]
    a = np.frombuffer(b"".join(x.to_bytes(48, "little") for x, _ in pairs), dtype="<u8").reshape(-1, 6)
    b = np.frombuffer(b"".join(y.to_bytes(48, "little") for _, y in pairs), dtype="<u8").reshape(-1, 6)

    def ints(arr):
        return [int.from_bytes(arr[i].tobytes(), "little") for i in range(arr.shape[0])]

    assert ints(C.fq_op(ctx, cv, 0, a, b)) == [x * y * ri % q for x, y in pairs]
    assert ints(C.fq_op(ctx, cv, 3, a, b)) == [(x + y) % q for x, y in pairs]
    assert ints(C.fq_op(ctx, cv, 4, a, b)) == [(x - y) % q for x, y in pairs]
    assert ints(C.fq_op(ctx, cv, 1, a)) == [x * x * ri % q for x, _ in pairs]
    inv = [pow(x, -1, q) * r2 % q if x else 0 for x, _ in pairs]
    assert ints(C.fq_op(ctx, cv, 2, a)) == inv
    assert ints(C.fq_op(ctx, cv, 5, a[:64])) == inv[:64]


@pytest.mark.gpu
@pytest.mark.parametrize("cv", ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("n", [1, 37, 600])
def test_gpu_msm_matches_oracle(ctx, cv, n):
    """capgpu_curve_msm_g1 against the definition, with edge scalars, repeated bases, a negated pair and infinity."""
    rng = random.Random(100 + n)
    q = cv.q
    G = (cv.gx, cv.gy)
    base = [O.mul(rng.randrange(1, cv.r), G, q) for _ in range(min(n, 24))]
    pts = [base[i % len(base)] for i in range(n)]
    sc = [rng.randrange(cv.r) for _ in range(n)]
    for i, s in enumerate([0, 1, cv.r - 1, cv.r - 2, (1 << 128) - 1, 1 << 200, (cv.r - 1) // 2, 2][: n]):
        sc[i] = s
    if n >= 30:
        pts[25] = O.neg(pts[1], q)   # P and -P in the same MSM
        sc[25] = sc[1] = 12345
        pts[26] = None               # infinity among the bases
        pts[27] = pts[3]
        sc[27] = sc[3]               # the same term twice: a bucket doubling
    got = C.g1_from_mont_array(cv, C.msm_g1(ctx, cv, C.g1_to_mont_array(cv, pts), sc))[0]
    assert got == O.msm_naive([p for p in pts], sc, q)


@pytest.mark.gpu
def test_gpu_curve_argument_errors(ctx):
    out = np.zeros(12, dtype=np.uint64)
    lib = ctx.lib
    assert lib.capgpu_curve_msm_g1(ctx.h, 7, None, None, 0, out.ctypes.data_as(ctypes.c_void_p)) == -2
    assert lib.capgpu_curve_msm_g1(ctx.h, 1, None, None, 0, out.ctypes.data_as(ctypes.c_void_p)) == 0 and not out.any()
    assert lib.capgpu_curve_fq_op(ctx.h, 1, 0, out.ctypes.data_as(ctypes.c_void_p), None, out.ctypes.data_as(ctypes.c_void_p), 2) == -2
