"""The device field / curve algorithms (cap_b200/csrc/fp.cuh, ec.cuh) and the library's host-side
helpers (hostfp.h, transcript.h), compiled for the CPU, against the big-int oracle.  The device
code runs with PTX add.cc / madc.hi.cc semantics emulated instruction by instruction."""
import ctypes
import os
import random

import numpy as np

from cap_b200.field import g1_from_mont_array, g1_to_mont_array
from oracle import bn254 as B
from oracle import keccak, transcript

RI = pow(1 << 256, -1, B.R)
QI = pow(1 << 256, -1, B.Q)


def _call(lib, fn, *xs):
    bufs = [(ctypes.c_uint32 * 8).from_buffer_copy(x.to_bytes(32, "little")) for x in xs]
    out = (ctypes.c_uint32 * 8)()
    getattr(lib, fn)(*bufs, out)
    return int.from_bytes(bytes(out), "little")


def test_device_field_ops(emu):
    rng = random.Random(1)
    edge = [0, 1, 2, B.R - 1, B.R - 2, 1 << 253, (1 << 256) % B.R, (1 << 32) - 1, (1 << 224)]
    pairs = [(a, b) for a in edge for b in edge] + [(None, None)] * 4000
    for a0, b0 in pairs:
        for mod, pre, mi in ((B.R, "fr", RI), (B.Q, "fq", QI)):
            a = rng.randrange(mod) if a0 is None else a0 % mod
            b = rng.randrange(mod) if b0 is None else b0 % mod
            assert _call(emu, f"emu_{pre}_mul", a, b) == a * b * mi % mod
            assert _call(emu, f"emu_{pre}_add", a, b) == (a + b) % mod
            assert _call(emu, f"emu_{pre}_sub", a, b) == (a - b) % mod
    # dedicated squaring (doubled-tail rows): edge limbs with top bits set exercise the a_j << 1 / d_j split
    sq_edge = edge + [(1 << 254) - 1, int("ffffffff" * 8, 16), int("80000000" * 8, 16), int("ffffffff00000000" * 4, 16),
                      int("00000000ffffffff" * 4, 16), int("7fffffff80000001" * 4, 16)]
    for a0 in sq_edge + [None] * 6000:
        for mod, pre, mi in ((B.R, "fr", RI), (B.Q, "fq", QI)):
            a = rng.randrange(mod) if a0 is None else a0 % mod
            assert _call(emu, f"emu_{pre}_sqr", a) == a * a * mi % mod, hex(a)
    # fused a*b + c*d / a*b - c*d (one reduction for two products): worst-case operands first
    big = [B.Q - 1, B.Q - 2, (1 << 253) % B.Q, 0, 1]
    quads = [(w, x, y, z) for w in big[:3] for x in big[:3] for y in big for z in big]
    quads += [tuple(rng.randrange(B.Q) for _ in range(4)) for _ in range(4000)]
    for w, x, y, z in quads:
        assert _call(emu, "emu_fq_mul_add", w, x, y, z) == (w * x + y * z) * QI % B.Q
    for _ in range(2000):
        w, x, y, z = (rng.randrange(B.R) for _ in range(4))
        assert _call(emu, "emu_fr_mul_sub", w, x, y, z) == (w * x - y * z) * RI % B.R
    m = B.R - 1
    assert _call(emu, "emu_fr_mul_sub", m, m, m, m) == 0 and _call(emu, "emu_fr_mul_sub", m, m, 0, m) == m * m * RI % B.R
    for _ in range(10):
        a = rng.randrange(1, B.R)
        am = B.to_mont(a, B.R)
        assert _call(emu, "emu_fr_inv", am) == B.to_mont(pow(a, -1, B.R), B.R)
        assert _call(emu, "emu_fr_from_mont", am) == a
        assert _call(emu, "emu_fr_to_mont", a) == am
        assert _call(emu, "emu_fr_neg", a) == (-a) % B.R
        a = rng.randrange(1, B.Q)
        assert _call(emu, "emu_fq_inv", B.to_mont(a, B.Q)) == B.to_mont(pow(a, -1, B.Q), B.Q)
    assert _call(emu, "emu_fr_neg", 0) == 0


def test_batched_gcd_inversion(emu):
    """fp_inv (batched binary GCD, the inversion on the critical path of every MSM result) against Python's
    modular inverse: edge values, values of every bit length, and random residues in both fields; the plain
    binary Euclid kept as cross-check agrees."""
    rng = random.Random(7)
    for mod, pre in ((B.R, "fr"), (B.Q, "fq")):
        R2 = pow(1 << 256, 2, mod)
        vals = [1, 2, 3, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, (1 << 253), (1 << 30), (1 << 30) - 1, (1 << 62), (1 << 62) - 1,
                (1 << 64) + 1, int("ffffffff" * 7, 16), int("80000000" * 7, 16) % mod]
        vals += [(1 << k) % mod for k in range(0, 254, 7)] + [((1 << k) - 1) % mod for k in range(1, 254, 5)]
        vals += [rng.randrange(1, 1 << k) for k in range(1, 254)] + [rng.randrange(1, mod) for _ in range(3000)]
        for x in vals:
            if x % mod == 0:
                continue
            want = pow(x, -1, mod) * R2 % mod  # input taken as a Montgomery residue x = yR: the result is R/y = R^2/x
            assert _call(emu, f"emu_{pre}_inv", x) == want, hex(x)
        for x in vals[:40]:
            if x % mod:
                assert _call(emu, f"emu_{pre}_inv_euclid", x) == pow(x, -1, mod) * R2 % mod
        assert _call(emu, f"emu_{pre}_inv", 0) == 0


def _chain(emu, pts, negs, bucket=False):
    a = g1_to_mont_array(pts) if pts else np.zeros((1, 8), dtype=np.uint64)
    ng = (ctypes.c_int * max(len(pts), 1))(*negs)
    out = np.zeros(8, dtype=np.uint64)
    fn = emu.emu_g1_bucket_chain if bucket else emu.emu_g1_add_mixed_chain
    fn(a.ctypes.data_as(ctypes.c_void_p), ng, len(pts), out.ctypes.data_as(ctypes.c_void_p))
    return g1_from_mont_array(out)[0]


def _full(emu, pa, pb):
    a = g1_to_mont_array(pa) if pa else np.zeros((1, 8), dtype=np.uint64)
    b = g1_to_mont_array(pb) if pb else np.zeros((1, 8), dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    emu.emu_g1_add_full(a.ctypes.data_as(ctypes.c_void_p), len(pa), b.ctypes.data_as(ctypes.c_void_p), len(pb), out.ctypes.data_as(ctypes.c_void_p))
    return g1_from_mont_array(out)[0]


def test_device_group_law(emu):
    rng = random.Random(2)
    pts = [B.g1_mul(B.G1_GEN, rng.randrange(B.R)) for _ in range(8)]

    def ref(ps, negs):
        acc = None
        for p, ng in zip(ps, negs):
            acc = B.g1_add(acc, B.g1_neg(p) if ng else p)
        return acc

    for _ in range(20):
        k = rng.randrange(1, 9)
        ps = [rng.choice(pts) for _ in range(k)]
        negs = [rng.randrange(2) for _ in range(k)]
        assert _chain(emu, ps, negs) == ref(ps, negs)
    # the accumulate kernel's bucket walk (affine + affine first addition) gives the same sums, exceptional starts included
    for _ in range(20):
        k = rng.randrange(0, 7)
        ps = [rng.choice(pts + [None]) for _ in range(k)]
        negs = [rng.randrange(2) for _ in range(k)]
        assert _chain(emu, ps, negs, bucket=True) == ref(ps, negs)
    Q = pts[1]
    for ps, negs in [([Q, Q], [0, 0]), ([Q, Q], [1, 0]), ([Q, Q, Q], [0, 1, 1]), ([Q, Q, pts[2]], [1, 1, 0]), ([None, Q, Q], [0, 0, 0]), ([Q, None], [1, 0])]:
        assert _chain(emu, ps, negs, bucket=True) == ref(ps, negs)
    P = pts[0]
    # exceptional cases of the mixed addition: doubling, cancellation, infinity operands
    assert _chain(emu, [P, P], [0, 0]) == B.g1_add(P, P)
    assert _chain(emu, [P, P], [0, 1]) is None
    assert _chain(emu, [P, P, P], [0, 1, 0]) == P
    assert _chain(emu, [P, None, P], [0, 0, 0]) == B.g1_add(P, P)
    # full XYZZ + XYZZ addition
    assert _full(emu, pts[:3], pts[3:7]) == ref(pts[:7], [0] * 7)
    assert _full(emu, pts[:3], pts[:3]) == B.g1_mul(ref(pts[:3], [0] * 3), 2)
    assert _full(emu, [P], [B.g1_neg(P)]) is None
    assert _full(emu, [], [P]) == P and _full(emu, [P], []) == P
    for k in (0, 1, 2, 3, 5, 255, 65535, 123456):
        a = g1_to_mont_array([P])
        out = np.zeros(8, dtype=np.uint64)
        emu.emu_g1_mul_small(a.ctypes.data_as(ctypes.c_void_p), k, out.ctypes.data_as(ctypes.c_void_p))
        assert g1_from_mont_array(out)[0] == B.g1_mul(P, k)


def test_host_field_and_transcript(host_emu):
    lib = host_emu
    rng = random.Random(5)
    for n in (0, 1, 135, 136, 137, 272, 1000):
        d = bytes(rng.randrange(256) for _ in range(n))
        out = (ctypes.c_uint8 * 32)()
        lib.emu_keccak256(d, n, out)
        assert bytes(out) == keccak.keccak256(d)

    def c4(x):
        return (ctypes.c_uint64 * 4)(*B.to_limbs(x))

    def call(fn, *xs):
        out = (ctypes.c_uint64 * 4)()
        getattr(lib, fn)(*[c4(x) for x in xs], out)
        return B.from_limbs(list(out))

    for _ in range(1000):
        a, b = rng.randrange(B.R), rng.randrange(B.R)
        assert call("emu_hfr_mul", a, b) == a * b * RI % B.R
        assert call("emu_hfr_add", a, b) == (a + b) % B.R
        assert call("emu_hfr_sub", a, b) == (a - b) % B.R
        a, b = rng.randrange(B.Q), rng.randrange(B.Q)
        assert call("emu_hfq_mul", a, b) == a * b * QI % B.Q
    for e in ((0, 0), (B.R - 1, B.R - 1), (B.R - 1, 1), (1, B.R - 1)):
        assert call("emu_hfr_add", *e) == sum(e) % B.R
        assert call("emu_hfr_sub", *e) == (e[0] - e[1]) % B.R
    a = rng.randrange(1, B.R)
    assert call("emu_hfr_inv", B.to_mont(a, B.R)) == B.to_mont(pow(a, -1, B.R), B.R)
    for n in (1, 31, 32, 48, 64):
        d = bytes(rng.randrange(256) for _ in range(n))
        out = (ctypes.c_uint64 * 4)()
        lib.emu_hfr_from_bytes(d, n, out)
        assert B.from_limbs(list(out)) == B.to_mont(int.from_bytes(d, "little") % B.R, B.R)
    out8 = (ctypes.c_uint8 * 32)()
    for _ in range(10):
        P = B.g1_mul(B.G1_GEN, rng.randrange(B.R))
        arr = g1_to_mont_array([P])
        lib.emu_g1_compress(arr.ctypes.data_as(ctypes.c_void_p), out8)
        assert bytes(out8) == transcript.g1_compressed(P)
    arr = g1_to_mont_array([None])
    lib.emu_g1_compress(arr.ctypes.data_as(ctypes.c_void_p), out8)
    assert bytes(out8) == transcript.g1_compressed(None)
    m1, m2 = os.urandom(200), os.urandom(77)
    out12 = (ctypes.c_uint64 * 12)()
    lib.emu_transcript(m1, len(m1), m2, len(m2), out12)
    t = transcript.SolidityTranscript()
    t.append_message(m1)
    a = t.get_and_append_challenge()
    b = t.get_and_append_challenge()
    t.append_message(m2)
    c = t.get_and_append_challenge()
    got = [B.from_limbs(list(out12)[4 * i:4 * i + 4]) for i in range(3)]
    assert got == [B.to_mont(x, B.R) for x in (a, b, c)]
