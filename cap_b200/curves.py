"""Host-side view of the other pairing curves the reference can be built for (BLS12-381 / BLS12-377,
/root/reference/src/config.rs:86-114): field constants, the Montgomery array layout of ark-ff `Fp384`
(6 x u64 little-endian limbs, R = 2^384) and thin wrappers over `capgpu_curve_msm_g1` /
`capgpu_curve_fq_op`.  No arithmetic happens here beyond converting integers to and from that layout."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib
from .device import Context, _ptr


@dataclass(frozen=True)
class Curve:
    name: str
    id: int      # CAPGPU_CURVE_*
    q: int       # base field modulus
    r: int       # scalar field modulus (group order)
    b: int       # y^2 = x^3 + b
    gx: int
    gy: int


BLS12_381 = Curve(
    "bls12_381", 1,
    0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
    0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 4,
    0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
    0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
BLS12_377 = Curve(
    "bls12_377", 2,
    0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001,
    0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001, 1,
    0x008848defe740a67c8fc6225bf87ff5485951e2caa9d41bb188282c8bd37cb5cd5481512ffcd394eeab9b16eb21be9ef,
    0x01914a69c5102eff1f674f5d30afeec4bd7fb348ca3e52d96d182ad44fb82305c2fe3d3634a9591afd82de55559c8ea6)
CURVES = {c.name: c for c in (BLS12_381, BLS12_377)}
R384 = 1 << 384


def fq_to_mont_array(curve: Curve, values) -> np.ndarray:
    buf = b"".join((v % curve.q * R384 % curve.q).to_bytes(48, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 6).copy()


def fq_from_mont_array(curve: Curve, arr: np.ndarray) -> list[int]:
    rinv = pow(R384, -1, curve.q)
    a = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 6)
    return [int.from_bytes(a[i].tobytes(), "little") * rinv % curve.q for i in range(a.shape[0])]


def g1_to_mont_array(curve: Curve, points) -> np.ndarray:
    """points: (x, y) integer pairs or None for infinity -> (n, 12) u64: x || y Montgomery, zeros = infinity."""
    flat = []
    for p in points:
        flat += [0, 0] if p is None else [p[0], p[1]]
    out = fq_to_mont_array(curve, flat).reshape(-1, 12)
    for i, p in enumerate(points):
        if p is None:
            out[i] = 0
    return out


def g1_from_mont_array(curve: Curve, arr: np.ndarray):
    a = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 12)
    vals = fq_from_mont_array(curve, a.reshape(-1, 6))
    return [None if not a[i].any() else (vals[2 * i], vals[2 * i + 1]) for i in range(a.shape[0])]


def fq_op(ctx: Context, curve: Curve, op: int, a: np.ndarray, b: np.ndarray | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype="<u8").reshape(-1, 6)
    out = np.zeros_like(a)
    bp = None
    if b is not None:
        b = np.ascontiguousarray(b, dtype="<u8").reshape(-1, 6)
        assert b.shape == a.shape
        bp = _ptr(b)
    _lib.check(ctx.lib.capgpu_curve_fq_op(ctx.h, curve.id, op, _ptr(a), bp, _ptr(out), a.shape[0]), ctx.h)
    return out


def msm_g1(ctx: Context, curve: Curve, points_xy: np.ndarray, scalars) -> np.ndarray:
    """sum_i s_i P_i on G1 of `curve`; points_xy (n, 12) u64 Montgomery, scalars integers < r.  Returns (12,) u64."""
    pts = np.ascontiguousarray(points_xy, dtype="<u8").reshape(-1, 12)
    sc = np.frombuffer(b"".join((s % curve.r).to_bytes(32, "little") for s in scalars), dtype="<u8").reshape(-1, 4).copy()
    assert sc.shape[0] == pts.shape[0]
    out = np.zeros(12, dtype="<u8")
    _lib.check(ctx.lib.capgpu_curve_msm_g1(ctx.h, curve.id, _ptr(pts) if len(pts) else None, _ptr(sc) if len(sc) else None, pts.shape[0], _ptr(out)), ctx.h)
    return out
