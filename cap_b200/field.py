"""BN254 field constants and host-side buffer packing for the C ABI.

The ABI moves field elements exactly as ark-ff 0.3.0 stores them (the layout the
reference's types carry, ``src/config.rs:78-83``): 4 little-endian u64 limbs holding the
MONTGOMERY form a*2^256 mod p; a slice of N elements is a contiguous N x 4 u64 array.
G1 affine points cross as x||y (8 u64), with the all-zero pattern meaning infinity.
"""
from __future__ import annotations

import numpy as np

Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_RINV_R = pow(1 << 256, -1, R)
_RINV_Q = pow(1 << 256, -1, Q)
FR_GENERATOR = 5
FR_TWO_ADICITY = 28
_ROOT = pow(FR_GENERATOR, (R - 1) >> FR_TWO_ADICITY, R)

# Coset representatives k_0..k_4 for 5 wire types (jf-plonk ``compute_coset_representatives``,
# the values hard-coded as COSET_K1..K4 in the CAP on-chain verifier) [UPSTREAM-RECALL].
COSET_K = (
    1,
    0x2F8DD1F1A7583C42C4E12A44E110404C73CA6C94813F85835DA4FB7BB1301D4A,
    0x1EE678A0470A75A6EAA8FE837060498BA828A3703B311D0F77F010424AFEB025,
    0x2042A587A90C187B0A087C03E29C968B950B1DB26D5C82D666905A6895790C0A,
    0x2E2B91456103698ADF57B799969DEA1C8F739DA5D8D40DD3EB9222DB7C81E881,
)


def root_of_unity(log_n: int) -> int:
    return pow(_ROOT, 1 << (FR_TWO_ADICITY - log_n), R)


def fr_to_mont_array(values) -> np.ndarray:
    """canonical ints -> (N, 4) uint64 Montgomery limbs."""
    buf = b"".join(((int(v) << 256) % R).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def fr_from_mont_array(arr) -> list[int]:
    b = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [int.from_bytes(b[i:i + 32], "little") * _RINV_R % R for i in range(0, len(b), 32)]


def fr_raw_array(values) -> np.ndarray:
    """ints -> (N, 4) uint64 limbs WITHOUT Montgomery conversion (canonical scalars)."""
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def fr_from_raw_array(arr) -> list[int]:
    b = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def g1_to_mont_array(points) -> np.ndarray:
    """affine (x, y) tuples / None -> (N, 8) uint64, x||y Montgomery, infinity = zeros."""
    parts = []
    for p in points:
        if p is None:
            parts.append(bytes(64))
        else:
            parts.append(((p[0] << 256) % Q).to_bytes(32, "little") + ((p[1] << 256) % Q).to_bytes(32, "little"))
    return np.frombuffer(b"".join(parts), dtype="<u8").reshape(-1, 8).copy()


def g1_from_mont_array(arr):
    b = np.ascontiguousarray(arr, dtype="<u8").tobytes()
    out = []
    for i in range(0, len(b), 64):
        if b[i:i + 64] == bytes(64):
            out.append(None)
            continue
        x = int.from_bytes(b[i:i + 32], "little") * _RINV_Q % Q
        y = int.from_bytes(b[i + 32:i + 64], "little") * _RINV_Q % Q
        out.append((x, y))
    return out
