"""BN254 fields and G1 with exact Python integers (oracle; test infrastructure only).

Follows ark-bn254 0.3.0 parameters (``Cargo.lock:81-84``) selected by the reference at
``src/config.rs:72-84`` (``Bn254`` pairing curve, ``ScalarField = Fr``,
``BaseField = Fq``) and ark-ff 0.3.0's ``Fp256`` Montgomery representation
(4 x u64 little-endian limbs, value stored as a*R mod p with R = 2^256).
Constants re-derived in SURVEY.md App. B and checked by ``tests/test_oracle_field.py``.
"""
from __future__ import annotations

# --- moduli -----------------------------------------------------------------
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47  # base field Fq
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001  # scalar field Fr

MONT_R = 1 << 256
MONT_R_Q = MONT_R % Q
MONT_R_R = MONT_R % R
MONT_RINV_Q = pow(MONT_R, -1, Q)
MONT_RINV_R = pow(MONT_R, -1, R)

FR_TWO_ADICITY = 28
FR_GENERATOR = 5  # ark-bn254 FrParameters::GENERATOR (multiplicative generator, coset shift)
FR_ROOT_OF_UNITY = pow(FR_GENERATOR, (R - 1) >> FR_TWO_ADICITY, R)  # 2^28-th primitive root

G1_B = 3
G1_GEN = (1, 2)


# --- representation helpers ---------------------------------------------------
def to_mont(x: int, mod: int) -> int:
    return (x << 256) % mod


def from_mont(x: int, mod: int) -> int:
    return (x * (MONT_RINV_Q if mod == Q else MONT_RINV_R)) % mod


def to_limbs(x: int) -> list[int]:
    """256-bit integer -> 4 little-endian u64 limbs (ark-ff BigInteger256 layout)."""
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


def fr_root_of_unity(log_n: int) -> int:
    """omega_n = g^((r-1)/n), as ark-poly Radix2EvaluationDomain::new derives group_gen."""
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_ROOT_OF_UNITY, 1 << (FR_TWO_ADICITY - log_n), R)


def inv(x: int, mod: int) -> int:
    return pow(x, -1, mod)


# --- G1: y^2 = x^3 + 3 over Fq, a = 0 ------------------------------------------
# Affine points are (x, y) tuples of canonical ints; None is the point at infinity.
def g1_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - G1_B) % Q == 0


def g1_neg(p):
    if p is None:
        return None
    return (p[0], (-p[1]) % Q)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1) * inv(2 * y1, Q) % Q
    else:
        lam = (y2 - y1) * inv(x2 - x1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    y3 = (lam * (x1 - x3) - y1) % Q
    return (x3, y3)


# Jacobian (X, Y, Z): x = X/Z^2, y = Y/Z^3; Z == 0 is infinity.  Mirrors the shape of
# ark-ec 0.3.0 short_weierstrass_jacobian::GroupProjective (double_in_place,
# add_assign_mixed) but only its mathematical result is relied upon.
JAC_INF = (1, 1, 0)


def jac_double(p):
    X, Y, Z = p
    if Z == 0:
        return p
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def jac_add_mixed(p, q):
    """p Jacobian + q affine."""
    if q is None:
        return p
    X1, Y1, Z1 = p
    x2, y2 = q
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % Q
    U2 = x2 * Z1Z1 % Q
    S2 = y2 * Z1 * Z1Z1 % Q
    if U2 == X1:
        if S2 == Y1:
            return jac_double(p)
        return JAC_INF
    H = (U2 - X1) % Q
    HH = H * H % Q
    I = 4 * HH % Q
    J = H * I % Q
    r = 2 * (S2 - Y1) % Q
    V = X1 * I % Q
    X3 = (r * r - J - 2 * V) % Q
    Y3 = (r * (V - X3) - 2 * Y1 * J) % Q
    Z3 = ((Z1 + H) * (Z1 + H) - Z1Z1 - HH) % Q
    return (X3, Y3, Z3)


def jac_add(p, q):
    X1, Y1, Z1 = p
    X2, Y2, Z2 = q
    if Z1 == 0:
        return q
    if Z2 == 0:
        return p
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return jac_double(p)
        return JAC_INF
    H = (U2 - U1) % Q
    I = 4 * H * H % Q
    J = H * I % Q
    r = 2 * (S2 - S1) % Q
    V = U1 * I % Q
    X3 = (r * r - J - 2 * V) % Q
    Y3 = (r * (V - X3) - 2 * S1 * J) % Q
    Z3 = ((Z1 + Z2) * (Z1 + Z2) - Z1Z1 - Z2Z2) * H % Q
    return (X3, Y3, Z3)


def jac_to_affine(p):
    X, Y, Z = p
    if Z == 0:
        return None
    zi = inv(Z, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def jac_from_affine(p):
    if p is None:
        return JAC_INF
    return (p[0], p[1], 1)


def g1_mul(p, k: int):
    """Double-and-add scalar multiplication (the naive cross-check for every MSM)."""
    k %= R
    if p is None or k == 0:
        return None
    acc = JAC_INF
    for bit in bin(k)[2:]:
        acc = jac_double(acc)
        if bit == "1":
            acc = jac_add_mixed(acc, p)
    return jac_to_affine(acc)


def batch_to_affine(points):
    """Montgomery-trick batch normalisation of Jacobian points."""
    zs = [p[2] for p in points]
    prefix = []
    acc = 1
    for z in zs:
        prefix.append(acc)
        if z:
            acc = acc * z % Q
    accinv = inv(acc, Q)
    out = [None] * len(points)
    for i in range(len(points) - 1, -1, -1):
        z = zs[i]
        if z == 0:
            continue
        zi = accinv * prefix[i] % Q
        accinv = accinv * z % Q
        zi2 = zi * zi % Q
        out[i] = (points[i][0] * zi2 % Q, points[i][1] * zi2 * zi % Q)
    return out


def srs_powers(tau: int, n: int, g=G1_GEN):
    """[tau^i]G for i < n, the shape of ark-poly-commit 0.3.0 KZG10::setup's powers_of_g
    (reference call site src/proof/mod.rs:59-69 -> PlonkKzgSnark::universal_setup).
    Uses a fixed-base 8-bit window table so 2^15 points take seconds, not minutes."""
    tau %= R
    # window table: T[w][d] = d * 2^(8w) * g   (Jacobian -> affine once)
    nwin = 32
    rows = []
    base = jac_from_affine(g)
    for _ in range(nwin):
        row = [JAC_INF]
        cur = JAC_INF
        for _d in range(255):
            cur = jac_add(cur, base)
            row.append(cur)
        rows.append(row)
        for _ in range(8):
            base = jac_double(base)
    flat = batch_to_affine([p for row in rows for p in row])
    table = [flat[i * 256:(i + 1) * 256] for i in range(nwin)]
    out_j = []
    s = 1
    for _ in range(n):
        acc = JAC_INF
        k = s
        w = 0
        while k:
            d = k & 0xFF
            if d:
                acc = jac_add_mixed(acc, table[w][d])
            k >>= 8
            w += 1
        out_j.append(acc)
        s = s * tau % R
    return batch_to_affine(out_j)
