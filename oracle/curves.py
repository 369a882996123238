"""TEST INFRASTRUCTURE ONLY.  Exact-integer group law of short-Weierstrass curves y^2 = x^3 + b over a prime
field: the oracle for the G1 arithmetic of the other pairing curves the reference can be configured with
(BLS12-381 / BLS12-377, /root/reference/src/config.rs:86-114; ark-ec 0.3.0
`short_weierstrass_jacobian::GroupAffine` for `ark_bls12_381::g1::Parameters` / `ark_bls12_377::g1::Parameters`).
Constants (moduli, group orders, generators of ark-bls12-381 0.3.0 / ark-bls12-377 @ 677b4ae) are checked by
tests/test_oracle_curves.py: generator on the curve, order * G = O, q = 1 mod 2^? facts used nowhere else."""
from __future__ import annotations


def add(P, Q, q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % q == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, q) % q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, q) % q
    x3 = (lam * lam - x1 - x2) % q
    return x3, (lam * (x1 - x3) - y1) % q


def neg(P, q):
    return None if P is None else (P[0], (-P[1]) % q)


def mul(k: int, P, q):
    R = None
    while k:
        if k & 1:
            R = add(R, P, q)
        P = add(P, P, q)
        k >>= 1
    return R


def on_curve(P, q, b) -> bool:
    return P is None or (P[1] * P[1] - P[0] ** 3 - b) % q == 0


def msm_naive(points, scalars, q):
    """sum_i s_i P_i by double-and-add per term (the definition VariableBaseMSM::multi_scalar_mul computes)."""
    acc = None
    for P, s in zip(points, scalars):
        acc = add(acc, mul(s, P, q), q)
    return acc
