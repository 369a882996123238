"""BN254 optimal-ate pairing with exact Python integers (oracle; test infrastructure only).

Lets the oracle's verifier restatement check the KZG opening equation of a proof the way the
reference's verifier does -- e(A, [tau]_2) = e(B, [1]_2) (jf-plonk ``PlonkKzgSnark::verify`` /
``batch_verify``, reached from ``/root/reference/src/proof/transfer.rs:192-212`` and
``src/lib.rs:517``; curve = ark-bn254 ``Bn254``, ``src/config.rs:77-84``) -- without using the
trapdoor tau on the G1 side.  Textbook construction: Fq12 = Fq[w]/(w^12 - 18 w^6 + 82), G2 on the
sextic twist y^2 = x^3 + 3/(9+u) over Fq2 = Fq[u]/(u^2+1), Miller loop over 6x+2 with
x = 4965661367192848881, final exponentiation by (q^12-1)/r.  Pinned by bilinearity,
non-degeneracy and the group order of the standard G2 generator (tests/test_oracle_pairing.py).
Slow (about a second per pairing): used on a handful of proofs only.
"""
from __future__ import annotations

from .bn254 import Q, R

ATE_LOOP_COUNT = 29793968203157093288  # 6x + 2
LOG_ATE = 63

# standard generator of G2 (ark-bn254 g2::Parameters / EIP-197), coordinates c0 + c1*u
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


# ---- Fq2 (pairs), used only for G2 scalar multiplication --------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)
def f2_scalar(a, k): return (a[0] * k % Q, a[1] * k % Q)


TWIST_B = f2_mul((3, 0), f2_inv((9, 1)))  # 3 / (9 + u)


def g2_is_on_curve(p) -> bool:
    if p is None:
        return True
    x, y = p
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), TWIST_B)) == (0, 0)


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return None
        lam = f2_mul(f2_scalar(f2_mul(x1, x1), 3), f2_inv(f2_scalar(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(p, k: int):
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


# ---- Fq12 as polynomials in w modulo w^12 - 18 w^6 + 82 ---------------------------------------
class F12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % Q for x in c]

    @staticmethod
    def one():
        return F12([1] + [0] * 11)

    @staticmethod
    def zero():
        return F12([0] * 12)

    def __eq__(self, o):
        return self.c == o.c

    def __add__(self, o):
        return F12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return F12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return F12([-a for a in self.c])

    def scale(self, k: int):
        return F12([a * k for a in self.c])

    def __mul__(self, o):
        t = [0] * 23
        for i, a in enumerate(self.c):
            if a:
                for j, b in enumerate(o.c):
                    t[i + j] += a * b
        for i in range(22, 11, -1):  # w^12 = 18 w^6 - 82
            top = t[i]
            if top:
                t[i - 6] += 18 * top
                t[i - 12] -= 82 * top
        return F12(t[:12])

    def __pow__(self, e: int):
        r, b = F12.one(), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def inv(self):
        """Extended Euclid on polynomials over Fq against the modulus w^12 - 18 w^6 + 82:
        invariant r_i = s_i * self (mod modulus)."""
        def deg(p):
            d = len(p) - 1
            while d > 0 and p[d] == 0:
                d -= 1
            return d

        r0, s0 = [82, 0, 0, 0, 0, 0, -18 % Q, 0, 0, 0, 0, 0, 1], [0] * 14
        r1, s1 = self.c[:] + [0], [1] + [0] * 13
        while True:
            d1 = deg(r1)
            if d1 == 0:
                if r1[0] == 0:
                    raise ZeroDivisionError("Fq12 inverse of zero")
                k = pow(r1[0], -1, Q)
                return F12([x * k for x in s1[:12]])
            while True:
                d0 = deg(r0)
                if d0 < d1 or (d0 == 0 and r0[0] == 0):
                    break
                f = r0[d0] * pow(r1[d1], -1, Q) % Q
                sh = d0 - d1
                for i in range(d1 + 1):
                    r0[i + sh] = (r0[i + sh] - f * r1[i]) % Q
                for i in range(14 - sh):
                    s0[i + sh] = (s0[i + sh] - f * s1[i]) % Q
            r0, r1, s0, s1 = r1, r0, s1, s0

    def __truediv__(self, o):
        return self * o.inv()


W = F12([0, 1] + [0] * 10)
W2, W3 = W * W, W * W * W


def _cast_g1(p):
    return (F12([p[0]] + [0] * 11), F12([p[1]] + [0] * 11))


def _twist(q):
    (x0, x1), (y0, y1) = q
    nx = F12([(x0 - 9 * x1) % Q] + [0] * 5 + [x1] + [0] * 5)
    ny = F12([(y0 - 9 * y1) % Q] + [0] * 5 + [y1] + [0] * 5)
    return (nx * W2, ny * W3)


def _double(p):
    x, y = p
    lam = (x * x).scale(3) / y.scale(2)
    nx = lam * lam - x.scale(2)
    return (nx, lam * (x - nx) - y)


def _add(p, q):
    x1, y1 = p
    x2, y2 = q
    if x1 == x2 and y1 == y2:
        return _double(p)
    lam = (y2 - y1) / (x2 - x1)
    nx = lam * lam - x1 - x2
    return (nx, lam * (x1 - nx) - y1)


def _line(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not (x1 == x2):
        lam = (y2 - y1) / (x2 - x1)
        return lam * (xt - x1) - (yt - y1)
    if y1 == y2:
        lam = (x1 * x1).scale(3) / y1.scale(2)
        return lam * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q2, p1) -> F12:
    """Un-exponentiated pairing value for q2 in G2 (Fq2 affine) and p1 in G1 (affine ints)."""
    if q2 is None or p1 is None:
        return F12.one()
    Qt, P = _twist(q2), _cast_g1(p1)
    Rr, f = Qt, F12.one()
    for i in range(LOG_ATE, -1, -1):
        f = f * f * _line(Rr, Rr, P)
        Rr = _double(Rr)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _line(Rr, Qt, P)
            Rr = _add(Rr, Qt)
    Q1 = (Qt[0] ** Q, Qt[1] ** Q)
    nQ2 = (Q1[0] ** Q, -(Q1[1] ** Q))
    f = f * _line(Rr, Q1, P)
    Rr = _add(Rr, Q1)
    f = f * _line(Rr, nQ2, P)
    return f


def final_exponentiation(f: F12) -> F12:
    return f ** ((Q ** 12 - 1) // R)


def pairing(q2, p1) -> F12:
    return final_exponentiation(miller_loop(q2, p1))


def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 for [(p1, q2), ...] (one shared final exponentiation)."""
    f = F12.one()
    for p1, q2 in pairs:
        f = f * miller_loop(q2, p1)
    return final_exponentiation(f) == F12.one()
