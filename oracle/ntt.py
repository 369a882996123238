"""Radix-2 NTT over BN254 Fr (oracle; test infrastructure only).

Restates ark-poly 0.3.0 ``Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}``
(``Cargo.lock:194-196``; reached from the reference through
``PlonkKzgSnark::prove`` at ``src/proof/transfer.rs:181``; domain sizes asserted at
``src/utils/mod.rs:137-193``):
  * fft:  evals[i] = sum_j c_j * omega^(i*j), natural order in and out,
          input shorter than the domain is zero-padded;
  * ifft: inverse, including the n^-1 scaling;
  * coset_fft: multiply c_j by g^j (g = Fr::multiplicative_generator() = 5) then fft;
  * coset_ifft: ifft then multiply by g^-j.
All values are canonical ints in [0, r).
"""
from __future__ import annotations

from .bn254 import R, FR_GENERATOR, fr_root_of_unity, inv


def _bitrev(i: int, bits: int) -> int:
    return int(bin(i)[2:].zfill(bits)[::-1], 2) if bits else 0


def _ntt_core(a: list[int], omega: int) -> list[int]:
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    a = [a[_bitrev(i, log_n)] for i in range(n)]
    m = 1
    while m < n:
        w_m = pow(omega, n // (2 * m), R)
        tw = [1] * m
        for k in range(1, m):
            tw[k] = tw[k - 1] * w_m % R
        for s in range(0, n, 2 * m):
            for k in range(m):
                t = tw[k] * a[s + k + m] % R
                u = a[s + k]
                a[s + k] = (u + t) % R
                a[s + k + m] = (u - t) % R
        m *= 2
    return a


def fft(coeffs: list[int], log_n: int) -> list[int]:
    n = 1 << log_n
    assert len(coeffs) <= n
    return _ntt_core(list(coeffs) + [0] * (n - len(coeffs)), fr_root_of_unity(log_n))


def ifft(evals: list[int], log_n: int) -> list[int]:
    n = 1 << log_n
    assert len(evals) <= n
    out = _ntt_core(list(evals) + [0] * (n - len(evals)), inv(fr_root_of_unity(log_n), R))
    ninv = inv(n, R)
    return [x * ninv % R for x in out]


def coset_fft(coeffs: list[int], log_n: int, shift: int = FR_GENERATOR) -> list[int]:
    out = []
    s = 1
    for c in coeffs:
        out.append(c * s % R)
        s = s * shift % R
    return fft(out, log_n)


def coset_ifft(evals: list[int], log_n: int, shift: int = FR_GENERATOR) -> list[int]:
    c = ifft(evals, log_n)
    si = inv(shift, R)
    s = 1
    for j in range(len(c)):
        c[j] = c[j] * s % R
        s = s * si % R
    return c


# ---- the 3 * 2^log_n-point domain of capgpu_ntt3_dev ------------------------------------------
# Not a reference function: jf-plonk's compute_quotient_polynomial evaluates the quotient (degree 5n + 7) on ark-poly's
# 8n-point coset; the CUDA prover uses the 6n points g <rho>, rho = 5^((r-1) / (6n)), as the three cosets g rho^k H_2n.
# This is the checker for that transform; tests/test_oracle_ntt_msm.py shows both routes interpolate the same polynomial.
def domain3_shifts(log_n: int) -> list[int]:
    rho = pow(FR_GENERATOR, (R - 1) // (3 << log_n), R)
    assert pow(rho, 3, R) == fr_root_of_unity(log_n)
    return [FR_GENERATOR * pow(rho, k, R) % R for k in range(3)]


def fft3(coeffs: list[int], log_n: int) -> list[list[int]]:
    """values[k][i] = f(g rho^k w^i) for a polynomial of degree < 3 * 2^log_n."""
    n = 1 << log_n
    out = []
    for s in domain3_shifts(log_n):
        c = pow(s, n, R)
        folded = [0] * n  # f mod (X^n - s^n) takes the same values on the coset s H
        for j, v in enumerate(coeffs):
            folded[j % n] = (folded[j % n] + v * pow(c, j // n, R)) % R
        out.append(coset_fft(folded, log_n, s))
    return out


def ifft3(values: list[list[int]], log_n: int) -> list[int]:
    """The 3 * 2^log_n coefficients from the values on the three cosets: per-coset interpolation u_k = f mod (X^n - c_k),
    c_k = (g rho^k)^n = g^n zeta^k, then the 3 x 3 Vandermonde solve t_{j + an} = (1/3) g^(-an) sum_k zeta^(-ak) u_k[j]."""
    n = 1 << log_n
    shifts = domain3_shifts(log_n)
    u = [coset_ifft(values[k], log_n, shifts[k]) for k in range(3)]
    zeta_inv = inv(pow(shifts[1] * inv(shifts[0], R) % R, n, R), R)
    gn_inv = inv(pow(shifts[0], n, R), R)
    third = inv(3, R)
    out = [0] * (3 * n)
    for a in range(3):
        scale = third * pow(gn_inv, a, R) % R
        for j in range(n):
            acc = sum(pow(zeta_inv, a * k, R) * u[k][j] for k in range(3)) % R
            out[a * n + j] = acc * scale % R
    return out


def dft_naive(coeffs: list[int], log_n: int, shift: int = 1) -> list[int]:
    """O(n^2) definition, used to pin the fast transform on tiny sizes."""
    n = 1 << log_n
    w = fr_root_of_unity(log_n)
    out = []
    for i in range(n):
        x = shift * pow(w, i, R) % R
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % R
        out.append(acc)
    return out


def poly_eval(coeffs: list[int], x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc
