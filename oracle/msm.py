"""G1 multi-scalar multiplication (oracle; test infrastructure only).

``msm_arkworks`` restates ark-ec 0.3.0 ``VariableBaseMSM::multi_scalar_mul``
(``msm/variable_base.rs``; ``Cargo.lock:103-105``), the routine behind every
``KZG10::commit`` of the prover the reference calls at ``src/proof/transfer.rs:181``:
  c = 3 if N < 32 else ln_without_floats(N) + 2, unsigned c-bit windows over 254 bits,
  2^c - 1 buckets per window, zero scalars skipped, unit scalars added to window 0 only,
  running-sum bucket reduction, Horner fold of the windows (c doublings each).
``msm_naive`` (double-and-add) and ``kzg_commit_tau`` (p(tau)*G for a synthetic SRS with
known tau) are the independent cross-checks.  Scalars are canonical ints in [0, r),
bases affine tuples or None.
"""
from __future__ import annotations

from .bn254 import (R, JAC_INF, jac_add, jac_add_mixed, jac_double, jac_to_affine,
                    g1_mul, g1_add, G1_GEN)
from .ntt import poly_eval


def _log2_ceil(x: int) -> int:
    # ark_std::log2: ceil(log2(x)), 0 for x <= 1
    return 0 if x <= 1 else (x - 1).bit_length()


def ln_without_floats(a: int) -> int:
    return _log2_ceil(a) * 69 // 100


def arkworks_window_bits(n: int) -> int:
    return 3 if n < 32 else ln_without_floats(n) + 2


def msm_arkworks(bases, scalars):
    n = min(len(bases), len(scalars))
    bases, scalars = bases[:n], scalars[:n]
    c = arkworks_window_bits(n)
    num_bits = 254
    window_sums = []
    for w_start in range(0, num_bits, c):
        res = JAC_INF
        buckets = [JAC_INF] * ((1 << c) - 1)
        for s, b in zip(scalars, bases):
            if s == 0:
                continue
            if s == 1:
                if w_start == 0:
                    res = jac_add_mixed(res, b)
                continue
            d = (s >> w_start) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = jac_add_mixed(buckets[d - 1], b)
        running = JAC_INF
        for bk in reversed(buckets):
            running = jac_add(running, bk)
            res = jac_add(res, running)
        window_sums.append(res)
    lowest = window_sums[0]
    total = JAC_INF
    for ws in reversed(window_sums[1:]):
        total = jac_add(total, ws)
        for _ in range(c):
            total = jac_double(total)
    return jac_to_affine(jac_add(lowest, total))


def msm_naive(bases, scalars):
    acc = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s))
    return acc


def kzg_commit_tau(coeffs, tau: int, g=G1_GEN):
    """Commitment to a coefficient vector under the synthetic SRS [tau^i]g, computed as
    p(tau) * g.  Equals KZG10::commit's MSM (ark-poly-commit 0.3.0)."""
    return g1_mul(g, poly_eval([c % R for c in coeffs], tau % R))
