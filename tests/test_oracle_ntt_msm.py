"""Oracle NTT / MSM: fast algorithms against their definitions and the golden fixtures."""
import random

import pytest

from oracle import bn254 as B
from oracle import msm, ntt

from conftest import TAU


def _h(xs):
    return [int(x, 16) for x in xs]


def _pt(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


@pytest.mark.parametrize("log_n", [1, 2, 4, 6])
def test_fft_matches_definition(log_n):
    rng = random.Random(log_n)
    n = 1 << log_n
    for ln in {n, max(1, n // 2 + 1)}:
        x = [rng.randrange(B.R) for _ in range(ln)]
        assert ntt.fft(x, log_n) == ntt.dft_naive(x, log_n)
        assert ntt.coset_fft(x, log_n) == ntt.dft_naive(x, log_n, shift=B.FR_GENERATOR)
        full = x + [0] * (n - ln)
        assert ntt.ifft(ntt.fft(x, log_n), log_n) == full
        assert ntt.coset_ifft(ntt.coset_fft(x, log_n), log_n) == full


def test_ntt_golden(golden):
    for v in golden["ntt"]:
        x = _h(v["input"])
        assert ntt.fft(x, v["log_n"]) == _h(v["fft"])
        assert ntt.ifft(x, v["log_n"]) == _h(v["ifft"])
        assert ntt.coset_fft(x, v["log_n"]) == _h(v["coset_fft"])
        assert ntt.coset_ifft(x, v["log_n"]) == _h(v["coset_ifft"])


def test_quotient_domain_equivalence():
    """N(x) / Z_H(x) evaluated point-wise on the 6n-point domain (three cosets of 2n points, oracle fft3 / ifft3 = the checker of
    capgpu_ntt3_dev) interpolates the same quotient as on jf-plonk's 8n-point coset, for a quotient of the PLONK degree 5n + 7."""
    rng = random.Random(6)
    log_n = 4
    n = 1 << log_n
    t = [rng.randrange(B.R) for _ in range(5 * n + 8)]
    numer = [0] * (len(t) + n)  # t * (X^n - 1)
    for j, v in enumerate(t):
        numer[j + n] = (numer[j + n] + v) % B.R
        numer[j] = (numer[j] - v) % B.R
    zh = [B.R - 1] + [0] * (n - 1) + [1]
    # 8n-point coset (degree of the numerator 6n + 7 < 8n)
    ev_n, ev_z = ntt.coset_fft(numer, log_n + 3), ntt.coset_fft(zh, log_n + 3)
    q8 = ntt.coset_ifft([a * B.inv(b, B.R) % B.R for a, b in zip(ev_n, ev_z)], log_n + 3)
    assert q8[:len(t)] == t and not any(q8[len(t):])
    # 6n points: the numerator does not fit (6n + 7 >= 6n), its values on the domain still do
    v_n, v_z = ntt.fft3(numer, log_n + 1), ntt.fft3(zh, log_n + 1)
    w, shifts = B.fr_root_of_unity(log_n + 1), ntt.domain3_shifts(log_n + 1)
    for k, i in [(0, 0), (1, 5), (2, 2 * n - 1)]:
        assert v_n[k][i] == ntt.poly_eval(numer, shifts[k] * pow(w, i, B.R) % B.R)
    q6 = ntt.ifft3([[a * B.inv(b, B.R) % B.R for a, b in zip(v_n[k], v_z[k])] for k in range(3)], log_n + 1)
    assert q6[:len(t)] == t and not any(q6[len(t):])
    # Z_H has period 2 along a coset of 2n points
    assert all(v_z[k][i] == v_z[k][i % 2] for k in range(3) for i in range(2 * n))


def test_arkworks_window_size():
    # SURVEY.md App. A.1 (ark-ec 0.3.0 ln_without_floats): values for the prover's MSM sizes
    assert msm.arkworks_window_bits(31) == 3
    assert msm.arkworks_window_bits(1 << 12) == 10
    assert msm.arkworks_window_bits((1 << 15) + 2) == 13
    assert msm.arkworks_window_bits(1 << 17) == 13
    assert msm.arkworks_window_bits((1 << 17) + 2) == 14


def test_msm_three_ways():
    rng = random.Random(9)
    n = 35
    srs = B.srs_powers(TAU, n)
    for sc in ([rng.randrange(B.R) for _ in range(n)], [rng.randrange(3) for _ in range(n)], [0] * n, [B.R - 1] * n):
        a = msm.msm_arkworks(srs, sc)
        assert a == msm.msm_naive(srs, sc)
        assert a == msm.kzg_commit_tau(sc, TAU)
    # fewer than 32 points takes the c = 3 branch
    assert msm.msm_arkworks(srs[:5], [1, 2, 3, 4, 5]) == msm.msm_naive(srs[:5], [1, 2, 3, 4, 5])
    # repeated base: bucket accumulation must handle P + P and P - P
    P = srs[3]
    assert msm.msm_arkworks([P, P, B.g1_neg(P)] * 12, [7] * 36) == B.g1_mul(P, 7 * 12)


def test_msm_golden(golden):
    g = golden["msm"]
    assert int(g["tau"], 16) == TAU
    srs = [_pt(p) for p in g["srs"]]
    assert srs == B.srs_powers(TAU, len(srs))
    for c in g["cases"]:
        assert msm.msm_arkworks(srs, _h(c["scalars"])) == _pt(c["result"])
