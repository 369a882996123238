"""The C restatement of the reference's CPU path (oracle/c/plonk_cpu.c, the timed CPU baseline)
against the Python big-int oracle: NTT, SRS, MSM (arkworks Pippenger), preprocess, full proofs."""
import random

import numpy as np
import pytest

from cap_b200 import _lib, field, plonk as host, synth
from oracle import bn254 as B
from oracle import cpu, msm as omsm, ntt as ontt, plonk as oplonk

from conftest import TAU


def test_c_ntt():
    rng = random.Random(1)
    for log_n in (1, 4, 9, 12):
        x = [rng.randrange(B.R) for _ in range(1 << log_n)]
        xm = field.fr_to_mont_array(x)
        assert field.fr_from_mont_array(cpu.ntt(xm, log_n)) == ontt.fft(x, log_n)
        assert field.fr_from_mont_array(cpu.ntt(xm, log_n, True, False, nthreads=3)) == ontt.ifft(x, log_n)
        assert field.fr_from_mont_array(cpu.ntt(xm, log_n, False, True, nthreads=2)) == ontt.coset_fft(x, log_n)
        assert field.fr_from_mont_array(cpu.ntt(xm, log_n, True, True)) == ontt.coset_ifft(x, log_n)


def test_c_srs_and_msm(golden):
    tau_m = field.fr_to_mont_array([TAU])[0]
    srs = cpu.srs(tau_m, 200, nthreads=4)
    g = golden["msm"]
    exp = [None if p is None else (int(p[0], 16), int(p[1], 16)) for p in g["srs"]]
    assert field.g1_from_mont_array(srs[:len(exp)]) == exp
    rng = random.Random(2)
    for n in (5, 31, 32, 200):
        for sc in ([rng.randrange(B.R) for _ in range(n)], [rng.randrange(3) for _ in range(n)], [B.R - 1] * n, [0] * n):
            want = omsm.kzg_commit_tau(sc, TAU)
            assert field.g1_from_mont_array(cpu.msm(srs, field.fr_to_mont_array(sc), True, nthreads=4))[0] == want
            assert field.g1_from_mont_array(cpu.msm(srs, field.fr_raw_array(sc), False, nthreads=1))[0] == want
    for c in g["cases"]:
        sc = [int(x, 16) for x in c["scalars"]]
        r = c["result"]
        assert field.g1_from_mont_array(cpu.msm(srs, field.fr_to_mont_array(sc)))[0] == (None if r is None else (int(r[0], 16), int(r[1], 16)))


@pytest.mark.parametrize("log_n,nin", [(5, 3), (8, 4)])
def test_c_prover_matches_python_oracle(log_n, nin):
    rng = random.Random(log_n)
    circ = synth.make_circuit(log_n, nin, seed=log_n)
    n = circ.n
    srs = cpu.srs(field.fr_to_mont_array([TAU])[0], n + 3, nthreads=4)
    opk = oplonk.preprocess(circ, tau=TAU)
    sel_e = np.stack([field.fr_to_mont_array(s) for s in circ.selectors])
    sig_e = np.stack([field.fr_to_mont_array(s) for s in host.sigma_evals(circ)])
    sel, sig, sc, gc = cpu.preprocess(log_n, sel_e, sig_e, srs, nthreads=2)
    assert [field.fr_from_mont_array(s) for s in sel] == opk["selectors"]
    assert [field.fr_from_mont_array(s) for s in sig] == opk["sigmas"]
    assert field.g1_from_mont_array(sc) == opk["vk"]["selector_comms"]
    assert field.g1_from_mont_array(gc) == opk["vk"]["sigma_comms"]
    bl = [rng.randrange(B.R) for _ in range(17)]
    want = oplonk.prove(circ, opk, bl, tau=TAU, ext_msg=b"hi")
    pub = field.fr_to_mont_array(host.public_input(circ))
    for nthreads in (1, 3):
        rc, pr = cpu.prove(log_n, nin, sel, sig, sig_e, field.fr_to_mont_array(circ.k), srs, sc, gc, host.wire_values(circ), pub,
                           field.fr_to_mont_array(bl), b"hi", nthreads=nthreads)
        assert rc == 0
        assert host.proof_to_dict(_lib.Proof.from_buffer_copy(bytes(pr))) == want
    # unsatisfied witness -> WrongQuotientPolyDegree (-3), like the CUDA path
    w = host.wire_values(circ).copy()
    w[4, nin + 2, 0] ^= 1
    rc, _ = cpu.prove(log_n, nin, sel, sig, sig_e, field.fr_to_mont_array(circ.k), srs, sc, gc, w, pub, field.fr_to_mont_array(bl), b"hi", nthreads=2)
    assert rc == -3
