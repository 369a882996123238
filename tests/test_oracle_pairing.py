"""BN254 pairing restatement (oracle/pairing.py) and the pairing-based verifier check."""
import random

from cap_b200 import synth
from oracle import bn254 as B
from oracle import pairing as P
from oracle import plonk

from conftest import TAU


def test_fq12_inverse_and_g2_generator():
    rng = random.Random(1)
    a = P.F12([rng.randrange(B.Q) for _ in range(12)])
    assert a * a.inv() == P.F12.one()
    assert P.F12([5] + [0] * 11).inv() == P.F12([pow(5, -1, B.Q)] + [0] * 11)
    assert P.g2_is_on_curve(P.G2_GEN)
    assert P.g2_mul(P.G2_GEN, B.R) is None  # prime-order subgroup
    assert P.g2_mul(P.G2_GEN, 5) == P.g2_add(P.g2_mul(P.G2_GEN, 2), P.g2_mul(P.G2_GEN, 3))


def test_pairing_is_bilinear_and_non_degenerate():
    rng = random.Random(2)
    e = P.pairing(P.G2_GEN, B.G1_GEN)
    assert e != P.F12.one() and e ** B.R == P.F12.one()
    a, b = rng.randrange(B.R), rng.randrange(B.R)
    assert P.pairing(P.g2_mul(P.G2_GEN, b), B.g1_mul(B.G1_GEN, a)) == e ** (a * b % B.R)
    assert P.pairing_product_is_one([(B.g1_mul(B.G1_GEN, a), P.G2_GEN), (B.g1_neg(B.G1_GEN), P.g2_mul(P.G2_GEN, a))])
    assert not P.pairing_product_is_one([(B.g1_mul(B.G1_GEN, a), P.G2_GEN), (B.G1_GEN, P.g2_mul(P.G2_GEN, a))])


def test_pairing_verifier_accepts_and_rejects():
    circ = synth.make_circuit(5, num_inputs=3, seed=5)
    pk = plonk.preprocess(circ, tau=TAU)
    rng = random.Random(3)
    bl = [rng.randrange(B.R) for _ in range(17)]
    proof = plonk.prove(circ, pk, bl, tau=TAU, ext_msg=b"m")
    pub = plonk.public_input(circ)
    g2_tau = P.g2_mul(P.G2_GEN, TAU)
    assert plonk.verify(pk["vk"], pub, proof, ext_msg=b"m", g2_tau=g2_tau)
    assert plonk.verify(pk["vk"], pub, proof, tau=TAU, ext_msg=b"m")  # the two checks agree
    bad = dict(proof)
    bad["perm_next_eval"] = (proof["perm_next_eval"] + 1) % B.R
    assert not plonk.verify(pk["vk"], pub, bad, ext_msg=b"m", g2_tau=g2_tau)
    assert not plonk.verify(pk["vk"], pub, proof, ext_msg=b"m", g2_tau=P.g2_mul(P.G2_GEN, TAU + 1))
