"""Pins the oracle's BN254 constants and group law (ark-bn254 0.3.0)."""
import random

from oracle import bn254 as B


def test_moduli_and_montgomery_constants():
    assert B.Q.bit_length() == 254 and B.R.bit_length() == 254
    assert (-pow(B.Q, -1, 1 << 64)) % (1 << 64) == 0x87D20782E4866389
    assert (-pow(B.R, -1, 1 << 64)) % (1 << 64) == 0xC2E1F593EFFFFFFF
    assert B.MONT_R_Q == 0x0E0A77C19A07DF2F666EA36F7879462C0A78EB28F5C70B3DD35D438DC58F0D9D
    assert B.MONT_R_R == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert pow(2, 512, B.Q) == 0x06D89F71CAB8351F47AB1EFF0A417FF6B5E71911D44501FBF32CFC5B538AFA89
    assert pow(2, 512, B.R) == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7


def test_two_adicity_and_generator():
    assert (B.R - 1) % (1 << 28) == 0 and (B.R - 1) % (1 << 29) != 0
    w = B.FR_ROOT_OF_UNITY
    assert w == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert pow(w, 1 << 28, B.R) == 1 and pow(w, 1 << 27, B.R) != 1
    # 5 is a quadratic non-residue, hence generates the 2-Sylow part of the multiplicative group
    assert pow(5, (B.R - 1) // 2, B.R) == B.R - 1
    # Montgomery form of the generator 5, the limb constant ark-bn254 ships
    assert B.to_limbs(B.to_mont(5, B.R)) == [0x1B0D0EF99FFFFFE6, 0xEABA68A3A32A913F, 0x47D8EB76D8DD0689, 0x15D0085520F5BBC3]
    for log_n in (1, 5, 15, 18, 20):
        wn = B.fr_root_of_unity(log_n)
        assert pow(wn, 1 << log_n, B.R) == 1 and pow(wn, 1 << (log_n - 1), B.R) == B.R - 1


def test_g1_group_law():
    G = B.G1_GEN
    assert B.g1_is_on_curve(G)
    assert B.g1_mul(G, B.R) is None  # prime order r, cofactor 1
    assert B.g1_mul(G, B.R - 1) == B.g1_neg(G)
    rng = random.Random(1)
    a, b = rng.randrange(B.R), rng.randrange(B.R)
    P, Qp = B.g1_mul(G, a), B.g1_mul(G, b)
    assert B.g1_is_on_curve(P) and B.g1_is_on_curve(Qp)
    assert B.g1_add(P, Qp) == B.g1_mul(G, (a + b) % B.R)
    assert B.g1_add(P, P) == B.g1_mul(G, 2 * a % B.R)
    assert B.g1_add(P, B.g1_neg(P)) is None
    assert B.g1_add(P, None) == P
    # Jacobian formulas agree with the affine chord-tangent law
    J = B.jac_add(B.jac_from_affine(P), B.jac_from_affine(Qp))
    assert B.jac_to_affine(J) == B.g1_add(P, Qp)
    assert B.jac_to_affine(B.jac_double(B.jac_from_affine(P))) == B.g1_add(P, P)
    assert B.jac_to_affine(B.jac_add_mixed(B.jac_from_affine(P), P)) == B.g1_add(P, P)
    assert B.jac_add_mixed(B.jac_from_affine(P), B.g1_neg(P))[2] == 0


def test_srs_powers_and_limbs():
    tau = 123456789
    srs = B.srs_powers(tau, 5)
    assert srs[0] == B.G1_GEN
    for i, p in enumerate(srs):
        assert p == B.g1_mul(B.G1_GEN, pow(tau, i, B.R))
    x = random.Random(2).randrange(1 << 256)
    assert B.from_limbs(B.to_limbs(x)) == x
    assert B.from_mont(B.to_mont(12345, B.R), B.R) == 12345
    assert B.from_mont(B.to_mont(12345, B.Q), B.Q) == 12345


def test_g1_known_answers_from_the_alt_bn128_precompile_vectors():
    """External pin of the curve arithmetic: 2G and 3G on alt_bn128 (= ark-bn254's G1, generator
    (1, 2)) as published with the EIP-196 ecAdd / ecMul precompile test vectors."""
    g2 = (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
          0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    g3 = (0x0769BF9AC56BEA3FF40232BCB1B6BD159315D84715B8E679F2D355961915ABF0,
          0x2AB799BEE0489429554FDB7C8D086475319E63B40B9C5B57CDF1FF3DD9FE2261)
    assert B.g1_add(B.G1_GEN, B.G1_GEN) == g2
    assert B.g1_mul(B.G1_GEN, 2) == g2
    assert B.g1_add(g2, B.G1_GEN) == g3 and B.g1_mul(B.G1_GEN, 3) == g3
    assert B.g1_is_on_curve(g2) and B.g1_is_on_curve(g3)
