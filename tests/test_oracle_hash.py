"""External known answers that pin the oracle's hash / RNG restatements."""
import hashlib
import random

from cap_b200.field import COSET_K
from oracle import bn254 as B
from oracle import chacha, keccak
from oracle.transcript import SolidityTranscript, g1_compressed


def test_keccak256_known_answers():
    # published Keccak-256 (pre-NIST padding) digests
    assert keccak.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert keccak.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"


def test_keccak_permutation_against_hashlib_sha3():
    rng = random.Random(3)
    for n in (0, 1, 135, 136, 137, 271, 272, 273, 1000):
        m = bytes(rng.randrange(256) for _ in range(n))
        assert keccak.sha3_256(m) == hashlib.sha3_256(m).digest()


def test_chacha20_rfc8439_block():
    # RFC 8439 section 2.3.2: key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00
    key = [int.from_bytes(bytes(range(32))[4 * i:4 * i + 4], "little") for i in range(8)]
    blk = chacha.chacha_block(key, [1, 0x09000000, 0x4A000000, 0], 20)
    assert blk[:4] == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3]
    assert blk[-1] == 0x4E3C50A2


def test_coset_representatives_pin_rng_and_field_sampling():
    """jf-plonk draws k_1..k_4 with Fr::rand from ChaChaRng::from_seed([0; 32]); the values are
    public (hard-coded in CAP's on-chain verifier).  Reproducing them pins the ChaCha word stream,
    next_u64 ordering, the 2-bit mask + rejection loop and the 'limbs are Montgomery' reading."""
    rng = chacha.ChaChaRng(bytes(32), rounds=20)
    ks = [B.from_mont(rng.fr_rand_mont(), B.R) for _ in range(4)]
    assert tuple(ks) == COSET_K[1:]
    # distinct cosets of the size-2^k subgroups used by CAP
    for log_n in (14, 15, 16, 17):
        n = 1 << log_n
        assert len({pow(k, n, B.R) for k in COSET_K}) == 5


def test_test_rng_is_deterministic():
    a, b = chacha.ChaChaRng(), chacha.ChaChaRng()
    xs = [a.fr_rand_mont() for _ in range(5)]
    assert xs == [b.fr_rand_mont() for _ in range(5)]
    assert all(x < B.R for x in xs) and len(set(xs)) == 5


def test_point_compression_and_transcript_determinism():
    P = B.g1_mul(B.G1_GEN, 7)
    c = g1_compressed(P)
    assert int.from_bytes(c, "little") & ((1 << 254) - 1) == P[0]
    assert (c[31] >> 7) == (1 if P[1] > B.Q - P[1] else 0)
    assert g1_compressed(B.g1_neg(P))[31] >> 7 != c[31] >> 7
    assert g1_compressed(None)[31] == 0x40
    t1, t2 = SolidityTranscript(), SolidityTranscript()
    for t in (t1, t2):
        t.append_message(b"x" * 100)
        t.append_commitment(P)
    a1, a2 = t1.get_and_append_challenge(), t2.get_and_append_challenge()
    assert a1 == a2 and 0 <= a1 < B.R
    assert t1.get_and_append_challenge() != a1  # the state advances between challenges
