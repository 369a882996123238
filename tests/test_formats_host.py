"""Host-only helpers of the format layer (no device involved): SHA-256 gate, `Fr::rand` replay from
RNG words, `Proof` serialisation -- each against an independent statement (hashlib, the oracle)."""
import ctypes
import hashlib
import random
from ctypes import byref, c_size_t

import numpy as np

from cap_b200 import _lib, field
from cap_b200.device import _ptr
from oracle import bn254 as B
from oracle import serialize as S
from oracle.chacha import ChaChaRng


def test_sha256_matches_hashlib():
    lib = _lib.load()
    rng = random.Random(1)
    for ln in (0, 1, 55, 56, 63, 64, 65, 119, 120, 1000, 4099):
        data = bytes(rng.randrange(256) for _ in range(ln))
        out = (ctypes.c_uint8 * 32)()
        buf = (ctypes.c_uint8 * max(ln, 1)).from_buffer_copy(data or b"\0")
        assert lib.capgpu_sha256(buf, ln, out) == 0
        assert bytes(out) == hashlib.sha256(data).digest(), ln
    # the digest the reference pins for data/aztec-crs-131072.bin has this shape (src/proof/mod.rs:100)
    assert len(bytes.fromhex("6b81e75fb9c14fd0e58fb2b29e48978cdad5511503685a61f1391dc4a4fc7cbf")) == 32


def test_fr_rand_from_words_matches_the_oracle_rng():
    """ark_std::test_rng() word stream -> 17 blinders: the library's replay helper equals the oracle's
    ChaCha `Fr::rand` (itself pinned by jf-plonk's coset representatives, tests/test_oracle_hash.py)."""
    lib = _lib.load()
    rng = ChaChaRng()
    words = [rng.next_u64() for _ in range(200)]
    want, used = S.fr_rand_from_words(words, 17)
    rng2 = ChaChaRng()
    assert want == [rng2.fr_rand_mont() for _ in range(17)]
    w = np.array(words, dtype=np.uint64)
    out = np.zeros((17, 4), dtype=np.uint64)
    u = c_size_t()
    assert lib.capgpu_fr_rand_from_words(_ptr(w), len(w), _ptr(out), 17, byref(u)) == 0
    assert u.value == used and field.fr_from_raw_array(out) == want
    assert lib.capgpu_fr_rand_from_words(_ptr(w), 8, _ptr(out), 17, byref(u)) == -2  # words run out


def test_proof_serialize_matches_the_oracle_writer():
    lib = _lib.load()
    rng = random.Random(9)
    pts = [B.g1_mul(B.G1_GEN, rng.randrange(1, B.R)) for _ in range(12)] + [None]
    frs = [rng.randrange(B.R) for _ in range(10)]
    proof = {"wires_poly_comms": pts[:5], "prod_perm_poly_comm": pts[5], "split_quot_poly_comms": pts[6:11],
             "opening_proof": pts[11], "shifted_opening_proof": pts[12],
             "wires_evals": frs[:5], "wire_sigma_evals": frs[5:9], "perm_next_eval": frs[9]}
    p = _lib.Proof()
    g = field.g1_to_mont_array(pts)
    f = field.fr_to_mont_array(frs)
    raw = np.concatenate([g.reshape(-1), f.reshape(-1)]).astype(np.uint64)
    ctypes.memmove(byref(p), raw.ctypes.data, ctypes.sizeof(p))
    ln = c_size_t()
    assert lib.capgpu_proof_serialize(byref(p), None, 0, byref(ln)) == 0 and ln.value == len(S.write_proof(proof))
    out = (ctypes.c_uint8 * ln.value)()
    assert lib.capgpu_proof_serialize(byref(p), out, ln.value - 1, byref(ln)) == -2
    assert lib.capgpu_proof_serialize(byref(p), out, ln.value, byref(ln)) == 0
    assert bytes(out) == S.write_proof(proof)
    assert S.read_proof(bytes(out)) == proof
