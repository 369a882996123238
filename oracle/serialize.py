"""ark-serialize 0.3 ``CanonicalSerialize`` restatement for the containers the hot path reads and
writes, plus the replay-fixture container (oracle; test infrastructure only).

Reference call sites: ``store_data`` / ``load_data`` (``/root/reference/src/parameters.rs:557-592``)
serialise ``UniversalSrs`` and the note proving keys with ``CanonicalSerialize::serialize``;
``load_srs`` (``src/proof/mod.rs:74-109``) deserialises the Aztec CRS behind a SHA-256 gate;
``TransferProvingKey`` (``src/proof/transfer.rs:60``) = ``ProvingKey`` + n_inputs + n_outputs +
tree_depth.  [UPSTREAM-RECALL: ark-serialize 0.3.0 / ark-poly-commit @ cafc05e3 / jf-plonk 0.1.2 @
bcd92b2c are not vendored here; the field orders are those of the published
``derive(CanonicalSerialize)`` structs, restated; parity with upstream bytes is UNPINNED until a
fixture written by rust/parity-dump is replayed (tests/test_replay.py).]

Grammar (little-endian): usize/u64 = 8 bytes; bool = 1 byte; Option<T> = 1 tag byte (+ T);
Vec<T> = u64 length + items; BTreeMap<K,V> = u64 length + (K, V) pairs; Fr = 32-byte canonical
value; G1Affine = 32 bytes compressed (x, bit 255 = y is the larger root, bit 254 = infinity);
G2Affine = 64 bytes compressed (x.c0 | x.c1, same flag bits in the last byte; "larger" compares
c1 first, then c0); DensePolynomial = Vec<Fr> with leading zeros trimmed.

Replay fixture (``CAPFIX01``), written by ``rust/parity-dump`` from inside a CAP test and by
``write_fixture`` below from the oracle:

    magic "CAPFIX01" | u64 n_sections | sections: 8-byte ASCII tag (zero padded) | u64 length | payload

    META     4 x u64: note type (0 transfer, 1 mint, 2 freeze), n_inputs, n_outputs, tree_depth
    PK       ProvingKey::serialize (embeds the commit key and the verifying key)
    WIRES    Vec<Vec<Fr>>: the 5 witness columns, witness[wire_variables[i][j]], n values each
    PUBIN    Vec<Fr>: public inputs
    EXTMSG   Vec<u8>: extra_transcript_init_msg (empty = none)
    RNGU64   Vec<u64>: every next_u64 the prover's RNG returned during prove(), in order
    PROOF    Proof::serialize
    CHALLS   (optional) Vec<Fr>: beta, gamma, alpha, zeta, v -- only a fork exposing them can write it
"""
from __future__ import annotations

import struct

from .bn254 import Q, R
from .transcript import fr_bytes, g1_compressed

MAGIC = b"CAPFIX01"


# ---- writers ---------------------------------------------------------------------------------
def ser_u64(x: int) -> bytes:
    return struct.pack("<Q", x)


def ser_vec(items, f) -> bytes:
    return ser_u64(len(items)) + b"".join(f(x) for x in items)


def ser_poly(coeffs) -> bytes:
    c = [x % R for x in coeffs]
    while c and c[-1] == 0:
        c.pop()
    return ser_vec(c, fr_bytes)


def _f2_larger(y) -> bool:
    ny = ((-y[0]) % Q, (-y[1]) % Q)
    return (y[1], y[0]) > (ny[1], ny[0])


def g2_compressed(p) -> bytes:
    if p is None:
        b = bytearray(64)
        b[63] |= 0x40
        return bytes(b)
    (x0, x1), y = p
    b = bytearray(x0.to_bytes(32, "little") + x1.to_bytes(32, "little"))
    if _f2_larger(y):
        b[63] |= 0x80
    return bytes(b)


def write_universal_srs(powers_of_g, h, beta_h, powers_of_gamma_g=None, neg_powers_of_h=None) -> bytes:
    """``UniversalSrs`` = ``kzg10::UniversalParams`` (what ``data/aztec-crs-131072.bin`` holds)."""
    gm = powers_of_gamma_g or {}
    nh = neg_powers_of_h or {}
    out = ser_vec(powers_of_g, g1_compressed)
    out += ser_u64(len(gm)) + b"".join(ser_u64(k) + g1_compressed(v) for k, v in sorted(gm.items()))
    out += g2_compressed(h) + g2_compressed(beta_h)
    out += ser_u64(len(nh)) + b"".join(ser_u64(k) + g2_compressed(v) for k, v in sorted(nh.items()))
    return out


def write_verifying_key(vk, g, h, beta_h, gamma_g=None) -> bytes:
    out = ser_u64(vk["domain_size"]) + ser_u64(vk["num_inputs"])
    out += ser_vec(vk["sigma_comms"], g1_compressed) + ser_vec(vk["selector_comms"], g1_compressed)
    out += ser_vec(vk["k"], fr_bytes)
    out += g1_compressed(g) + g1_compressed(gamma_g) + g2_compressed(h) + g2_compressed(beta_h)
    out += b"\x00"  # is_merged
    out += b"\x00"  # plookup_vk: None
    return out


def write_proving_key(pk, powers_of_g, h, beta_h, powers_of_gamma_g=()) -> bytes:
    """``ProvingKey::serialize`` for an oracle key (``oracle.plonk.preprocess``)."""
    out = ser_vec(pk["sigmas"], ser_poly) + ser_vec(pk["selectors"], ser_poly)
    out += ser_vec(list(powers_of_g), g1_compressed) + ser_vec(list(powers_of_gamma_g), g1_compressed)
    out += write_verifying_key(pk["vk"], powers_of_g[0], h, beta_h)
    out += b"\x00"  # plookup_pk: None
    return out


def write_note_proving_key(pk_bytes: bytes, n_inputs: int, n_outputs: int, tree_depth: int) -> bytes:
    """CAP ``TransferProvingKey`` (src/proof/transfer.rs:60): proving_key, n_inputs, n_outputs, tree_depth: u8."""
    return pk_bytes + ser_u64(n_inputs) + ser_u64(n_outputs) + bytes([tree_depth])


def write_proof(proof) -> bytes:
    out = ser_vec(proof["wires_poly_comms"], g1_compressed) + g1_compressed(proof["prod_perm_poly_comm"])
    out += ser_vec(proof["split_quot_poly_comms"], g1_compressed)
    out += g1_compressed(proof["opening_proof"]) + g1_compressed(proof["shifted_opening_proof"])
    out += ser_vec(proof["wires_evals"], fr_bytes) + ser_vec(proof["wire_sigma_evals"], fr_bytes) + fr_bytes(proof["perm_next_eval"])
    out += b"\x00"  # plookup_proof: None
    return out


# ---- readers ---------------------------------------------------------------------------------
class Reader:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0

    def take(self, k: int) -> bytes:
        if self.o + k > len(self.d):
            raise ValueError("truncated blob")
        b = self.d[self.o:self.o + k]
        self.o += k
        return b

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def u8(self) -> int:
        return self.take(1)[0]

    def fr(self) -> int:
        v = int.from_bytes(self.take(32), "little")
        if v >= R:
            raise ValueError("non-canonical field element")
        return v

    def g1(self):
        return g1_decompress(self.take(32))

    def vec(self, f):
        return [f() for _ in range(self.u64())]

    def done(self) -> bool:
        return self.o == len(self.d)


def g1_decompress(b: bytes):
    v = int.from_bytes(b, "little")
    larger, inf = (v >> 255) & 1, (v >> 254) & 1
    x = v & ((1 << 254) - 1)
    if inf:
        return None
    if x >= Q:
        raise ValueError("x coordinate out of range")
    rhs = (x * x * x + 3) % Q
    y = pow(rhs, (Q + 1) // 4, Q)
    if y * y % Q != rhs:
        raise ValueError("point is not on the curve")
    if (y > Q - y) != bool(larger):
        y = Q - y
    return (x, y)


def read_proof(data: bytes) -> dict:
    r = Reader(data)
    p = {"wires_poly_comms": r.vec(r.g1), "prod_perm_poly_comm": r.g1(), "split_quot_poly_comms": r.vec(r.g1),
         "opening_proof": r.g1(), "shifted_opening_proof": r.g1(),
         "wires_evals": r.vec(r.fr), "wire_sigma_evals": r.vec(r.fr), "perm_next_eval": r.fr()}
    if r.u8() != 0 or not r.done():
        raise ValueError("unexpected plookup proof / trailing bytes")
    return p


def read_proving_key(data: bytes, allow_trailing: bool = False) -> dict:
    """Returns {"sigmas", "selectors" (coefficient lists), "powers_of_g", "vk", "consumed"}."""
    r = Reader(data)
    sigmas = [r.vec(r.fr) for _ in range(r.u64())]
    selectors = [r.vec(r.fr) for _ in range(r.u64())]
    powers = r.vec(r.g1)
    r.vec(r.g1)  # powers_of_gamma_g
    vk = {"domain_size": r.u64(), "num_inputs": r.u64()}
    vk["sigma_comms"] = r.vec(r.g1)
    vk["selector_comms"] = r.vec(r.g1)
    vk["k"] = r.vec(r.fr)
    r.take(32 + 32 + 64 + 64)  # open_key
    r.u8()  # is_merged
    if r.u8() != 0 or r.u8() != 0:
        raise ValueError("plookup keys are not supported")
    if not allow_trailing and not r.done():
        raise ValueError("trailing bytes")
    n = vk["domain_size"]
    pad = lambda c: c + [0] * (n - len(c))
    return {"sigmas": [pad(c) for c in sigmas], "selectors": [pad(c) for c in selectors], "powers_of_g": powers, "vk": vk, "consumed": r.o}


def read_universal_srs_points(data: bytes) -> list:
    r = Reader(data)
    pts = r.vec(r.g1)
    for _ in range(r.u64()):
        r.take(8 + 32)
    r.take(128)
    for _ in range(r.u64()):
        r.take(8 + 64)
    if not r.done():
        raise ValueError("trailing bytes")
    return pts


# ---- ark-ff Fr::rand from recorded RNG words ------------------------------------------------------
def fr_rand_from_words(words, count: int):
    """(Montgomery representations as drawn, words consumed): 4 words per attempt, top two bits of the
    last limb cleared, rejected if >= r (ark-ff 0.3 ``Fp256::rand``; same as ChaChaRng.fr_rand_mont)."""
    out, w = [], 0
    while len(out) < count:
        if w + 4 > len(words):
            raise ValueError("ran out of RNG words")
        limbs = list(words[w:w + 4])
        w += 4
        limbs[3] &= 0xFFFFFFFFFFFFFFFF >> 2
        v = sum(l << (64 * i) for i, l in enumerate(limbs))
        if v < R:
            out.append(v)
    return out, w


# ---- replay fixture --------------------------------------------------------------------------------
def write_fixture(meta, pk_bytes: bytes, wires, pub_inputs, ext_msg: bytes, rng_words, proof_bytes: bytes, challenges=None) -> bytes:
    """wires: 5 lists of n canonical values.  meta: (note_type, n_inputs, n_outputs, tree_depth)."""
    secs = [
        (b"META", b"".join(ser_u64(x) for x in meta)),
        (b"PK", pk_bytes),
        (b"WIRES", ser_vec(wires, lambda col: ser_vec(col, fr_bytes))),
        (b"PUBIN", ser_vec(pub_inputs, fr_bytes)),
        (b"EXTMSG", ser_u64(len(ext_msg)) + ext_msg),
        (b"RNGU64", ser_vec(rng_words, ser_u64)),
        (b"PROOF", proof_bytes),
    ]
    if challenges is not None:
        secs.append((b"CHALLS", ser_vec(challenges, fr_bytes)))
    out = MAGIC + ser_u64(len(secs))
    for tag, payload in secs:
        out += tag.ljust(8, b"\x00") + ser_u64(len(payload)) + payload
    return out


def read_fixture(data: bytes) -> dict:
    """Sections as raw payload bytes keyed by tag, plus the decoded small ones."""
    if data[:8] != MAGIC:
        raise ValueError("not a CAPFIX01 file")
    r = Reader(data)
    r.take(8)
    secs = {}
    for _ in range(r.u64()):
        tag = r.take(8).rstrip(b"\x00").decode()
        secs[tag] = r.take(r.u64())
    if not r.done():
        raise ValueError("trailing bytes")
    for need in ("META", "PK", "WIRES", "PUBIN", "EXTMSG", "RNGU64", "PROOF"):
        if need not in secs:
            raise ValueError(f"fixture lacks section {need}")
    out = {"sections": secs}
    m = Reader(secs["META"])
    out["meta"] = tuple(m.u64() for _ in range(4))
    w = Reader(secs["WIRES"])
    out["wires"] = [w.vec(w.fr) for _ in range(w.u64())]
    p = Reader(secs["PUBIN"])
    out["pub_inputs"] = p.vec(p.fr)
    e = Reader(secs["EXTMSG"])
    out["ext_msg"] = e.take(e.u64())
    g = Reader(secs["RNGU64"])
    out["rng_words"] = g.vec(g.u64)
    if "CHALLS" in secs:
        c = Reader(secs["CHALLS"])
        out["challenges"] = c.vec(c.fr)
    return out
