"""Fiat-Shamir transcript restatement (oracle; test infrastructure only).

Restates jf-plonk 0.1.2 ``SolidityTranscript`` (jellyfish @ bcd92b2c, the transcript type
chosen by the reference at ``src/proof/transfer.rs:44,181``, ``src/proof/mint.rs:113``,
``src/proof/freeze.rs:151``) [UPSTREAM-RECALL -- source not vendored; parity unpinned]:

  * an append-only byte vector plus a 64-byte state initialised to zero; labels ignored;
  * field elements are appended as ark-serialize 0.3 canonical bytes (32 B little-endian
    of the canonical value); G1 points in ark-serialize *compressed* form (x little-endian,
    bit 7 of the last byte = "y is the larger of {y, -y}", bit 6 = infinity);
  * challenge: state = Keccak256(state|transcript|0x00) | Keccak256(state|transcript|0x01),
    challenge = Fr::from_le_bytes_mod_order(state[..48]); the transcript vector is kept;
  * ``append_vk_and_pub_input``: size_in_bits, domain_size, num_inputs as 8-byte LE usize,
    then k_i, selector commitments, sigma commitments, public inputs one by one.

Everything that is protocol convention rather than mathematics is confined to this file
(and its C++ twin ``cap_b200/csrc/transcript.h``) so it can be corrected against upstream
without touching any kernel.  The round-level C ABI takes challenges from the caller, so a
Rust host that keeps using upstream's own transcript is unaffected by any error here.
"""
from __future__ import annotations

from .bn254 import Q, R
from .keccak import keccak256


def fr_bytes(x: int) -> bytes:
    return (x % R).to_bytes(32, "little")


def g1_compressed(p) -> bytes:
    if p is None:
        b = bytearray(32)
        b[31] |= 0x40
        return bytes(b)
    x, y = p
    b = bytearray(x.to_bytes(32, "little"))
    if y > (Q - y) % Q:
        b[31] |= 0x80
    return bytes(b)


class SolidityTranscript:
    def __init__(self):
        self.transcript = bytearray()
        self.state = bytes(64)

    def append_message(self, msg: bytes):
        self.transcript += msg

    def append_commitment(self, p):
        self.append_message(g1_compressed(p))

    def append_commitments(self, ps):
        for p in ps:
            self.append_commitment(p)

    def append_field(self, x: int):
        self.append_message(fr_bytes(x))

    def append_vk_and_pub_input(self, vk, pub_input):
        self.append_message((254).to_bytes(8, "little"))
        self.append_message(int(vk["domain_size"]).to_bytes(8, "little"))
        self.append_message(int(vk["num_inputs"]).to_bytes(8, "little"))
        for k in vk["k"]:
            self.append_field(k)
        self.append_commitments(vk["selector_comms"])
        self.append_commitments(vk["sigma_comms"])
        for x in pub_input:
            self.append_field(x)

    def append_proof_evaluations(self, wires_evals, wire_sigma_evals, perm_next_eval):
        for x in wires_evals:
            self.append_field(x)
        for x in wire_sigma_evals:
            self.append_field(x)
        self.append_field(perm_next_eval)

    def get_and_append_challenge(self) -> int:
        base = self.state + bytes(self.transcript)
        h0 = keccak256(base + b"\x00")
        h1 = keccak256(base + b"\x01")
        self.state = h0 + h1
        return int.from_bytes(self.state[:48], "little") % R
