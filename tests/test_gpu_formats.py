"""The reference's on-disk formats -> device: `UniversalSrs` and `ProvingKey`
`CanonicalSerialize` blobs written by the oracle's serializer (oracle/serialize.py) are loaded through
the C ABI and must give the same commit key / proving key / proofs as the plain upload paths; the
SHA-256 gate and malformed blobs are refused.  Mirrors `load_srs` (/root/reference/src/proof/mod.rs:
74-109) and `load_data` (src/parameters.rs:572-580)."""
import ctypes
import hashlib
import random
from ctypes import byref, c_size_t, c_void_p

import numpy as np
import pytest

from cap_b200 import _lib, field, plonk, synth
from cap_b200.device import _ptr
from oracle import bn254 as B
from oracle import pairing, plonk as oplonk, serialize as S

from conftest import TAU

pytestmark = pytest.mark.gpu


def _buf(b: bytes):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(b)


def test_universal_srs_blob(ctx):
    lib = ctx.lib
    n = 70
    powers = B.srs_powers(TAU, n)
    powers[5] = None  # a point at infinity survives the round trip
    h, beta_h = pairing.G2_GEN, pairing.g2_mul(pairing.G2_GEN, TAU)
    blob = S.write_universal_srs(powers, h, beta_h, powers_of_gamma_g={0: powers[1], 3: powers[2]}, neg_powers_of_h={1: beta_h})
    assert S.read_universal_srs_points(blob) == powers
    dig = hashlib.sha256(blob).digest()
    hsrs = c_void_p()
    _lib.check(lib.capgpu_srs_load_serialized(ctx.h, _buf(blob), len(blob), _buf(dig), 0, 0, byref(hsrs)), ctx.h)
    assert lib.capgpu_srs_size(hsrs) == n
    out = np.zeros((n, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_srs_export(ctx.h, hsrs, _ptr(out), n), ctx.h)
    assert field.g1_from_mont_array(out) == powers
    lib.capgpu_srs_destroy(hsrs)
    # trim: only the first max_points powers are kept
    _lib.check(lib.capgpu_srs_load_serialized(ctx.h, _buf(blob), len(blob), None, 40, 0, byref(hsrs)), ctx.h)
    assert lib.capgpu_srs_size(hsrs) == 40
    lib.capgpu_srs_destroy(hsrs)
    # the integrity gate of src/proof/mod.rs:98-107
    bad = bytearray(dig)
    bad[0] ^= 1
    assert lib.capgpu_srs_load_serialized(ctx.h, _buf(blob), len(blob), _buf(bytes(bad)), 0, 0, byref(hsrs)) == -2
    assert b"sha256" in lib.capgpu_last_error(ctx.h)
    # malformed blobs: truncated, trailing bytes, a point off the curve
    assert lib.capgpu_srs_load_serialized(ctx.h, _buf(blob[:-1]), len(blob) - 1, None, 0, 0, byref(hsrs)) == -2
    assert lib.capgpu_srs_load_serialized(ctx.h, _buf(blob + b"\0"), len(blob) + 1, None, 0, 0, byref(hsrs)) == -2
    off = bytearray(blob)
    off[8 + 32 * 7] ^= 1
    rc = lib.capgpu_srs_load_serialized(ctx.h, _buf(bytes(off)), len(off), None, 0, 0, byref(hsrs))
    if rc == 0:  # the flipped x may still be on the curve: then the point simply differs
        lib.capgpu_srs_destroy(hsrs)
    else:
        assert rc == -2


@pytest.mark.parametrize("log_n,nin", [(5, 2), (8, 6)])
def test_proving_key_blob(ctx, log_n, nin):
    lib = ctx.lib
    circ = synth.make_circuit(log_n, num_inputs=nin, seed=40 + log_n)
    n = circ.n
    powers = B.srs_powers(TAU, n + 3)
    opk = oplonk.preprocess(circ, tau=TAU)
    h, beta_h = pairing.G2_GEN, pairing.g2_mul(pairing.G2_GEN, TAU)
    blob = S.write_proving_key(opk, powers, h, beta_h)
    assert S.read_proving_key(blob)["vk"] == opk["vk"]
    # CAP's TransferProvingKey = ProvingKey + n_inputs + n_outputs + tree_depth (src/proof/transfer.rs:60)
    wrapped = S.write_note_proving_key(blob, 2, 2, 26)
    hpk = c_void_p()
    used = c_size_t()
    _lib.check(lib.capgpu_pk_load_serialized(ctx.h, _buf(wrapped), len(wrapped), byref(used), byref(hpk)), ctx.h)
    assert used.value == len(blob) and wrapped[used.value:] == S.ser_u64(2) + S.ser_u64(2) + bytes([26])
    assert lib.capgpu_pk_load_serialized(ctx.h, _buf(wrapped), len(wrapped), None, byref(c_void_p())) == -2  # trailing bytes
    lg, ni = ctypes.c_uint(), c_size_t()
    kk = np.zeros((5, 4), dtype=np.uint64)
    _lib.check(lib.capgpu_pk_info(hpk, byref(lg), byref(ni), _ptr(kk)))
    assert (lg.value, ni.value) == (log_n, nin) and field.fr_from_mont_array(kk) == list(circ.k)
    sel = np.zeros((13, n, 4), dtype=np.uint64)
    sig = np.zeros((5, n, 4), dtype=np.uint64)
    sc = np.zeros((13, 8), dtype=np.uint64)
    gc = np.zeros((5, 8), dtype=np.uint64)
    _lib.check(lib.capgpu_pk_export(ctx.h, hpk, _ptr(sel), _ptr(sig), _ptr(sc), _ptr(gc)), ctx.h)
    assert [field.fr_from_mont_array(s) for s in sel] == opk["selectors"]
    assert [field.fr_from_mont_array(s) for s in sig] == opk["sigmas"]
    assert field.g1_from_mont_array(sc) == opk["vk"]["selector_comms"] and field.g1_from_mont_array(gc) == opk["vk"]["sigma_comms"]
    # a proof under the loaded key (its embedded commit key included) equals the oracle's, byte for byte
    rng = random.Random(log_n)
    bl = [rng.randrange(B.R) for _ in range(17)]
    want = oplonk.prove(circ, opk, bl, tau=TAU, ext_msg=b"from-blob")
    wires = plonk.wire_values(circ)
    pub = field.fr_to_mont_array(plonk.public_input(circ))
    blm = field.fr_raw_array([B.to_mont(b, B.R) for b in bl])
    proof = _lib.Proof()
    msg = _buf(b"from-blob")
    _lib.check(lib.capgpu_prove(ctx.h, hpk, _ptr(wires), _ptr(pub), _ptr(blm), msg, 9, byref(proof)), ctx.h)
    assert plonk.proof_to_dict(proof) == want
    ln = c_size_t()
    out = (ctypes.c_uint8 * 1024)()
    _lib.check(lib.capgpu_proof_serialize(byref(proof), out, 1024, byref(ln)))
    assert bytes(out[: ln.value]) == S.write_proof(want)
    lib.capgpu_pk_destroy(hpk)
    # malformed keys
    for bad in (blob[:-3], blob[:8] + b"\x07" + blob[9:], blob[:-1] + b"\x01"):
        assert lib.capgpu_pk_load_serialized(ctx.h, _buf(bad), len(bad), None, byref(c_void_p())) == -2
