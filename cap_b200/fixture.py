"""Reader of CAPFIX01 replay fixtures (written by rust/parity-dump inside the reference, or by
tests/golden/make_fixture.py) for the product side: `bench.py --fixture` and any host that wants to
replay a recorded `PlonkKzgSnark::prove` call (/root/reference/src/proof/transfer.rs:159-188) through
the C ABI.  Layout: INTEGRATION.md section 5.  Only containers are parsed here; the proving key and
the proof stay opaque `CanonicalSerialize` byte strings handled by the library
(capgpu_pk_load_serialized / capgpu_proof_serialize)."""
from __future__ import annotations

import ctypes
import struct
from ctypes import byref, c_size_t, c_void_p

import numpy as np

from . import _lib
from .device import Context, _ptr
from .field import R

MAGIC = b"CAPFIX01"


def _vec_fr(payload: bytes, off: int = 0):
    (cnt,) = struct.unpack_from("<Q", payload, off)
    off += 8
    raw = np.frombuffer(payload, dtype="<u8", count=cnt * 4, offset=off).reshape(cnt, 4)
    return raw, off + 32 * cnt


def _canonical_to_mont(raw: np.ndarray) -> np.ndarray:
    vals = [int.from_bytes(raw[i].tobytes(), "little") for i in range(raw.shape[0])]
    if any(v >= R for v in vals):
        raise ValueError("non-canonical field element in fixture")
    buf = b"".join(((v << 256) % R).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


class Fixture:
    def __init__(self, path: str):
        data = open(path, "rb").read()
        if data[:8] != MAGIC:
            raise ValueError(f"{path}: not a CAPFIX01 file")
        (nsec,) = struct.unpack_from("<Q", data, 8)
        off = 16
        self.sections = {}
        for _ in range(nsec):
            tag = data[off:off + 8].rstrip(b"\0").decode()
            (ln,) = struct.unpack_from("<Q", data, off + 8)
            self.sections[tag] = data[off + 16:off + 16 + ln]
            off += 16 + ln
        if off != len(data):
            raise ValueError(f"{path}: trailing bytes")
        self.path = path
        self.meta = struct.unpack("<4Q", self.sections["META"])
        self.pk_bytes = self.sections["PK"]
        self.proof_bytes = self.sections["PROOF"]
        w = self.sections["WIRES"]
        (cols,) = struct.unpack_from("<Q", w, 0)
        if cols != 5:
            raise ValueError("fixture must hold 5 witness columns")
        off, out = 8, []
        for _ in range(5):
            raw, off = _vec_fr(w, off)
            out.append(_canonical_to_mont(raw))
        self.wires = np.stack(out)  # (5, n, 4) Montgomery
        raw, _ = _vec_fr(self.sections["PUBIN"])
        self.pub_inputs = _canonical_to_mont(raw) if raw.shape[0] else np.zeros((0, 4), dtype=np.uint64)
        e = self.sections["EXTMSG"]
        (ln,) = struct.unpack_from("<Q", e, 0)
        self.ext_msg = e[8:8 + ln]
        g = self.sections["RNGU64"]
        (cnt,) = struct.unpack_from("<Q", g, 0)
        self.rng_words = np.frombuffer(g, dtype="<u8", count=cnt, offset=8).copy()

    def blinders(self, lib):
        """17 x 4 Montgomery limbs re-drawn from the recorded RNG words the way `Fr::rand` consumes
        them; tries 17 draws, then 13 (a revision that does not mask the split quotient)."""
        for count in (17, 13):
            bl = np.zeros((17, 4), dtype=np.uint64)
            used = c_size_t()
            rc = lib.capgpu_fr_rand_from_words(_ptr(self.rng_words), len(self.rng_words), _ptr(bl), count, byref(used))
            if rc == 0 and used.value == len(self.rng_words):
                return bl, count
        raise ValueError("recorded RNG words match neither 17 nor 13 Fr::rand draws")

    def load_key(self, ctx: Context):
        """capgpu_pk handle (with its embedded commit key) from the PK section."""
        buf = (ctypes.c_uint8 * len(self.pk_bytes)).from_buffer_copy(self.pk_bytes)
        h = c_void_p()
        _lib.check(ctx.lib.capgpu_pk_load_serialized(ctx.h, buf, len(self.pk_bytes), None, byref(h)), ctx.h)
        return h


def proof_bytes(lib, proof: _lib.Proof) -> bytes:
    ln = c_size_t()
    out = (ctypes.c_uint8 * 1024)()
    _lib.check(lib.capgpu_proof_serialize(byref(proof), out, 1024, byref(ln)))
    return bytes(out[: ln.value])
