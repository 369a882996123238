"""Host-side mirror of the reference's proving interface for the hot path.

Same names, argument meaning and error behaviour as the jf-plonk 0.1.2 ``UniversalSNARK``
surface that CAP calls (``/root/reference/src/proof/mod.rs:59-69`` universal_setup,
``src/proof/transfer.rs:124-155`` preprocess, ``src/proof/transfer.rs:159-188`` prove), with the
arithmetic executed by libcapgpu on the GPU.  The circuit side (``Arithmetization``:
wire columns, permutation, public inputs) is taken from any object shaped like
``cap_b200.synth.SynthCircuit``.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_size_t, c_void_p

import numpy as np

from . import _lib, field
from .device import Context, Srs, _ptr
from .field import R

NUM_WIRES = 5
NUM_SELECTORS = 13


class PlonkError(RuntimeError):
    """Mirror of jf-plonk's PlonkError (mapped by CAP to TxnApiError::FailedSnark)."""


# ---- Arithmetization helpers (jf-relation PlonkCircuit after finalize_for_arithmetization) ----
def wire_values(circ) -> np.ndarray:
    """(5, n, 4) Montgomery limbs of witness[wire_variables[i][j]] (compute_wire_polynomials' input)."""
    wit_bytes = [((int(v) << 256) % R).to_bytes(32, "little") for v in circ.witness]
    buf = b"".join(wit_bytes[int(v)] for i in range(NUM_WIRES) for v in circ.wire_variables[i])
    return np.frombuffer(buf, dtype="<u8").reshape(NUM_WIRES, circ.n, 4).copy()


def public_input(circ) -> list[int]:
    return [int(circ.witness[int(circ.wire_variables[4][j])]) for j in range(circ.num_inputs)]


def sigma_evals(circ) -> list[list[int]]:
    """sigma_i(omega^j) = k_{i'} omega^{j'} for (i', j') the next cell of (i, j)'s variable cycle
    (jf-relation compute_wire_permutation + compute_extended_permutation_polynomials)."""
    n = circ.n
    w = field.root_of_unity(circ.log_n)
    pw = [1] * n
    for j in range(1, n):
        pw[j] = pw[j - 1] * w % R
    cells: dict[int, list] = {}
    for i in range(NUM_WIRES):
        col = circ.wire_variables[i]
        for j in range(n):
            cells.setdefault(int(col[j]), []).append((i, j))
    out = [[0] * n for _ in range(NUM_WIRES)]
    for lst in cells.values():
        for a, b in zip(lst, lst[1:] + lst[:1]):
            out[a[0]][a[1]] = circ.k[b[0]] * pw[b[1]] % R
    return out


class ProvingKey:
    """Device-resident ``ProvingKey`` (selectors, sigmas, commit key, vk)."""

    def __init__(self, ctx: Context, srs: Srs, handle, log_n: int, num_inputs: int, k):
        self.ctx, self.srs, self.h = ctx, srs, handle
        self.log_n, self.n, self.num_inputs, self.k = log_n, 1 << log_n, num_inputs, tuple(k)
        self._vk = None

    def export(self):
        n = self.n
        sel = np.zeros((NUM_SELECTORS, n, 4), dtype=np.uint64)
        sig = np.zeros((NUM_WIRES, n, 4), dtype=np.uint64)
        sc = np.zeros((NUM_SELECTORS, 8), dtype=np.uint64)
        gc = np.zeros((NUM_WIRES, 8), dtype=np.uint64)
        _lib.check(self.ctx.lib.capgpu_pk_export(self.ctx.h, self.h, _ptr(sel), _ptr(sig), _ptr(sc), _ptr(gc)), self.ctx.h)
        return sel, sig, sc, gc

    @property
    def vk(self) -> dict:
        """Verifying key as plain values (domain size, k, commitments as affine int tuples)."""
        if self._vk is None:
            sc = np.zeros((NUM_SELECTORS, 8), dtype=np.uint64)
            gc = np.zeros((NUM_WIRES, 8), dtype=np.uint64)
            _lib.check(self.ctx.lib.capgpu_pk_export(self.ctx.h, self.h, None, None, _ptr(sc), _ptr(gc)), self.ctx.h)
            self._vk = {
                "domain_size": self.n, "num_inputs": self.num_inputs, "k": list(self.k),
                "selector_comms": field.g1_from_mont_array(sc), "sigma_comms": field.g1_from_mont_array(gc),
            }
        return self._vk

    def set_lagrange(self, enable: bool):
        """Commit wire columns from their evaluations (default when the key holds the Lagrange commit
        key) or from coefficients."""
        _lib.check(self.ctx.lib.capgpu_pk_lagrange(self.h, 1 if enable else 0), self.ctx.h)

    def lagrange_bases(self, count: int | None = None):
        """[L_0..L_{n-1}, P_0, P_1, P_n, P_{n+1}] as affine int tuples."""
        count = self.n + 4 if count is None else count
        out = np.zeros((count, 8), dtype=np.uint64)
        _lib.check(self.ctx.lib.capgpu_pk_lagrange_export(self.ctx.h, self.h, _ptr(out), count), self.ctx.h)
        return field.g1_from_mont_array(out)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.capgpu_pk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def proof_to_dict(p: _lib.Proof) -> dict:
    """capgpu_proof -> canonical values (same field names as jf-plonk's ``Proof``)."""
    g = lambda a: field.g1_from_mont_array(np.frombuffer(bytes(a), dtype="<u8").reshape(-1, 8))
    f = lambda a: field.fr_from_mont_array(np.frombuffer(bytes(a), dtype="<u8").reshape(-1, 4))
    return {
        "wires_poly_comms": g(p.wires_poly_comms),
        "prod_perm_poly_comm": g(p.prod_perm_poly_comm)[0],
        "split_quot_poly_comms": g(p.split_quot_poly_comms),
        "opening_proof": g(p.opening_proof)[0],
        "shifted_opening_proof": g(p.shifted_opening_proof)[0],
        "wires_evals": f(p.wires_evals),
        "wire_sigma_evals": f(p.wire_sigma_evals),
        "perm_next_eval": f(p.perm_next_eval)[0],
    }


class PlonkKzgSnark:
    """``jf_plonk::proof_system::PlonkKzgSnark`` restricted to TurboPlonk proving."""

    @staticmethod
    def universal_setup(ctx: Context, max_degree: int, tau: int) -> Srs:
        """``universal_setup(max_degree, rng)``: powers_of_g[i] = tau^i * g, max_degree + 1 points.
        The caller supplies tau (the reference draws it from its RNG); g is the G1 generator."""
        return Srs(ctx, tau_mont=field.fr_to_mont_array([tau % R])[0], size=max_degree + 1)

    @staticmethod
    def preprocess(ctx: Context, srs: Srs, circ) -> ProvingKey:
        """``preprocess(srs, circuit)``: interpolates and commits selectors / sigmas on the GPU.
        Raises PlonkError if the SRS is smaller than domain size + 3 (``compute_universal_param_size``,
        /root/reference/src/utils/mod.rs:109-113, is domain + 2 = max_degree, i.e. domain + 3 points)."""
        sel = np.stack([field.fr_to_mont_array(s) for s in circ.selectors])
        sig = np.stack([field.fr_to_mont_array(s) for s in sigma_evals(circ)])
        k = field.fr_to_mont_array(circ.k)
        h = c_void_p()
        rc = ctx.lib.capgpu_preprocess(ctx.h, srs.h, circ.log_n, circ.num_inputs, _ptr(sel), _ptr(sig), _ptr(k), byref(h))
        if rc != 0:
            raise PlonkError(str(_lib.CapGpuError(rc, ctx.lib.capgpu_strerror(rc).decode(), ctx.lib.capgpu_last_error(ctx.h).decode())))
        return ProvingKey(ctx, srs, h, circ.log_n, circ.num_inputs, circ.k)

    @staticmethod
    def upload_proving_key(ctx: Context, srs: Srs, log_n: int, num_inputs: int, selectors, sigmas, k, selector_comms, sigma_comms) -> ProvingKey:
        """Uploads an existing jf-plonk ``ProvingKey`` (coefficient-form polynomials + vk commitments)."""
        sel = np.ascontiguousarray(selectors, dtype=np.uint64)
        sig = np.ascontiguousarray(sigmas, dtype=np.uint64)
        kk = field.fr_to_mont_array(k)
        sc = np.ascontiguousarray(selector_comms, dtype=np.uint64)
        gc = np.ascontiguousarray(sigma_comms, dtype=np.uint64)
        h = c_void_p()
        _lib.check(ctx.lib.capgpu_pk_upload(ctx.h, srs.h, log_n, num_inputs, _ptr(sel), _ptr(sig), _ptr(kk), _ptr(sc), _ptr(gc), byref(h)), ctx.h)
        return ProvingKey(ctx, srs, h, log_n, num_inputs, k)

    @staticmethod
    def prove_raw(ctx: Context, pk: ProvingKey, wires: np.ndarray, pub_inputs: np.ndarray, blinders: np.ndarray,
                  extra_transcript_init_msg: bytes | None = None) -> _lib.Proof:
        """One ``capgpu_prove`` call on ABI-layout buffers (the call a Rust shim makes)."""
        proof = _lib.Proof()
        msg = extra_transcript_init_msg or b""
        mbuf = (ctypes.c_uint8 * len(msg)).from_buffer_copy(msg) if msg else None
        rc = ctx.lib.capgpu_prove(ctx.h, pk.h, _ptr(wires), _ptr(pub_inputs) if pub_inputs.size else None, _ptr(blinders),
                                  mbuf, len(msg), byref(proof))
        if rc != 0:
            raise PlonkError(str(_lib.CapGpuError(rc, ctx.lib.capgpu_strerror(rc).decode(), ctx.lib.capgpu_last_error(ctx.h).decode())))
        return proof

    @staticmethod
    def prove(ctx: Context, circ, pk: ProvingKey, blinders_mont, extra_transcript_init_msg: bytes | None = None) -> dict:
        """``prove(rng, circuit, pk, extra_transcript_init_msg)``.  ``blinders_mont``: the 17 values the
        prover would draw from ``rng`` (Montgomery representation, as ``Fr::rand`` returns them)."""
        wires = wire_values(circ)
        pub = field.fr_to_mont_array(public_input(circ)) if circ.num_inputs else np.zeros((0, 4), dtype=np.uint64)
        bl = field.fr_raw_array(blinders_mont)
        assert bl.shape == (17, 4)
        return proof_to_dict(PlonkKzgSnark.prove_raw(ctx, pk, wires, pub, bl, extra_transcript_init_msg))


class ProvingQueue:
    """``capgpu_queue``: asynchronous proving (submit / poll / wait) over the contexts of one GPU.
    The host-side overlap point is /root/reference/src/proof/transfer.rs:167-181: a caller builds
    the witness of note k+1 while note k is being proved."""

    def __init__(self, ctxs, pk: ProvingKey, ring_slots: int = 0):
        self.lib = ctxs[0].lib
        self.ctxs, self.pk = list(ctxs), pk
        cx = (c_void_p * len(ctxs))(*[c.h for c in ctxs])
        h = c_void_p()
        _lib.check(self.lib.capgpu_queue_create(cx, len(ctxs), pk.h, ring_slots, byref(h)), ctxs[0].h)
        self.h = h

    def submit(self, wires: np.ndarray, pub_inputs: np.ndarray, blinders: np.ndarray, ext_msg: bytes = b"") -> int:
        """Copies the inputs (they may be dropped on return) and returns a ticket."""
        t = ctypes.c_uint64()
        mbuf = (ctypes.c_uint8 * len(ext_msg)).from_buffer_copy(ext_msg) if ext_msg else None
        _lib.check(self.lib.capgpu_submit(self.h, _ptr(wires), _ptr(pub_inputs) if pub_inputs.size else None, _ptr(blinders), mbuf,
                                          len(ext_msg), byref(t)))
        return t.value

    def submit_ptr(self, wires_ptr: int, pub_inputs: np.ndarray, blinders: np.ndarray, ext_msg: bytes = b"") -> int:
        t = ctypes.c_uint64()
        mbuf = (ctypes.c_uint8 * len(ext_msg)).from_buffer_copy(ext_msg) if ext_msg else None
        _lib.check(self.lib.capgpu_submit(self.h, c_void_p(wires_ptr), _ptr(pub_inputs) if pub_inputs.size else None, _ptr(blinders), mbuf,
                                          len(ext_msg), byref(t)))
        return t.value

    def poll(self, ticket: int) -> bool:
        d = ctypes.c_int()
        _lib.check(self.lib.capgpu_poll(self.h, ticket, byref(d)))
        return bool(d.value)

    def wait(self, ticket: int) -> _lib.Proof:
        """Blocks until the proof is ready; raises PlonkError with the note's status otherwise."""
        p = _lib.Proof()
        rc = self.lib.capgpu_wait(self.h, ticket, byref(p))
        if rc != 0:
            raise PlonkError(f"capgpu_wait: {self.lib.capgpu_strerror(rc).decode()} ({rc})")
        return p

    def stats(self) -> dict:
        a, b, g = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        c, w = ctypes.c_double(), ctypes.c_double()
        _lib.check(self.lib.capgpu_queue_stats(self.h, byref(a), byref(b), byref(g), byref(c), byref(w)))
        return {"submitted": a.value, "completed": b.value, "groups": g.value, "copy_ms": c.value, "wait_slot_ms": w.value}

    def close(self):
        if getattr(self, "h", None):
            self.lib.capgpu_queue_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prove_batch_raw(ctxs, pk: ProvingKey, wire_ptrs, pubs, blinders, ext_msgs=None, on_device: bool = False, raise_on_error: bool = True):
    """``capgpu_prove_batch``: independent notes over one proving key, one worker thread per
    context inside the library (the reference's rayon loop over notes,
    /root/reference/src/utils/params_builder.rs:195-233).  ``wire_ptrs``: host addresses (ints) of
    each note's 5 x n wire values; ``pubs`` / ``blinders``: lists of ABI-layout numpy arrays.
    Returns (list of _lib.Proof, per-note status codes)."""
    count = len(wire_ptrs)
    lib = ctxs[0].lib
    proofs = (_lib.Proof * count)()
    status = (ctypes.c_int * count)()
    cx = (c_void_p * len(ctxs))(*[c.h for c in ctxs])
    wp = (c_void_p * count)(*wire_ptrs)
    pp = (c_void_p * count)(*[p.ctypes.data if p.size else None for p in pubs])
    bp = (c_void_p * count)(*[b.ctypes.data for b in blinders])
    if ext_msgs is not None:
        bufs = [ctypes.create_string_buffer(m, len(m)) if m else None for m in ext_msgs]
        mp = (c_void_p * count)(*[ctypes.addressof(b) if b is not None else None for b in bufs])
        ml = (c_size_t * count)(*[len(m) if m else 0 for m in ext_msgs])
    else:
        bufs, mp, ml = None, None, None
    fn = lib.capgpu_prove_batch_dev if on_device else lib.capgpu_prove_batch
    rc = fn(cx, len(ctxs), pk.h, count, wp, pp, bp, mp, ml, proofs, status)
    if rc != 0 and raise_on_error:
        raise PlonkError(f"capgpu_prove_batch: {lib.capgpu_strerror(rc).decode()} (first failing status {rc})")
    return list(proofs), list(status)


def debug_read(ctx: Context, what: int, max_elems: int) -> list[int]:
    out = np.zeros((max_elems, 4), dtype=np.uint64)
    n = c_size_t()
    _lib.check(ctx.lib.capgpu_debug_read(ctx.h, what, _ptr(out), max_elems, byref(n)), ctx.h)
    return field.fr_from_mont_array(out[: n.value])
