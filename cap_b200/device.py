"""Host-side handles over the C ABI: context, SRS, standalone MSM / NTT.

These mirror the call sites the backend replaces: ``VariableBaseMSM::multi_scalar_mul`` (via
``KZG10::commit``) and ``Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}`` under
``PlonkKzgSnark::prove`` (/root/reference/src/proof/transfer.rs:181).  All arrays are numpy
``uint64`` in the ABI layout (see ``cap_b200.field``).
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_void_p

import numpy as np

from . import _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def _as_u64(a, cols: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim == 1:
        a = a.reshape(-1, cols)
    assert a.shape[-1] == cols, f"expected trailing dimension {cols}, got {a.shape}"
    return a


class Context:
    """One per (host thread, GPU): owns a CUDA stream and its workspaces."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = c_void_p()
        _lib.check(self.lib.capgpu_ctx_create(device, byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.capgpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_group(self, group: int):
        """Notes proved in lockstep by this context inside capgpu_prove_batch / a ProvingQueue."""
        _lib.check(self.lib.capgpu_ctx_set_group(self.h, group), self.h)

    def set_latency_mode(self, on: bool):
        _lib.check(self.lib.capgpu_ctx_set_latency_mode(self.h, int(on)), self.h)

    def sync(self):
        _lib.check(self.lib.capgpu_ctx_sync(self.h), self.h)

    @property
    def stream(self) -> int:
        return int(self.lib.capgpu_ctx_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.capgpu_launch_count(self.h))

    def calibrate(self) -> dict:
        a, b, c = c_double(), c_double(), c_double()
        _lib.check(self.lib.capgpu_calibrate(self.h, byref(a), byref(b), byref(c)), self.h)
        return {"gimad_per_s": a.value, "gimad_wide_per_s": b.value, "gfmul_per_s": c.value}

    # -- Radix2EvaluationDomain -----------------------------------------------------------
    def ntt(self, data, log_n: int, inverse: bool = False, coset: bool = False) -> np.ndarray:
        """data: (in_len, 4) or (batch, in_len, 4) Montgomery limbs; returns (.., 2^log_n, 4)."""
        a = np.ascontiguousarray(data, dtype=np.uint64)
        single = a.ndim == 2
        if single:
            a = a[None]
        batch, in_len, _ = a.shape
        out = np.empty((batch, 1 << log_n, 4), dtype=np.uint64)
        _lib.check(self.lib.capgpu_ntt(self.h, _ptr(a), in_len, _ptr(out), log_n, batch, int(inverse), int(coset)), self.h)
        return out[0] if single else out

    def ntt3(self, data, log_n: int, inverse: bool = False) -> np.ndarray:
        """Transforms on the 3 * 2^log_n-point domain g <rho> the prover evaluates the quotient on (capgpu_ntt3_dev; the
        buffers pass through torch device tensors).  Forward: (batch, in_len <= 2^log_n, 4) coefficients ->
        (batch, 3, 2^log_n, 4) values, value (k, i) = f(g rho^k w^i).  Inverse: (batch, 3, 2^log_n, 4) values ->
        (batch, 3 * 2^log_n, 4) coefficients."""
        import torch

        a = np.ascontiguousarray(data, dtype=np.uint64)
        n = 1 << log_n
        if inverse:
            a = a.reshape(-1, 3 * n, 4)
        batch, in_len, _ = a.shape
        d_in = torch.from_numpy(a.view(np.int64)).cuda()
        d_out = torch.empty((batch, 3 * n, 4), dtype=torch.int64, device="cuda")
        _lib.check(self.lib.capgpu_ntt3_dev(self.h, c_void_p(d_in.data_ptr()), in_len, c_void_p(d_out.data_ptr()), log_n, batch, int(inverse)), self.h)
        self.sync()
        out = d_out.cpu().numpy().view(np.uint64)
        return out if inverse else out.reshape(batch, 3, n, 4)


class Srs:
    """Device-resident commit key (``powers_of_g``), with the MSM's window-shifted tables."""

    def __init__(self, ctx: Context, points_xy=None, window_bits: int = 0, tau_mont=None, size: int = 0, compressed=None):
        """Upload ``points_xy`` ((n, 8) uint64), or ``compressed`` (n x 32 bytes of ark-serialize
        compressed G1 points, decompressed on the device), or generate tau^i * g on the device
        from ``tau_mont`` ((4,) uint64 Montgomery limbs) for ``size`` points."""
        self.ctx = ctx
        self.lib = ctx.lib
        h = c_void_p()
        if compressed is not None:
            raw = np.frombuffer(bytes(compressed), dtype=np.uint8)
            assert raw.size % 32 == 0, "compressed G1 points are 32 bytes each"
            _lib.check(self.lib.capgpu_srs_upload_compressed(ctx.h, _ptr(raw), raw.size // 32, window_bits, byref(h)), ctx.h)
            self.size = raw.size // 32
        elif points_xy is not None:
            pts = _as_u64(points_xy, 8)
            _lib.check(self.lib.capgpu_srs_upload(ctx.h, _ptr(pts), pts.shape[0], window_bits, byref(h)), ctx.h)
            self.size = pts.shape[0]
        else:
            t = np.ascontiguousarray(tau_mont, dtype=np.uint64).reshape(4)
            _lib.check(self.lib.capgpu_srs_setup(ctx.h, _ptr(t), size, window_bits, byref(h)), ctx.h)
            self.size = size
        self.h = h

    def export(self, n: int | None = None) -> np.ndarray:
        n = self.size if n is None else n
        out = np.zeros((n, 8), dtype=np.uint64)
        _lib.check(self.lib.capgpu_srs_export(self.ctx.h, self.h, _ptr(out), n), self.ctx.h)
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.capgpu_srs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def msm(self, scalars, mont: bool = True, base_off: int = 0, ctx: Context | None = None) -> np.ndarray:
        """scalars: (n, 4) or (batch, n, 4); returns (8,) or (batch, 8) affine x||y."""
        ctx = ctx or self.ctx
        a = np.ascontiguousarray(scalars, dtype=np.uint64)
        single = a.ndim == 2
        if single:
            a = a[None]
        batch, n, _ = a.shape
        out = np.zeros((batch, 8), dtype=np.uint64)
        _lib.check(self.lib.capgpu_msm_g1(ctx.h, self.h, base_off, _ptr(a), n, batch, int(mont), _ptr(out)), ctx.h)
        return out[0] if single else out


def msm_adhoc(ctx: Context, points_xy, scalars, mont: bool = True) -> np.ndarray:
    """``capgpu_msm_g1_adhoc``: sum_i scalars[i] * points[i] over caller-supplied affine points
    ((n, 8) Montgomery x||y) -- the G1 sums of the batched verifier.  Returns (8,) affine x||y."""
    p = np.ascontiguousarray(points_xy, dtype=np.uint64)
    a = np.ascontiguousarray(scalars, dtype=np.uint64)
    assert p.shape[0] == a.shape[0]
    out = np.zeros(8, dtype=np.uint64)
    _lib.check(ctx.lib.capgpu_msm_g1_adhoc(ctx.h, _ptr(p), _ptr(a), p.shape[0], int(mont), _ptr(out)), ctx.h)
    return out

