"""ctypes binding of libcapgpu.so (the C ABI declared in include/capgpu.h).

Fails loudly when the CUDA library is missing: there is no CPU fallback and this module
never imports ``oracle``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcapgpu.so")

# every symbol include/capgpu.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "capgpu_strerror", "capgpu_last_error", "capgpu_ctx_create", "capgpu_ctx_destroy", "capgpu_ctx_sync",
    "capgpu_ctx_stream", "capgpu_srs_upload", "capgpu_srs_setup", "capgpu_srs_export", "capgpu_msm_g1_dev", "capgpu_ntt_dev", "capgpu_ntt3_dev", "capgpu_srs_destroy", "capgpu_srs_size", "capgpu_msm_g1",
    "capgpu_ntt", "capgpu_pk_upload", "capgpu_preprocess", "capgpu_pk_export", "capgpu_pk_destroy", "capgpu_pk_lagrange", "capgpu_pk_lagrange_export", "capgpu_msm_g1_adhoc",
    "capgpu_prove", "capgpu_job_begin", "capgpu_job_round1", "capgpu_job_round2", "capgpu_job_round3",
    "capgpu_job_round4", "capgpu_job_round5", "capgpu_job_end", "capgpu_debug_read", "capgpu_launch_count",
    "capgpu_calibrate", "capgpu_prove_dev", "capgpu_profile_enable", "capgpu_profile_read", "capgpu_ctx_set_latency_mode", "capgpu_g1_sum_dev", "capgpu_srs_upload_compressed", "capgpu_prove_batch",
    "capgpu_ctx_set_group", "capgpu_pk_info", "capgpu_prove_batch_dev", "capgpu_queue_create", "capgpu_queue_destroy", "capgpu_submit", "capgpu_poll", "capgpu_wait", "capgpu_queue_stats",
    "capgpu_sha256", "capgpu_srs_load_serialized", "capgpu_pk_load_serialized", "capgpu_proof_serialize", "capgpu_fr_rand_from_words", "capgpu_msm_g1_dev_part",
    "capgpu_curve_msm_g1", "capgpu_curve_fq_op", "capgpu_msm_g1_dev_part_xyzz", "capgpu_g1_sum_xyzz_dev",
    "capgpu_msm_g1_dev_part_peer", "capgpu_g1_sum_xyzz_wait_dev",
]


class CapGpuError(RuntimeError):
    def __init__(self, code: int, text: str, detail: str = ""):
        super().__init__(f"capgpu error {code}: {text}" + (f" [{detail}]" if detail else ""))
        self.code = code


class Proof(ctypes.Structure):
    """Mirror of ``capgpu_proof`` (jf-plonk ``Proof``: 13 G1 + 10 Fr)."""
    _fields_ = [
        ("wires_poly_comms", c_uint64 * 8 * 5),
        ("prod_perm_poly_comm", c_uint64 * 8),
        ("split_quot_poly_comms", c_uint64 * 8 * 5),
        ("opening_proof", c_uint64 * 8),
        ("shifted_opening_proof", c_uint64 * 8),
        ("wires_evals", c_uint64 * 4 * 5),
        ("wire_sigma_evals", c_uint64 * 4 * 4),
        ("perm_next_eval", c_uint64 * 4),
    ]


_lib = None


def load() -> ctypes.CDLL:
    """Loads libcapgpu.so.  `build.build()` runs first: it is a no-op when the stamp beside the
    library equals the digest of the current sources and flags, and rebuilds otherwise, so a stale
    library is never loaded silently (CAPGPU_NO_BUILD=1 skips the check, e.g. on a box without nvcc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("CAPGPU_NO_BUILD") or not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: the CUDA extension is required (no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    P64 = POINTER(c_uint64)
    sig = {
        "capgpu_strerror": (c_char_p, [c_int]),
        "capgpu_last_error": (c_char_p, [c_void_p]),
        "capgpu_ctx_create": (c_int, [c_int, POINTER(c_void_p)]),
        "capgpu_ctx_destroy": (None, [c_void_p]),
        "capgpu_ctx_sync": (c_int, [c_void_p]),
        "capgpu_ctx_stream": (c_void_p, [c_void_p]),
        "capgpu_ctx_set_latency_mode": (c_int, [c_void_p, c_int]),
        "capgpu_srs_upload": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_void_p)]),
        "capgpu_srs_upload_compressed": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_void_p)]),
        "capgpu_srs_setup": (c_int, [c_void_p, c_void_p, c_size_t, c_int, POINTER(c_void_p)]),
        "capgpu_srs_export": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
        "capgpu_msm_g1_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_void_p]),
        "capgpu_ntt_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_uint, c_size_t, c_int, c_int]),
        "capgpu_msm_g1_dev_part": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_size_t, c_size_t, c_void_p]),
        "capgpu_g1_sum_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
        "capgpu_msm_g1_dev_part_xyzz": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_size_t, c_size_t, c_void_p]),
        "capgpu_g1_sum_xyzz_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
        "capgpu_msm_g1_dev_part_peer": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_size_t, c_size_t, c_void_p, c_void_p, c_size_t, c_uint]),
        "capgpu_g1_sum_xyzz_wait_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_uint, c_void_p]),
        "capgpu_srs_destroy": (None, [c_void_p]),
        "capgpu_srs_size": (c_size_t, [c_void_p]),
        "capgpu_msm_g1": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_int, c_void_p]),
        "capgpu_ntt": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_uint, c_size_t, c_int, c_int]),
        "capgpu_ntt3_dev": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_uint, c_size_t, c_int]),
        "capgpu_pk_upload": (c_int, [c_void_p, c_void_p, c_uint, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
        "capgpu_preprocess": (c_int, [c_void_p, c_void_p, c_uint, c_size_t, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
        "capgpu_pk_export": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        "capgpu_pk_destroy": (None, [c_void_p]),
        "capgpu_msm_g1_adhoc": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
        "capgpu_pk_lagrange": (c_int, [c_void_p, c_int]),
        "capgpu_pk_lagrange_export": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
        "capgpu_prove": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(Proof)]),
        "capgpu_prove_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(Proof)]),
        "capgpu_profile_enable": (c_int, [c_void_p, c_int]),
        "capgpu_profile_read": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_uint64), POINTER(c_double)]),
        "capgpu_prove_batch": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        "capgpu_prove_batch_dev": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        "capgpu_ctx_set_group": (c_int, [c_void_p, c_int]),
        "capgpu_pk_info": (c_int, [c_void_p, POINTER(c_uint), POINTER(c_size_t), c_void_p]),
        "capgpu_queue_create": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_void_p)]),
        "capgpu_queue_destroy": (None, [c_void_p]),
        "capgpu_submit": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(c_uint64)]),
        "capgpu_poll": (c_int, [c_void_p, c_uint64, POINTER(c_int)]),
        "capgpu_wait": (c_int, [c_void_p, c_uint64, POINTER(Proof)]),
        "capgpu_sha256": (c_int, [c_void_p, c_size_t, c_void_p]),
        "capgpu_srs_load_serialized": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, POINTER(c_void_p)]),
        "capgpu_pk_load_serialized": (c_int, [c_void_p, c_void_p, c_size_t, POINTER(c_size_t), POINTER(c_void_p)]),
        "capgpu_proof_serialize": (c_int, [POINTER(Proof), c_void_p, c_size_t, POINTER(c_size_t)]),
        "capgpu_fr_rand_from_words": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, POINTER(c_size_t)]),
        "capgpu_queue_stats": (c_int, [c_void_p, POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_double), POINTER(c_double)]),
        "capgpu_job_begin": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
        "capgpu_job_round1": (c_int, [c_void_p, c_void_p, c_void_p]),
        "capgpu_job_round2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
        "capgpu_job_round3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        "capgpu_job_round4": (c_int, [c_void_p, c_void_p, c_void_p]),
        "capgpu_job_round5": (c_int, [c_void_p, c_void_p, c_void_p]),
        "capgpu_job_end": (None, [c_void_p]),
        "capgpu_debug_read": (c_int, [c_void_p, c_int, c_void_p, c_size_t, POINTER(c_size_t)]),
        "capgpu_launch_count": (c_uint64, [c_void_p]),
        "capgpu_calibrate": (c_int, [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_double)]),
        "capgpu_curve_msm_g1": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
        "capgpu_curve_fq_op": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue  # reported by tests/test_abi.py; callers get AttributeError on use
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, ctx_handle=None):
    if rc == 0:
        return
    lib = load()
    text = lib.capgpu_strerror(rc).decode()
    detail = lib.capgpu_last_error(ctx_handle).decode() if ctx_handle else ""
    raise CapGpuError(rc, text, detail)
