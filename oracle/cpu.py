"""ctypes loader for oracle/c/libcapcpu.so, the C restatement of the reference's CPU path.

Test / baseline infrastructure only (see oracle/__init__.py): used by tests/ as a second,
fast oracle at full sizes and by bench.py as the timed same-host CPU baseline ("port").
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_int, c_size_t, c_uint, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CDIR = os.path.join(HERE, "c")
SO = os.path.join(CDIR, "libcapcpu.so")


class CpuProof(ctypes.Structure):
    _fields_ = [
        ("wires_poly_comms", ctypes.c_uint64 * 8 * 5),
        ("prod_perm_poly_comm", ctypes.c_uint64 * 8),
        ("split_quot_poly_comms", ctypes.c_uint64 * 8 * 5),
        ("opening_proof", ctypes.c_uint64 * 8),
        ("shifted_opening_proof", ctypes.c_uint64 * 8),
        ("wires_evals", ctypes.c_uint64 * 4 * 5),
        ("wire_sigma_evals", ctypes.c_uint64 * 4 * 4),
        ("perm_next_eval", ctypes.c_uint64 * 4),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(CDIR, "plonk_cpu.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", CDIR, "-B", "libcapcpu.so"], check=True, stdout=subprocess.DEVNULL)
    return SO


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(SO)
        lib.capcpu_preprocess.argtypes = [c_uint, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.capcpu_prove.argtypes = [c_uint, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_int, ctypes.POINTER(CpuProof)]
        lib.capcpu_msm.argtypes = [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]
        lib.capcpu_ntt.argtypes = [c_void_p, c_uint, c_int, c_int, c_int]
        lib.capcpu_srs.argtypes = [c_void_p, c_size_t, c_int, c_void_p]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(c_void_p)


def srs(tau_mont: np.ndarray, n: int, nthreads: int = 1) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    load().capcpu_srs(_p(np.ascontiguousarray(tau_mont, dtype=np.uint64)), n, nthreads, _p(out))
    return out


def msm(srs_xy: np.ndarray, scalars: np.ndarray, mont: bool = True, nthreads: int = 1) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    sc = np.ascontiguousarray(scalars, dtype=np.uint64)
    load().capcpu_msm(_p(srs_xy), _p(sc), sc.shape[0], int(mont), nthreads, _p(out))
    return out


def ntt(data: np.ndarray, log_n: int, inverse: bool = False, coset: bool = False, nthreads: int = 1) -> np.ndarray:
    a = np.zeros((1 << log_n, 4), dtype=np.uint64)
    a[: data.shape[0]] = data
    load().capcpu_ntt(_p(a), log_n, int(inverse), int(coset), nthreads)
    return a


def preprocess(log_n: int, sel_evals: np.ndarray, sig_evals: np.ndarray, srs_xy: np.ndarray, nthreads: int = 1):
    n = 1 << log_n
    sel = np.zeros((13, n, 4), dtype=np.uint64)
    sig = np.zeros((5, n, 4), dtype=np.uint64)
    sc = np.zeros((13, 8), dtype=np.uint64)
    gc = np.zeros((5, 8), dtype=np.uint64)
    load().capcpu_preprocess(log_n, _p(np.ascontiguousarray(sel_evals)), _p(np.ascontiguousarray(sig_evals)), _p(srs_xy), nthreads,
                             _p(sel), _p(sig), _p(sc), _p(gc))
    return sel, sig, sc, gc


def prove(log_n: int, num_inputs: int, sel_coef, sig_coef, sig_evals, k, srs_xy, sel_comms, sig_comms, wires, pub, blinders,
          ext_msg: bytes = b"", nthreads: int = 1):
    proof = CpuProof()
    pubp = _p(pub) if num_inputs else None
    rc = load().capcpu_prove(log_n, num_inputs, _p(sel_coef), _p(sig_coef), _p(np.ascontiguousarray(sig_evals)), _p(k), _p(srs_xy),
                             _p(sel_comms), _p(sig_comms), _p(np.ascontiguousarray(wires)), pubp, _p(blinders), ext_msg, len(ext_msg),
                             nthreads, ctypes.byref(proof))
    return rc, proof
