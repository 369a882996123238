"""The C-ABI shared library loads and exports every symbol include/capgpu.h declares; without a
GPU it refuses to create a context (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from cap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "capgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(capgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libcapgpu.so lacks {missing}"
    assert set(_lib.EXPORTS) == set(names)


def test_rust_sys_crate_declares_the_same_entry_points():
    """rust/capgpu-sys (the binding INTEGRATION.md describes; unbuilt here) stays in step with the header."""
    text = open(os.path.join(ROOT, "rust", "capgpu-sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"\bpub fn (capgpu_[a-z0-9_]+)\s*\(", text))
    assert rust == set(_declared())


def test_error_strings():
    lib = _lib.load()
    assert lib.capgpu_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert lib.capgpu_strerror(code) not in (b"ok", b"unknown error")
    assert lib.capgpu_strerror(-99) == b"unknown error"


def test_proof_struct_layout():
    # 13 G1 (64 B) + 10 Fr (32 B), no padding: matches capgpu_proof in the header
    assert ctypes.sizeof(_lib.Proof) == 13 * 64 + 10 * 32


def test_null_arguments_are_rejected():
    lib = _lib.load()
    assert lib.capgpu_ctx_create(0, None) == -2
    assert lib.capgpu_ctx_sync(None) == -2
    assert lib.capgpu_srs_size(None) == 0
    assert lib.capgpu_launch_count(None) == 0


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.capgpu_ctx_create(0, ctypes.byref(h)) == -1  # CAPGPU_ERR_CUDA
    assert not h.value


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "cap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src or f.endswith((".cu", ".cuh", ".h")), f
