import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
TAU = 0x2B7E151628AED2A6ABF7158809CF4F3C762E7160F38B4DA56A784D9045190CFE % 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _build_emu(name: str) -> ctypes.CDLL:
    src = os.path.join(ROOT, "tests", "cpu_emu", name + ".cpp")
    outdir = os.path.join(ROOT, "tests", "cpu_emu", "build")
    os.makedirs(outdir, exist_ok=True)
    so = os.path.join(outdir, name + ".so")
    deps = [src] + [os.path.join(ROOT, "cap_b200", "csrc", f) for f in ("fp.cuh", "ec.cuh", "fpn.cuh", "hostfp.h", "transcript.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def emu():
    """Device field / curve algorithms compiled for the host with emulated PTX carry semantics."""
    return _build_emu("emu")


@pytest.fixture(scope="session")
def host_emu():
    """The library's host-side C++ helpers (Montgomery arithmetic, Keccak transcript)."""
    return _build_emu("host_emu")


@pytest.fixture(scope="session")
def golden():
    return {name: json.load(open(os.path.join(GOLDEN, name + ".json"))) for name in ("ntt", "msm", "proof_n32")}


@pytest.fixture(scope="session")
def ctx():
    from cap_b200 import device
    c = device.Context(0)
    yield c
    c.close()
