"""A COMPILED host over the C ABI: include/capgpu.hpp (C++ mirror of the jf-plonk interface CAP calls, the counterpart
of rust/jf-plonk-gpu for a toolchain this image has) and tests/cpp/replay_fixture.cpp, built with plain g++ against
libcapgpu.so.  not gpu: it compiles and links; gpu: it replays every CAPFIX01 fixture (blocking call, asynchronous
queue, error path) and the `Proof` bytes equal the recorded ones."""
import glob
import os
import subprocess

import pytest

from cap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "fixtures", "*.capfix")))


@pytest.fixture(scope="module")
def replay_binary():
    _lib.load()  # builds libcapgpu.so when the sources changed
    out_dir = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "replay_fixture")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "replay_fixture.cpp"),
                    "-L", os.path.join(ROOT, "cap_b200"), "-lcapgpu", "-o", exe], check=True)
    return exe


def _run(exe, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "cap_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe, *args], env=env, capture_output=True, text=True, timeout=300)


def test_cpp_host_compiles_and_links(replay_binary):
    r = _run(replay_binary)
    assert r.returncode == 2 and "usage" in r.stderr  # no fixture given; nothing touches the GPU


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cpp_host_replays_fixture(replay_binary, path):
    r = _run(replay_binary, path)
    assert r.returncode == 0 and "REPLAY_OK" in r.stdout, r.stdout + r.stderr
