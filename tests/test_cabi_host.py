"""A plain C++ host of the C ABI (tests/cabi/cabi_host.cpp): no Python, no torch between the caller and
libvcr_b200.so.  CPU: it compiles and links against the header and the library.  GPU: it runs and its kNN / FPS
indices equal the CPU oracle's bit for bit."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "cabi_host.cpp")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    from oracle import canon                      # builds oracle/libvcr_canon.so on first use
    assert os.path.exists(os.path.join(ROOT, "oracle", "libvcr_canon.so")), canon
    exe = str(tmp_path / "cabi_host")
    cmd = ["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"), SRC, "-o", exe,
           "-L", os.path.join(ROOT, "vcr_net_b200"), "-lvcr_b200", "-L", os.path.join(ROOT, "oracle"), "-lvcr_canon",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_cabi_host_compiles_and_links(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")
def test_cabi_host_runs_bit_exact(tmp_path):
    exe = _build(tmp_path)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.pathsep.join([os.path.join(ROOT, "vcr_net_b200"), os.path.join(ROOT, "oracle"),
                                              os.path.join(CUDA, "lib64"), env.get("LD_LIBRARY_PATH", "")])
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 mismatches" in r.stdout
