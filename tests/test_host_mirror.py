"""The C++ host-side mirror (go-tfhe_b200/host/gotfhe.hpp) over the C ABI: compiles everywhere, runs on the GPU."""
import importlib
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_host_mirror")


def _build():
    T = importlib.import_module("go-tfhe_b200")
    T.build()
    libdir = os.path.join(ROOT, "go-tfhe_b200", "lib")
    src = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
    hdr = os.path.join(ROOT, "go-tfhe_b200", "host", "gotfhe.hpp")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", BIN, src, "-L" + libdir, "-ltfhe_b200", "-ltfhe_b200_client",
                               "-Wl,-rpath," + libdir, "-Wl,-rpath-link," + libdir, "-L/usr/local/cuda/lib64", "-lcudart"])
    return BIN


def test_host_mirror_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_host_mirror_truth_tables_on_gpu():
    res = subprocess.run([_build(), "80"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failures" in res.stdout
