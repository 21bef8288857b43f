import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def host_math():
    """g++ build of tests/host_math/harness.cpp (the kernels' per-Gaussian maths on the CPU)."""
    import ctypes
    import subprocess
    d = os.path.join(ROOT, "tests", "host_math")
    so = os.path.join(d, "_harness.so")
    src = os.path.join(d, "harness.cpp")
    hdr = os.path.join(ROOT, "mobgs_b200", "csrc", "gs_math.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    return ctypes.CDLL(so)
