"""The C++ host-side mirror (include/wtg.hpp) and its parity program (tests/cpp/mirror_parity.cpp).
CPU: it compiles and links against the product library and the oracle headers.  GPU: it runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "mirror_parity")


def _build():
    import __graft_entry__ as g

    g.build()
    lib = os.path.join(ROOT, "wittgenstein_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "mirror_parity.cpp"),
                           "-L" + lib, "-lwtg_b200", "-Wl,-rpath," + lib, "-lpthread"])


def test_cpp_mirror_builds():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_mirror_parity():
    if not os.path.exists(EXE):
        _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MIRROR PARITY OK" in out.stdout, out.stdout + out.stderr
