"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/wtg.h declares (no compute call is made without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "wtg.h")).read()
    return sorted(set(re.findall(r"\b(wtg_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_loads_and_exports_header_symbols():
    import __graft_entry__ as g

    g.build()
    from wittgenstein_b200 import _lib

    lib = C.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wtg.h but not exported"
    # and the Python mirror binds exactly the declared entry points
    assert sorted(_lib.Api().symbols()) == syms


def test_no_cpu_fallback_in_product_package():
    """The product must not import the oracle or the host debugging build."""
    pkg = os.path.join(ROOT, "wittgenstein_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in src and "libwtg_oracle" not in src and "libwtg_emu" not in src, f
                assert "#include \"../../oracle" not in src and "oracle/" not in src.replace("oracle/ ", ""), f


def test_create_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from wittgenstein_b200 import Network, WtgError

    with pytest.raises(WtgError):
        Network()


def test_jni_shim_compiles_against_the_header():
    """bindings/jni/wtg_jni.c (the reference-side binding) type-checks against include/wtg.h; without a JDK
    the check uses a minimal stub of <jni.h> (tests/cpp/jni_stub)."""
    import subprocess

    jni = None
    jh = os.environ.get("JAVA_HOME")
    if jh and os.path.exists(os.path.join(jh, "include", "jni.h")):
        jni = ["-I" + os.path.join(jh, "include"), "-I" + os.path.join(jh, "include", "linux")]
    inc = jni or ["-I" + os.path.join(ROOT, "tests", "cpp", "jni_stub")]
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Werror"] + inc + ["-I" + os.path.join(ROOT, "include"),
                                                                                os.path.join(ROOT, "bindings", "jni", "wtg_jni.c")])
    # every native method of NativeNetwork.java has its wrapper
    java = open(os.path.join(ROOT, "bindings", "jni", "NativeNetwork.java")).read()
    c = open(os.path.join(ROOT, "bindings", "jni", "wtg_jni.c")).read()
    for name in re.findall(r"native\s+[\w\[\]]+\s+(\w+)\(", java):
        assert f"NN({name})" in c, name
