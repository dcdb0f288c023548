"""TEST INFRASTRUCTURE — loads the host-compiled debugging build of the device logic (tests/emu).
Never used by the product package."""
import os
import subprocess

from wittgenstein_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
_api = None


def api():
    global _api
    if _api is not None:
        return _api
    so = os.path.join(EMU_DIR, "libwtg_emu.so")
    csrc = os.path.join(ROOT, "wittgenstein_b200", "csrc")
    srcs = [os.path.join(EMU_DIR, "wtg_emu.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                               os.path.join(EMU_DIR, "wtg_emu.cpp")])
    _api = _lib.Api(so, "wtgemu_")
    return _api
