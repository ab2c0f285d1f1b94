"""TEST INFRASTRUCTURE ONLY: loads the host emulation of the CUDA kernels' logic
(tests/emul/_build/librapier_b200_emul.so, built from rapier_b200/csrc with -DRB_EMULATE)."""
import ctypes as C
import os
import subprocess

from rapier_b200._lib import declare

EMUL_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
LIB_PATH = os.path.join(EMUL_DIR, "_build", "librapier_b200_emul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", EMUL_DIR])
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib
