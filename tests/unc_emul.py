"""ctypes loader of tests/unc_emul.cpp -- the device source putslam_b200/csrc/unc_point.cuh compiled for the CPU.
TEST INFRASTRUCTURE; nothing in the product imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "unc_emul.cpp")
HDR = os.path.join(os.path.dirname(HERE), "putslam_b200", "csrc", "unc_point.cuh")
OUT = os.path.join(HERE, "_build", "libunc_emul.so")
_lib = None


def compute(A, B, CA, CB, T, mode="euler"):
    global _lib
    if _lib is None:
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wall", "-fPIC", "-shared", "-o", OUT, SRC])
        _lib = C.CDLL(OUT)
        _lib.unc_emul.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
    CA = np.ascontiguousarray(CA, np.float64).reshape(-1, 9); CB = np.ascontiguousarray(CB, np.float64).reshape(-1, 9)
    T12 = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4].T.ravel())
    U = np.zeros(36)
    ok = _lib.unc_emul(A.ctypes.data, B.ctypes.data, CA.ctypes.data, CB.ctypes.data, len(A), T12.ctypes.data,
                       0 if mode == "euler" else 1, U.ctypes.data)
    return U.reshape(6, 6), bool(ok)
