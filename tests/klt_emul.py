"""ctypes loader of tests/klt_emul.cpp -- the device source putslam_b200/csrc/klt_point.cuh compiled for the CPU, lanes as a
loop.  TEST INFRASTRUCTURE: lets the kernel's arithmetic be checked against the oracle and the cv2 golden vectors without
a GPU.  Nothing in the product imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "klt_emul.cpp")
HDR = os.path.join(os.path.dirname(HERE), "putslam_b200", "csrc", "klt_point.cuh")
OUT = os.path.join(HERE, "_build", "libklt_emul.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            # -ffp-contract=off: one IEEE rounding per float operation, like nvcc -fmad=false on the device
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wall", "-fPIC", "-shared", "-o", OUT, SRC])
        _lib = C.CDLL(OUT)
        vp, ci, cd = C.c_void_p, C.c_int, C.c_double
        _lib.klt_emul_track.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci, ci, ci, ci, cd, ci, cd, vp, vp, ci, C.POINTER(ci)]
        _lib.klt_emul_prune.argtypes = [vp, vp, vp, ci, cd, cd, C.c_float, vp]
        _lib.klt_emul_pyrdown.argtypes = [vp, ci, ci, ci, vp]
    return _lib


def track(a, b, pts, win=7, max_level=3, max_iter=30, eps=0.01, init=None, min_eig_err=False, min_eig_thr=1e-4, specialised=True):
    L = load()
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    H, W = a.shape[:2]; cn = 1 if a.ndim == 2 else a.shape[2]
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); n = len(pts)
    cur = np.zeros((n, 2), np.float32) if init is None else np.ascontiguousarray(init, np.float32).reshape(-1, 2).copy()
    st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32); lv = C.c_int()
    flags = (1 if init is not None else 0) | (2 if min_eig_err else 0)
    r = L.klt_emul_track(a.ctypes.data, b.ctypes.data, W, H, cn, pts.ctypes.data, cur.ctypes.data, n, win, max_level, max_iter,
                         eps, flags, min_eig_thr, st.ctypes.data, err.ctypes.data, int(specialised), C.byref(lv))
    assert r == 0
    return cur, st, err, lv.value


def pyr_down(img):
    L = load()
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape[:2]; cn = 1 if img.ndim == 2 else img.shape[2]
    out = np.zeros(((H + 1) // 2, (W + 1) // 2) + ((cn,) if img.ndim == 3 else ()), np.uint8)
    L.klt_emul_pyrdown(img.ctypes.data, W, H, cn, out.ctypes.data)
    return out


def sq_threshold(d):
    """smallest double T with sqrt(T) >= d (what the library's host side passes to the prune kernel)"""
    if not d > 0: return 0.0
    if np.isinf(d): return float("inf")
    t = np.float64(d) * np.float64(d)
    while t > 0 and np.sqrt(t) >= d: t = np.nextafter(t, 0.0)
    while np.sqrt(t) < d: t = np.nextafter(t, np.inf)
    return float(t)


def prune_limit(d):
    """smallest float >= d, 0 for d <= 0 or NaN (the library's host side, prune_limit in ctx.cu)"""
    if not d > 0: return 0.0
    f = np.float32(d)
    if float(f) < d: f = np.nextafter(f, np.float32(np.inf))
    return float(f)


def prune(xy, err, status, err_thr, min_distance):
    L = load()
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); n = len(xy)
    err = np.ascontiguousarray(err, np.float32); status = np.ascontiguousarray(status, np.uint8)
    keep = np.zeros(n, np.uint8)
    L.klt_emul_prune(xy.ctypes.data, err.ctypes.data, status.ctypes.data, n, err_thr, sq_threshold(min_distance),
                     prune_limit(min_distance), keep.ctypes.data)
    return np.nonzero(keep)[0]
