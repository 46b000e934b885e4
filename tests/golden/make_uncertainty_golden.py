"""Golden vectors for TransformEst::computeUncertainty / computeUncertaintyG2O (reference
include/putslam/TransformEst/transformEst.h:29-144 and :147-272).

The reference cannot be compiled here (Eigen is absent), but the two functions are straight-line code: per point a list
of `dgdTheta(r,c) += <expression>;` / `dgdX(<row>,c) = <expression>;` statements (machine-generated derivatives), then
`uncertainty = dgdTheta^-1 * dgdX^T * Cx * dgdX * dgdTheta^-1`.  This script READS those statements from the reference
header where it lies (nothing of it is copied into the repository), evaluates them with Python's math library on seeded
inputs, and stores inputs and outputs in uncertainty_ref.npz -- the reference's own formulas, run here.

    python tests/golden/make_uncertainty_golden.py          (needs /root/reference; the .npz is committed)
"""
import math
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = "/root/reference/include/putslam/TransformEst/transformEst.h"


def quaternion_from_rotation(m):
    """Eigen::Quaternion(Matrix3) (the published trace / largest-diagonal method) -> (w, x, y, z)"""
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        t = math.sqrt(t + 1.0); w = 0.5 * t; t = 0.5 / t
        return w, (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t
    i = 0
    if m[1, 1] > m[0, 0]: i = 1
    if m[2, 2] > m[i, i]: i = 2
    j = (i + 1) % 3; k = (j + 1) % 3
    t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
    q = [0.0, 0.0, 0.0]
    q[i] = 0.5 * t; t = 0.5 / t
    w = (m[k, j] - m[j, k]) * t
    q[j] = (m[j, i] + m[i, j]) * t; q[k] = (m[k, i] + m[i, k]) * t
    return w, q[0], q[1], q[2]


def function_body(src, name):
    start = src.index("& " + name + "(")
    end = src.index("return uncertainty;", start)
    return src[start:end]


def statements(body):
    acc, setx, sym = [], [], []
    for line in body.splitlines():
        line = line.strip()
        m = re.match(r"dgdTheta\((\d),(\d)\)\s*\+=\s*(.*);$", line)
        if m: acc.append((int(m.group(1)), int(m.group(2)), m.group(3))); continue
        m = re.match(r"dgdTheta\((\d),(\d)\)\s*=\s*dgdTheta\((\d),(\d)\);$", line)
        if m: sym.append(tuple(int(g) for g in m.groups())); continue
        m = re.match(r"dgdX\((.*?),(\d)\)\s*=\s*(.*);$", line)
        if m: setx.append((m.group(1), int(m.group(2)), m.group(3)))
    return acc, setx, sym


def evaluate(name, A, B, CA, CB, T):
    src = open(HEADER).read()
    acc, setx, sym = statements(function_body(src, name))
    assert len(acc) >= 21 and len(setx) == 36, (name, len(acc), len(setx))
    n = len(A)
    env = {k: getattr(math, k) for k in ("sin", "cos", "pow")}
    w, qx, qy, qz = quaternion_from_rotation(T[:3, :3])
    if name == "computeUncertainty":
        q0, q1, q2, q3 = w, qx, qy, qz
        env.update(roll=math.atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2)), pitch=math.asin(2 * (q0 * q2 - q3 * q1)),
                   yaw=math.atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3)))
    else:
        env.update(qx=qx, qy=qy, qz=qz, qw=w)
    env.update(x=T[0, 3], y=T[1, 3], z=T[2, 3])
    H = np.zeros((6, 6)); G = np.zeros((6 * n, 6)); Cx = np.zeros((6 * n, 6 * n))
    for i in range(n):
        env.update(xa=A[i, 0], ya=A[i, 1], za=A[i, 2], xb=B[i, 0], yb=B[i, 1], zb=B[i, 2])
        for r, c, e in acc:
            H[r, c] += eval(e, {"__builtins__": {}}, env)
        for row, c, e in setx:
            G[eval(row.replace("setA.rows()", str(n)), {"__builtins__": {}}, {"i": i}), c] = eval(e, {"__builtins__": {}}, env)
        Cx[3 * i:3 * i + 3, 3 * i:3 * i + 3] = CA[i]
        Cx[3 * n + 3 * i:3 * n + 3 * i + 3, 3 * n + 3 * i:3 * n + 3 * i + 3] = CB[i]
    for r, c, r2, c2 in sym:
        H[r, c] = H[r2, c2]
    k = 1.0 / n
    G = k * G; H = k * H
    Hi = np.linalg.inv(H)
    return Hi @ G.T @ Cx @ G @ Hi, H, G


def rot(rpy):
    r, p, y = rpy
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def cases():
    rng = np.random.default_rng(83)
    out = {}
    names = []
    for ci, (n, rpy, noise) in enumerate(((100, (0.2, -0.1, 0.3), 0.01), (12, (-2.5, 1.2, 2.9), 0.05), (40, (0.0, 0.0, 0.0), 0.0),
                                          (7, (3.0, -0.4, -3.0), 0.02), (250, (1.0, 0.7, -2.0), 0.005))):
        B = rng.uniform(-1.5, 1.5, (n, 3))
        T = np.eye(4); T[:3, :3] = rot(rpy); T[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        A = B @ T[:3, :3].T + T[:3, 3] + rng.normal(0, noise, (n, 3))
        L = rng.normal(0, 0.01, (2, n, 3, 3))
        CA = L[0] @ L[0].transpose(0, 2, 1) + 1e-6 * np.eye(3); CB = L[1] @ L[1].transpose(0, 2, 1) + 1e-6 * np.eye(3)
        name = f"c{ci}"
        names.append(name)
        out.update({name + "_A": A, name + "_B": B, name + "_CA": CA, name + "_CB": CB, name + "_T": T})
        for fn, tag in (("computeUncertainty", "euler"), ("computeUncertaintyG2O", "quat")):
            U, H, G = evaluate(fn, A, B, CA, CB, T)
            out[f"{name}_{tag}_U"] = U; out[f"{name}_{tag}_H"] = H; out[f"{name}_{tag}_G"] = G
    out["names"] = np.array(names)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "uncertainty_ref.npz"), **cases())
    print("uncertainty golden vectors written")
