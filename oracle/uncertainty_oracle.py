"""CPU restatement (numpy) of TransformEst::computeUncertainty / computeUncertaintyG2O -- TEST INFRASTRUCTURE ONLY.

Reference: include/putslam/TransformEst/transformEst.h:29-144 (Euler angles) and :147-272 (quaternion vector part); called
by demos/demoKabsch.cpp:655,731,1028.  The reference spells the derivatives out as machine-generated
scalar expressions; here they are derived, which is what the kernel computes too:

    cost      J(theta; a, b) = sum_i |r_i|^2,  r_i = a_i - R(theta) b_i - t,   theta = (t, rotation parameters)
    g         = dJ/dtheta = -2 sum_i M_i^T r_i,            M_i = d(R b_i + t)/dtheta = [I | D_1 b_i, D_2 b_i, D_3 b_i]
    dg/dtheta = 2 sum_i (M_i^T M_i - S_i),                 S_i[3+k, 3+l] = r_i . (D_kl b_i)     (the full Hessian)
    dg/da_i   = -2 M_i          (3 x 6, rows = coordinates of a_i)
    dg/db_i   : translation columns 2 R^T ... i.e. [j, k] = 2 R[k, j];  rotation column k:  -2 D_k^T r_i + 2 R^T D_k b_i
    U         = H^-1 (sum_i Ga_i^T CA_i Ga_i + Gb_i^T CB_i Gb_i) H^-1,  H and the G's scaled by 1/n as in the reference

    Euler:  R = Rz(yaw) Ry(pitch) Rx(roll), parameters (roll, pitch, yaw) read off the quaternion of T (:32-39)
    G2O:    R = the quaternion rotation matrix written with 1 - 2(..) on the diagonal, differentiated w.r.t. qx, qy, qz
            with qw held constant (that is what the reference's expressions are)
D_k = dR/dparam_k, D_kl = d2R/dparam_k dparam_l.  Pinned against the reference's own expressions evaluated by
tests/golden/make_uncertainty_golden.py (tests/golden/uncertainty_ref.npz) to 1e-9 relative.
"""
import math

import numpy as np


def quaternion_from_rotation(m):
    """Eigen::Quaternion(Matrix3): trace / largest-diagonal method -> (w, x, y, z)"""
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


def euler_derivatives(w, x, y, z):
    roll = math.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = math.asin(2 * (w * y - z * x))
    yaw = math.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return euler_derivatives_from_angles(roll, pitch, yaw)


def euler_derivatives_from_angles(roll, pitch, yaw):
    def axis(a, k):       # rotation about axis k and its first / second derivative
        c, s = math.cos(a), math.sin(a)
        i, j = (k + 1) % 3, (k + 2) % 3
        R = np.eye(3); R[i, i] = c; R[i, j] = -s; R[j, i] = s; R[j, j] = c
        d = np.zeros((3, 3)); d[i, i] = -s; d[i, j] = -c; d[j, i] = c; d[j, j] = -s
        dd = np.zeros((3, 3)); dd[i, i] = -c; dd[i, j] = s; dd[j, i] = -s; dd[j, j] = -c
        return R, d, dd
    X, dX, ddX = axis(roll, 0); Y, dY, ddY = axis(pitch, 1); Z, dZ, ddZ = axis(yaw, 2)
    R = Z @ Y @ X
    D = [Z @ Y @ dX, Z @ dY @ X, dZ @ Y @ X]
    DD = [[Z @ Y @ ddX, Z @ dY @ dX, dZ @ Y @ dX], [None, Z @ ddY @ X, dZ @ dY @ X], [None, None, ddZ @ Y @ X]]
    return R, D, DD


def quat_derivatives(w, x, y, z):
    R = np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * z * w, 2 * x * z + 2 * y * w],
                  [2 * x * y + 2 * z * w, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * x * w],
                  [2 * x * z - 2 * y * w, 2 * y * z + 2 * x * w, 1 - 2 * x * x - 2 * y * y]])
    D = [np.array([[0, 2 * y, 2 * z], [2 * y, -4 * x, -2 * w], [2 * z, 2 * w, -4 * x]], float),
         np.array([[-4 * y, 2 * x, 2 * w], [2 * x, 0, 2 * z], [-2 * w, 2 * z, -4 * y]], float),
         np.array([[-4 * z, -2 * w, 2 * x], [2 * w, -4 * z, 2 * y], [2 * x, 2 * y, 0]], float)]
    E = lambda i, j: np.eye(3)[[i]].T @ np.eye(3)[[j]]
    DD = [[-4 * (E(1, 1) + E(2, 2)), 2 * (E(0, 1) + E(1, 0)), 2 * (E(0, 2) + E(2, 0))],
          [None, -4 * (E(0, 0) + E(2, 2)), 2 * (E(1, 2) + E(2, 1))],
          [None, None, -4 * (E(0, 0) + E(1, 1))]]
    return R, D, DD


def compute_uncertainty(A, B, CA, CB, T, mode="euler"):
    """A, B: n x 3; CA, CB: n x 3 x 3; T: 4 x 4 (A ~ R B + t) -> 6 x 6 (and the scaled H, stacked G for inspection)"""
    A = np.asarray(A, float); B = np.asarray(B, float); n = len(A)
    q = quaternion_from_rotation(np.asarray(T, float)[:3, :3])
    R, D, DD = (euler_derivatives if mode == "euler" else quat_derivatives)(*q)
    t = np.asarray(T, float)[:3, 3]
    H = np.zeros((6, 6)); Q = np.zeros((6, 6)); G = np.zeros((6 * n, 6))
    for i in range(n):
        a, b = A[i], B[i]
        r = a - R @ b - t
        M = np.zeros((3, 6)); M[:, :3] = np.eye(3)
        for k in range(3): M[:, 3 + k] = D[k] @ b
        S = np.zeros((6, 6))
        for k in range(3):
            for l in range(k, 3):
                S[3 + k, 3 + l] = S[3 + l, 3 + k] = r @ (DD[k][l] @ b)
        H += 2 * (M.T @ M - S)
        Ga = -2 * M
        Gb = np.zeros((3, 6)); Gb[:, :3] = 2 * R.T
        for k in range(3): Gb[:, 3 + k] = -2 * (D[k].T @ r) + 2 * (R.T @ (D[k] @ b))
        G[3 * i:3 * i + 3] = Ga; G[3 * n + 3 * i:3 * n + 3 * i + 3] = Gb
        Q += Ga.T @ CA[i] @ Ga + Gb.T @ CB[i] @ Gb
    k = 1.0 / n
    H = k * H; Q = (k * k) * Q
    Hi = np.linalg.inv(H)
    return Hi @ Q @ Hi, H, k * G
