"""runs pslam_transform_uncertainty_batch a few times (64 transforms x 100 point pairs, both parametrisations) for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from putslam_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(5)
A, B, CA, CB, T = [], [], [], [], []
for b in range(64):
    n = 100
    a = rng.uniform(-1, 1, (n, 3)) + [0, 0, 2.5]
    ang = 0.05
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    t = np.array([0.02, -0.01, 0.03])
    A.append(a); B.append(a @ R.T + t + rng.normal(0, 1e-3, (n, 3)))
    c = np.tile(np.eye(3) * 1e-4, (n, 1, 1))
    CA.append(c); CB.append(c.copy())
    M = np.eye(4); M[:3, :3] = R; M[:3, 3] = t
    T.append(M[:3, :4].copy())
for mode in ("euler", "g2o"):
    for _ in range(3):
        out = ctx.transform_uncertainty(A, B, CA, CB, T, mode=mode)
print("ok", out[0].shape, int(out[1].sum()))
