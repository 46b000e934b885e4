"""Transform uncertainty (TransformEst::computeUncertainty / computeUncertaintyG2O, reference
include/putslam/TransformEst/transformEst.h:29-272): the derived oracle and the device source run on the CPU, both against
the reference's own expressions (tests/golden/uncertainty_ref.npz, made by tests/golden/make_uncertainty_golden.py from
the reference header)."""
import numpy as np

GOLD = None


def gold():
    global GOLD
    if GOLD is None:
        import os
        GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "uncertainty_ref.npz"))
    return GOLD


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_uncertainty_oracle_matches_reference_expressions():
    from oracle import uncertainty_oracle as U
    g = gold()
    for name in g["names"]:
        for mode in ("euler", "quat"):
            u, H, G = U.compute_uncertainty(g[f"{name}_A"], g[f"{name}_B"], g[f"{name}_CA"], g[f"{name}_CB"], g[f"{name}_T"], mode)
            assert rel(H, g[f"{name}_{mode}_H"]) < 1e-12, (name, mode)       # d2J/dtheta2, scaled by 1/n
            assert rel(G, g[f"{name}_{mode}_G"]) < 1e-12, (name, mode)       # d2J/dtheta dX, all 6n rows
            assert rel(u, g[f"{name}_{mode}_U"]) < 1e-9, (name, mode)


def test_uncertainty_oracle_is_the_hessian_of_the_cost():
    """independent of the reference's expressions: H and G by central differences of the gradient of J"""
    from oracle import uncertainty_oracle as U
    rng = np.random.default_rng(3)
    n = 9
    B = rng.uniform(-1, 1, (n, 3)); rpy = np.array([0.3, -0.2, 0.5]); t = np.array([0.1, -0.2, 0.3])
    def Rof(p):
        return U.euler_derivatives_from_angles(*p)[0]
    A = B @ Rof(rpy).T + t + rng.normal(0, 0.05, (n, 3))
    def cost(theta, A_, B_):
        r = A_ - B_ @ Rof(theta[3:]).T - theta[:3]
        return (r * r).sum()
    th = np.concatenate([t, rpy])
    h = 1e-4
    H = np.zeros((6, 6))
    for i in range(6):
        for j in range(6):
            e_i = np.eye(6)[i] * h; e_j = np.eye(6)[j] * h
            H[i, j] = (cost(th + e_i + e_j, A, B) - cost(th + e_i - e_j, A, B) - cost(th - e_i + e_j, A, B) + cost(th - e_i - e_j, A, B)) / (4 * h * h)
    T = np.eye(4); T[:3, :3] = Rof(rpy); T[:3, 3] = t
    _, Hs, _ = U.compute_uncertainty(A, B, np.zeros((n, 3, 3)), np.zeros((n, 3, 3)), T, "euler")
    assert np.abs(Hs * n - H).max() < 1e-5 * np.abs(H).max()


def test_uncertainty_kernel_source_matches_reference_expressions():
    import unc_emul as E
    g = gold()
    for name in g["names"]:
        for mode in ("euler", "quat"):
            u, ok = E.compute(g[f"{name}_A"], g[f"{name}_B"], g[f"{name}_CA"], g[f"{name}_CB"], g[f"{name}_T"], mode)
            assert ok and rel(u, g[f"{name}_{mode}_U"]) < 1e-9, (name, mode)


def test_uncertainty_kernel_source_quaternion_branches():
    """rotations by more than 120 degrees take the largest-diagonal branches of the quaternion conversion"""
    import unc_emul as E
    from oracle import uncertainty_oracle as U
    rng = np.random.default_rng(4)
    for axis in range(3):
        v = np.zeros(3); v[axis] = 1.0; v += rng.normal(0, 0.05, 3); v /= np.linalg.norm(v)
        ang = 3.0
        K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        assert np.trace(R) < 0
        T = np.eye(4); T[:3, :3] = R; T[:3, 3] = rng.uniform(-1, 1, 3)
        B = rng.uniform(-1, 1, (20, 3)); A = B @ R.T + T[:3, 3] + rng.normal(0, 0.01, (20, 3))
        L = rng.normal(0, 0.01, (2, 20, 3, 3)); CA = L[0] @ L[0].transpose(0, 2, 1); CB = L[1] @ L[1].transpose(0, 2, 1)
        for mode in ("euler", "quat"):
            u, ok = E.compute(A, B, CA, CB, T, mode)
            ref, _, _ = U.compute_uncertainty(A, B, CA, CB, T, mode)
            assert ok and rel(u, ref) < 1e-9, (axis, mode)
