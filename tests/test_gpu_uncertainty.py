"""GPU parity of pslam_transform_uncertainty_batch (TransformEst::computeUncertainty / computeUncertaintyG2O, reference
include/putslam/TransformEst/transformEst.h:29-272) through the C ABI: against the reference's own expressions
(tests/golden/uncertainty_ref.npz) and the oracle.  Double precision; tolerance 1e-9 relative to the largest entry (the
reference sums sequentially, the kernel in a fixed tree order, and the 6 x 6 inverse is not Eigen's)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_uncertainty_golden_reference_expressions(ctx):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "uncertainty_ref.npz"))
    names = [str(n) for n in g["names"]]
    args = [[g[f"{n}_{k}"] for n in names] for k in ("A", "B", "CA", "CB", "T")]
    for mode in ("euler", "quat"):
        l0 = ctx.launches
        U, ok = ctx.transform_uncertainty(*args, mode=mode)          # the five problems as one batch
        assert ctx.launches - l0 == 1 and ok.all()
        for i, n in enumerate(names):
            assert rel(U[i], g[f"{n}_{mode}_U"]) < 1e-9, (n, mode)


def test_uncertainty_batch_vs_oracle(ctx):
    """ragged batch with an empty problem, rotations beyond 120 degrees (every branch of the quaternion conversion),
    sets larger than the CTA (several points per thread)"""
    from oracle import uncertainty_oracle as Uo
    rng = np.random.default_rng(12)
    A, B, CA, CB, T = [], [], [], [], []
    for n, ang, axis in ((30, 0.4, 0), (0, 0.0, 0), (700, 3.0, 0), (45, 3.0, 1), (129, 3.0, 2), (6, -1.0, 1)):
        v = np.zeros(3); v[axis] = 1.0; v += rng.normal(0, 0.05, 3); v /= np.linalg.norm(v)
        K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        t = np.eye(4); t[:3, :3] = R; t[:3, 3] = rng.uniform(-1, 1, 3)
        b = rng.uniform(-2, 2, (n, 3)); a = b @ R.T + t[:3, 3] + rng.normal(0, 0.02, (n, 3))
        L = rng.normal(0, 0.01, (2, n, 3, 3))
        A.append(a); B.append(b); CA.append(L[0] @ L[0].transpose(0, 2, 1)); CB.append(L[1] @ L[1].transpose(0, 2, 1)); T.append(t)
    for mode in ("euler", "quat"):
        U, ok = ctx.transform_uncertainty(A, B, CA, CB, T, mode=mode)
        assert ok.tolist() == [1, 0, 1, 1, 1, 1] and not U[1].any()
        for i in (0, 2, 3, 4, 5):
            ref, _, _ = Uo.compute_uncertainty(A[i], B[i], CA[i], CB[i], T[i], mode)
            assert rel(U[i], ref) < 1e-9, (i, mode)
            assert np.abs(U[i] - U[i].T).max() <= 1e-12 * np.abs(U[i]).max()
    U, ok = ctx.transform_uncertainty([], [], [], [], [])
    assert U.shape == (0, 6, 6)
