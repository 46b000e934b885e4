"""GPU test (-m gpu): the drop-in, end to end.  oracle/_ref/libref_tree.so holds the REFERENCE's own Matcher (matcher.cpp,
dbscan.cpp, RGBD.cpp, RANSAC.cpp ... compiled from /root/reference) with adapter/putslam_tree's `MatcherB200 : public
MatcherOpenCV` plugged into its virtual interface.  The reference's orchestration -- detectInitFeatures, runVO -> match, DBScan,
removeImageDistortion, keypoints2Dto3D, RANSAC, pointInlierRatio -- runs as compiled from its sources and reaches the GPU only
through detectFeatures / describeFeatures / performMatching, exactly as a PUTSLAM build with Matcher/matcherB200.h would.
The outcome must equal the same pipeline put together from OpenCV (cv2) and the CPU oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import bits

from oracle import ref_build as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.tree_available(), reason="oracle/_ref/libref_tree.so not present")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_dbscan(xy, eps):
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_dbscan.so"))
    xy = np.ascontiguousarray(xy, np.float32)
    kept = np.empty(max(1, len(xy)), np.int32)
    m = lib.orc_ref_dbscan(xy.ctypes.data_as(C.POINTER(C.c_float)), len(xy), C.c_double(eps), 2, 1, kept.ctypes.data_as(C.POINTER(C.c_int)))
    return kept[:m].copy()


def expected_features(cv2, gray, eps):
    """MatcherOpenCV::detectFeatures (grid 1 x 1: ORB detect, sort by response, cut to 500) -> DBScan -> describeFeatures"""
    kps = cv2.ORB_create().detect(gray)
    resp = np.array([k.response for k in kps])
    assert len(np.unique(resp)) == len(resp)            # no ties: every sort by response gives the same order
    kps = [kps[i] for i in np.argsort(-resp, kind="stable")][:500]
    keep = ref_dbscan(np.array([k.pt for k in kps], np.float32), eps)
    kps = [kps[i] for i in keep]
    kps, desc = cv2.ORB_create().compute(gray, kps)
    return np.array([[k.pt[0], k.pt[1], k.octave] for k in kps], np.float32), desc


@pytest.mark.parametrize("error_version,eps", [(0, 10.0), (2, 6.0)])
def test_reference_matcher_runs_vo_through_the_b200_virtuals(O, error_version, eps):
    import cv2
    import bench
    from putslam_b200 import synth
    rng = np.random.default_rng(5 + error_version)
    a = bench.orb_bench_image(rng)
    dx, dy = 7, -4
    b = np.roll(np.roll(a, dy, 0), dx, 1)
    b = np.clip(b.astype(np.int32) + rng.integers(-2, 3, b.shape), 0, 255).astype(np.uint8)
    rgb0 = np.stack([a, a, a], 2).copy(); rgb1 = np.stack([b, b, b], 2).copy()      # equal channels: RGB2GRAY is the identity
    depth = np.full((480, 640), 10000, np.uint16)                                    # a wall 2 m away: the shift is a rigid motion
    args = R.tree_args(vo_tracking=0, error_version=error_version, dbscan_eps=eps)
    out = R.tree_run_vo(rgb0, depth, rgb1, depth, args=args, seed=11)
    kp0, d0 = expected_features(cv2, a, eps)
    kp1, d1 = expected_features(cv2, b, eps)
    assert np.array_equal(out["kp0"], kp0) and np.array_equal(out["kp1"], kp1) and len(kp0) > 80
    cam = (synth.FX, synth.FY, synth.CX, synth.CY)
    zero = (0, 0, 0, 0, 0)
    x0, _ = O.backproject(O.undistort(kp0[:, :2], *cam, zero), depth, *cam, 5000.0)
    x1, _ = O.backproject(O.undistort(kp1[:, :2], *cam, zero), depth, *cam, 5000.0)
    assert np.array_equal(bits(out["xyz1"]), bits(x1))
    mq, mt, md = O.bf_mutual(d0, d1)
    ref = O.ransac(x0, x1, mq, mt, params=O.default_ransac_params(error_version), seed=11)
    assert np.array_equal(out["inliers"], np.stack([mq[ref["inliers"]], mt[ref["inliers"]]], 1)) and len(ref["inliers"]) > 30
    assert out["hyp_used"] == ref["hyp_used"] and np.array_equal(bits(out["T"]), bits(ref["T"]))
    assert out["ratio"] == O.point_inlier_ratio(mt[ref["inliers"]], mt, len(kp1))
    # the estimated motion is the planted shift: (dx, dy) px at 2 m
    t = out["T"][:3, 3]
    assert abs(t[0] + dx * 2.0 / synth.FX) < 0.01 and abs(t[1] + dy * 2.0 / synth.FY) < 0.01
