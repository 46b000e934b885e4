"""GPU parity of the KLT tracking seam (MatcherOpenCV::performTracking, reference src/Matcher/matcherOpenCV.cpp:209-300)
through the C ABI: pslam_klt_track == cv::calcOpticalFlowPyrLK bit for bit (golden vectors, the oracle on every output,
live cv2 at the reference's frame size), pslam_klt_perform_tracking == the oracle's restatement of the threshold and the
pairwise too-close rule."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def test_klt_track_golden_cv2(ctx, golden):
    g = golden["klt_cv2"]
    a, b, pts = g["a"], g["b"], g["pts"]
    cases = {"colour": (a, b, {}), "gray": (a[..., 0], b[..., 0], {}), "mineig": (a, b, {"min_eig_err": True}),
             "initflow": (a, b, {"init_xy": g["init"]})}
    for name in g["names"]:
        ia, ib, kw = cases[str(name)]
        r = ctx.klt_track(ia, ib, pts, **kw)
        assert np.array_equal(r["status"], g[f"{name}_status"]), name
        ok = r["status"] == 1
        assert 60 < ok.sum() < len(pts)
        assert np.array_equal(bits(r["xy"][ok]), bits(g[f"{name}_next"][ok])), name
        assert np.array_equal(bits(r["err"][ok]), bits(g[f"{name}_err"][ok])), name


def test_klt_track_vs_oracle_every_output(ctx):
    """lost points included (cv2 defines only the tracked ones): hard-edged frames (accumulation order matters), other
    windows, points outside the frame, a pyramid cut short by a small frame, clamped criteria, initial flow"""
    from oracle import klt_oracle as K
    rng = np.random.default_rng(71)
    up = lambda m: np.repeat(np.repeat(m, 2, 0), 2, 1)
    for win, lev, cn, shape in ((9, 2, 1, (40, 52)), (13, 1, 3, (40, 52)), (7, 3, 3, (38, 50)), (4, 2, 3, (30, 30)), (7, 5, 1, (9, 20)),
                                (21, 1, 3, (40, 40))):
        a = up(rng.integers(0, 2, shape + ((cn,) if cn == 3 else ()), dtype=np.uint8) * 255)
        b = np.roll(a, 1, 1); b[::3] = np.roll(b[::3], 1, 0)
        b = ((a.astype(np.int32) + b) // 2).astype(np.uint8)
        H, W = a.shape[:2]
        pts = np.stack([rng.uniform(-3, W + 3, 36), rng.uniform(-3, H + 3, 36)], 1).astype(np.float32)
        for okw, gkw in (({}, {}), ({"min_eig_err": True, "max_iter": 200, "eps": 0.0}, {"min_eig_err": True, "max_iter": 200, "eps": 0.0}),
                         ({"init": pts + np.float32(0.7)}, {"init_xy": pts + np.float32(0.7)})):
            o_n, o_s, o_e = K.lk_pyr(a, b, pts, win=win, max_level=lev, **okw)
            r = ctx.klt_track(a, b, pts, win=win, max_level=lev, **gkw)
            assert np.array_equal(r["status"], o_s), (win, cn, list(okw))
            assert np.array_equal(bits(r["xy"]), bits(o_n)), (win, cn, list(okw))
            assert np.array_equal(bits(r["err"]), bits(o_e)), (win, cn, list(okw))


def _sequence(rng, n_frames, H=480, W=640):
    import cv2
    g = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 2.0)
    g = cv2.normalize(g, None, 0, 255, cv2.NORM_MINMAX)
    frames = [g]
    for k in range(1, n_frames):
        M = np.float32([[0.999, 0.015, 3.1 + k], [-0.015, 0.999, -2.2]])
        f = cv2.warpAffine(frames[-1], M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        frames.append(np.clip(f.astype(np.int32) + rng.integers(-3, 4, f.shape), 0, 255).astype(np.uint8))
    return frames


def test_klt_track_full_frame_live_cv2_and_resident_previous_frame(ctx):
    """the reference's configuration at BASELINE's frame size: 640 x 480 x 3, 1000 points, window 7, 3 levels, 30 / 0.01,
    trackingMinEigThreshold 0; then the next frame with prev_image = NULL (pyramid kept from the last call)"""
    import cv2
    rng = np.random.default_rng(5)
    f0, f1, f2 = _sequence(rng, 3)
    pts = np.stack([rng.uniform(0, 640, 1000), rng.uniform(0, 480, 1000)], 1).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01)

    def check(prev, cur, p, r):
        p1, st, er = cv2.calcOpticalFlowPyrLK(prev, cur, p.reshape(-1, 1, 2), None, winSize=(7, 7), maxLevel=3, criteria=crit,
                                              minEigThreshold=0.0)
        st = st.ravel(); ok = st == 1
        assert np.array_equal(r["status"], st) and ok.sum() > 900
        assert np.array_equal(bits(r["xy"][ok]), bits(p1.reshape(-1, 2)[ok]))
        assert np.array_equal(bits(r["err"][ok]), bits(er.ravel()[ok]))
        return p1.reshape(-1, 2)[ok]

    l0 = ctx.launches
    r = ctx.klt_track(f0, f1, pts, min_eig_threshold=0.0)
    assert ctx.launches - l0 == 4                                  # 3 pyramid levels (both frames per launch) + tracker
    p1 = check(f0, f1, pts, r)
    l0 = ctx.launches
    r2 = ctx.klt_track(None, f2, p1, min_eig_threshold=0.0)        # f1's pyramid is resident
    assert ctx.launches - l0 == 4
    check(f1, f2, p1, r2)
    # row stride: a view into a wider buffer must be packed by the library
    wide = np.zeros((480, 700, 3), np.uint8); wide[:, :640] = f1
    wide0 = np.zeros((480, 700, 3), np.uint8); wide0[:, :640] = f0            # one row stride serves both frames
    from putslam_b200 import api
    import ctypes as C
    xy = np.zeros((1000, 2), np.float32); st = np.zeros(1000, np.uint8); err = np.zeros(1000, np.float32)
    rc = ctx.lib.pslam_klt_track(ctx.h, api._p(wide0, C.c_uint8), api._p(wide, C.c_uint8), 640, 480, 2100, 3, api._p(pts, C.c_float),
                                 api._p(xy, C.c_float), 1000, 7, 3, 3, 30, C.c_double(0.01), 0, C.c_double(0.0),
                                 api._p(st, C.c_uint8), api._p(err, C.c_float))
    assert rc == 0 and np.array_equal(st, r["status"]) and np.array_equal(bits(xy), bits(r["xy"]))


def test_klt_perform_tracking_vs_oracle(ctx, golden):
    """threshold + pairwise too-close rule + ordered survivors, on tracker output with planted coincident points, ties and
    the reference's parameters (trackingErrorThreshold 25, minimalReprojDistanceNewTrackingFeatures 3)"""
    from oracle import klt_oracle as K
    g = golden["klt_cv2"]
    a, b = g["a"], g["b"]
    rng = np.random.default_rng(9)
    pts = np.stack([rng.uniform(5, 195, 1500), rng.uniform(5, 145, 1500)], 1).astype(np.float32)
    pts[100:200] = pts[:100]                                       # identical tracks -> identical err: ties
    pts[200:300] = pts[:100] + np.float32(0.5)
    for thr, dist in ((25.0, 3.0), (4.0, 1.0), (1e9, 0.0), (2.0, 500.0)):
        r = ctx.klt_track(a, b, pts, min_eig_threshold=0.0, prune=(thr, dist))
        plain = ctx.klt_track(a, b, pts, min_eig_threshold=0.0)
        assert np.array_equal(r["status"], plain["status"]) and np.array_equal(bits(r["xy"]), bits(plain["xy"]))
        kept = K.perform_tracking(r["err"], r["status"], r["xy"], thr, dist)
        assert np.array_equal(r["kept"], kept), (thr, dist)
    assert 0 < len(ctx.klt_track(a, b, pts, min_eig_threshold=0.0, prune=(25.0, 3.0))["kept"]) < 1500


def test_klt_argument_errors(ctx):
    from putslam_b200 import api
    a = np.zeros((40, 50, 3), np.uint8)
    pts = np.zeros((3, 2), np.float32)
    fresh = api.Context(0)
    try:
        with pytest.raises(api.PslamError):                        # no resident previous frame yet
            fresh.klt_track(None, a, pts)
        fresh.klt_track(a, a, pts, max_level=1)
        with pytest.raises(api.PslamError):                        # resident pyramid is shallower than asked for
            fresh.klt_track(None, a, pts, max_level=2)
        with pytest.raises(api.PslamError):                        # another frame size
            fresh.klt_track(None, np.zeros((40, 52, 3), np.uint8), pts)
        with pytest.raises(api.PslamError):
            fresh.klt_track(a, a, pts, win=23)
        r = fresh.klt_track(a, a, np.zeros((0, 2), np.float32))
        assert len(r["xy"]) == 0
        r = fresh.klt_track(a, a, pts)                             # flat frame: nothing trackable at the default threshold
        assert not r["status"].any()
    finally:
        fresh.close()


def _klt_frame_case(rng, n=260, H=240, W=320):
    import cv2
    g = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 1.6)
    f0 = cv2.normalize(g, None, 0, 255, cv2.NORM_MINMAX)
    M = np.float32([[1, 0, 2.3], [0, 1, -1.4]])                     # camera slides parallel to a wall 2 m away
    f1 = cv2.warpAffine(f0, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    f1 = np.clip(f1.astype(np.int32) + rng.integers(-2, 3, f1.shape), 0, 255).astype(np.uint8)
    depth0 = (10000 + rng.integers(-3, 4, (H, W))).astype(np.uint16)
    depth1 = (10000 + rng.integers(-3, 4, (H, W))).astype(np.uint16)
    depth1[::37, ::11] = 0                                          # holes: invalid depth is filtered inside RANSAC
    pts = np.stack([rng.uniform(6, W - 7, n), rng.uniform(6, H - 7, n)], 1).astype(np.float32)
    pts[200:230] = pts[:30] + np.float32(0.6)                       # too close to others
    return f0, f1, depth0, depth1, pts


@pytest.mark.parametrize("undistort", [False, True])
def test_klt_frame_vs_oracle_chain(ctx, O, undistort):
    """pslam_klt_frame == performTracking -> removeImageDistortion -> keypoints2Dto3D -> RANSAC::estimateTransformation
    composed from the oracle's restatements of the reference functions (Matcher::trackKLT, matcher.cpp:151-207)"""
    from oracle import klt_oracle as K
    from putslam_b200 import synth
    rng = np.random.default_rng(31 + int(undistort))
    f0, f1, depth0, depth1, pts = _klt_frame_case(rng)
    fx, fy, cx, cy = synth.FX / 2, synth.FY / 2, synth.CX / 2, synth.CY / 2
    from putslam_b200 import api
    cam = api.make_camera(fx, fy, cx, cy)
    prm = api.default_ransac_params()
    prm.fx, prm.fy, prm.cx, prm.cy = fx, fy, cx, cy
    und0 = O.undistort(pts, fx, fy, cx, cy, synth.DIST) if undistort else pts
    prev_xyz, _ = O.backproject(und0, depth0, fx, fy, cx, cy, 5000.0)
    prev_xyz[::3] += rng.normal(0, 0.3, prev_xyz[::3].shape).astype(np.float32)   # a third of the 3-D points are outliers
    r = ctx.klt_frame(f0, f1, pts, prev_xyz, depth1, cam=cam, undistort=undistort, params=prm, seed=5)
    # the oracle chain
    nxt, st, err = K.lk_pyr(f0, f1, pts, min_eig_thr=0.0)
    kept = K.perform_tracking(err, st, nxt, 25.0, 3.0)
    assert 150 < len(kept) < len(pts)
    assert np.array_equal(r["status"], st) and np.array_equal(bits(r["xy"]), bits(nxt)) and np.array_equal(bits(r["err"]), bits(err))
    assert np.array_equal(r["kept"], kept) and r["n_matches"] == len(kept)
    und1 = O.undistort(nxt[kept], fx, fy, cx, cy, synth.DIST) if undistort else nxt[kept]
    xyz1, dd = O.backproject(und1, depth1, fx, fy, cx, cy, 5000.0)
    assert np.array_equal(bits(r["uv_undist"]), bits(np.ascontiguousarray(und1, np.float32)))
    assert np.array_equal(bits(r["xyz"]), bits(xyz1)) and np.array_equal(bits(r["det_dist"]), bits(dd))
    ref = O.ransac(prev_xyz, xyz1, kept.astype(np.int32), np.arange(len(kept), dtype=np.int32), params=None, seed=5)
    assert np.array_equal(r["inliers"], ref["inliers"]) and 100 < len(ref["inliers"]) < 0.8 * len(kept)
    assert np.abs(r["T"] - ref["T"]).max() <= 1e-5                  # 1e-5 m / 1e-5 rad (north_star)
    assert r["hyp_used"] == ref["hyp_used"] and r["best_ratio"] == ref["best_ratio"]
    assert r["inlier_ratio"] == len(ref["inliers"]) / len(kept)
    # the separate calls give the same survivors
    sep = ctx.klt_track(f0, f1, pts, min_eig_threshold=0.0, prune=(25.0, 3.0))
    assert np.array_equal(sep["kept"], r["kept"])


def test_klt_frame_degenerate(ctx, O):
    """no features; and a flat frame where nothing is tracked: identity, no inliers, ratio 0 (trackKLT's own values)"""
    a = np.full((60, 80, 3), 90, np.uint8)
    depth = np.full((60, 80), 10000, np.uint16)
    r = ctx.klt_frame(a, a, np.zeros((0, 2), np.float32), np.zeros((0, 3), np.float32), depth)
    assert len(r["kept"]) == 0 and np.array_equal(r["T"], np.eye(4, dtype=np.float32)) and r["inlier_ratio"] == 0.0
    pts = np.array([[10, 10], [40, 30], [70, 50]], np.float32)
    r = ctx.klt_frame(a, a, pts, np.ones((3, 3), np.float32), depth, min_eig_threshold=1e-4)
    assert not r["status"].any() and len(r["kept"]) == 0 and len(r["inliers"]) == 0
    assert np.array_equal(r["T"], np.eye(4, dtype=np.float32)) and r["inlier_ratio"] == 0.0
