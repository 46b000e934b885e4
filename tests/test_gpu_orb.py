"""GPU parity tests (-m gpu) of the ORB descriptor path (pslam_orb_describe == cv::ORB::compute with provided keypoints,
the reference's MatcherOpenCV::describeFeatures, src/Matcher/matcherOpenCV.cpp:181-195): surviving keypoints, their
order and every descriptor bit against the cv2 golden vectors, the oracle and live cv2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_orb_describe_golden_cv2(ctx, golden):
    g = golden["orb_cv2"]
    for name in g["names"]:
        order, desc = ctx.orb_describe(g[f"{name}_img"], g[f"{name}_xy"], g[f"{name}_octave"], g[f"{name}_angle"])
        assert np.array_equal(order, g[f"{name}_order"]), name
        assert np.array_equal(desc, g[f"{name}_desc"]), name


def scene(rng, H, W):
    import cv2
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W), dtype=np.uint8), (0, 0), 2.0)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)


@pytest.mark.parametrize("W,H,n,max_oct", [(640, 480, 1000, 7), (641, 479, 500, 7), (320, 240, 300, 5), (1280, 720, 800, 11),
                                            (200, 150, 100, 7)])
def test_orb_describe_vs_oracle_and_live_cv2(ctx, W, H, n, max_oct):
    import cv2
    from oracle import orb_oracle as OO
    rng = np.random.default_rng(W * 31 + H)
    img = scene(rng, H, W)
    xy = np.stack([rng.uniform(5, W - 5, n), rng.uniform(5, H - 5, n)], 1).astype(np.float32)
    octave = rng.integers(0, max_oct + 1, n).astype(np.int32)
    angle = rng.uniform(0, 360, n).astype(np.float32)
    order, desc = ctx.orb_describe(img, xy, octave, angle)
    eo, ed = OO.describe(img, xy, octave, angle)
    assert np.array_equal(order, eo) and 0 < order.size < n
    assert np.array_equal(desc, ed)
    if max_oct <= 7:   # cv::ORB::compute itself (levels beyond nlevels are handled by the same code in OpenCV too)
        kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, int(o)) for (x, y), a, o in zip(xy, angle, octave)]
        k2, d2 = cv2.ORB_create().compute(img, kps)
        assert len(k2) == order.size and np.array_equal(desc, d2)
        assert all(float(k.pt[0]) == float(xy[i, 0]) and float(k.pt[1]) == float(xy[i, 1]) for k, i in zip(k2, order))
    # a second call with another level count on the same context (tables are rebuilt), then the first again
    o2, d2b = ctx.orb_describe(img, xy[octave < 3], octave[octave < 3], angle[octave < 3])
    e2, ed2 = OO.describe(img, xy[octave < 3], octave[octave < 3], angle[octave < 3])
    assert np.array_equal(o2, e2) and np.array_equal(d2b, ed2)
    o3, d3 = ctx.orb_describe(img, xy, octave, angle)
    assert np.array_equal(o3, order) and np.array_equal(d3, desc)


def test_orb_describe_colour_stride_and_edges(ctx):
    import cv2
    from oracle import orb_oracle as OO
    from putslam_b200 import api
    rng = np.random.default_rng(9)
    g = scene(rng, 240, 320)
    bgr = np.stack([g, np.roll(g, 4, 1), 255 - g], 2).copy()
    n = 200
    xy = np.stack([rng.uniform(25, 295, n), rng.uniform(25, 215, n)], 1).astype(np.float32)
    xy[:4] = [[30.5, 100.0], [31.5, 100.0], [288.5, 80.0], [289.5, 80.0]]
    octave = rng.integers(0, 8, n).astype(np.int32); angle = rng.uniform(0, 360, n).astype(np.float32)
    order, desc = ctx.orb_describe(bgr, xy, octave, angle)
    eo, ed = OO.describe(bgr, xy, octave, angle)
    assert np.array_equal(order, eo) and np.array_equal(desc, ed)
    kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, int(o)) for (x, y), a, o in zip(xy, angle, octave)]
    _, d2 = cv2.ORB_create().compute(bgr, kps)
    assert np.array_equal(desc, d2)
    # every keypoint outside the border band -> nothing left; no keypoints -> nothing to do; bad octave -> error
    o, d = ctx.orb_describe(g, np.array([[5.0, 5.0], [318.0, 200.0]], np.float32), [0, 1], [0.0, 10.0])
    assert o.size == 0 and d.shape == (0, 32)
    o, d = ctx.orb_describe(g, np.zeros((0, 2), np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32))
    assert o.size == 0
    with pytest.raises(api.PslamError):
        ctx.orb_describe(g, np.array([[100.0, 100.0]], np.float32), [12], [0.0])
    with pytest.raises(api.PslamError):
        ctx.orb_describe(g[:60, :80].copy(), np.array([[40.0, 30.0]], np.float32), [7], [0.0])      # level 7 would be 22 x 17


# ---------------------------------------------------------------- detection
def _rows(d):
    return np.column_stack([d["xy"], d["size"], d["angle"], d["response"]]).astype(np.float32), d["octave"]


def test_orb_detect_golden_cv2(ctx, golden):
    """cv::ORB::detect recorded from cv2: same keypoints, same values, SAME ORDER (the host half runs OpenCV's
    retainBest with the same libstdc++ algorithms)"""
    g = golden["orb_detect_cv2"]
    for name in g["names"]:
        kp, octave = _rows(ctx.orb_detect(g[f"{name}_img"], int(g[f"{name}_nfeatures"])))
        assert np.array_equal(octave, g[f"{name}_octave"]), name
        assert np.array_equal(kp.view(np.uint32), g[f"{name}_kp"].view(np.uint32)), name


@pytest.mark.parametrize("W,H,nf,sigma", [(640, 480, 500, 1.6), (640, 480, 2000, 1.2), (641, 479, 300, 2.0), (1280, 720, 1000, 1.5),
                                           (512, 512, 50, 1.0)])
def test_orb_detect_vs_live_cv2_and_oracle(ctx, W, H, nf, sigma):
    import cv2
    from oracle import orb_oracle as OO
    rng = np.random.default_rng(W + 7 * H + nf)
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W), dtype=np.uint8), (0, 0), sigma)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    out = ctx.orb_detect(img, nf)
    ref = cv2.ORB_create(nfeatures=nf).detect(img)
    assert len(ref) == out["octave"].size and len(ref) > 0.5 * nf
    rk = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in ref], np.float32)
    kp, octave = _rows(out)
    assert np.array_equal(octave, [k.octave for k in ref])
    assert np.array_equal(kp.view(np.uint32), rk.view(np.uint32))            # order included
    if W * H <= 640 * 480:
        mine = OO.detect(img, nf)
        got = {(float(m[0]), float(m[1]), int(m[5])): (float(m[2]), float(m[3]), float(m[4])) for m in mine}
        dev = {(float(r[0]), float(r[1]), int(o)): (float(r[2]), float(r[3]), float(r[4])) for r, o in zip(kp, octave)}
        assert got == dev


def test_orb_detect_candidates_vs_oracle(ctx):
    """nfeatures large enough that nothing is culled: every FAST corner that survives suppression and the border filter
    comes back, with the oracle's Harris response and angle"""
    import cv2
    from oracle import orb_oracle as OO
    rng = np.random.default_rng(12)
    img = cv2.GaussianBlur(rng.integers(0, 256, (300, 400), dtype=np.uint8), (0, 0), 1.8)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    out = ctx.orb_detect(img, 200000, cap=300000)
    cands = OO.detect_candidates(img)
    assert len(cands) == out["octave"].size and len(cands) > 300
    got = {(float(x), float(y), int(o)): (float(a), float(r)) for (x, y), a, r, o in zip(out["xy"], out["angle"], out["response"], out["octave"])}
    for (l, x, y, s, hr, ang) in cands:
        sc = OO.level_scale(l)
        key = (float(np.float32(x) * sc), float(np.float32(y) * sc), l)
        assert got[key] == (float(ang), float(hr)), (l, x, y)


def test_orb_detect_colour_and_edges(ctx):
    import cv2
    from putslam_b200 import api
    rng = np.random.default_rng(3)
    g = cv2.GaussianBlur(rng.integers(0, 256, (480, 640), dtype=np.uint8), (0, 0), 1.5)
    rgb = np.stack([g, np.roll(g, 2, 1), 255 - np.roll(g, 3, 0)], 2).copy()
    for order, code in ((0, cv2.COLOR_BGR2GRAY), (1, cv2.COLOR_RGB2GRAY)):
        out = ctx.orb_detect(rgb, 400, colour_order=order)
        ref = cv2.ORB_create(nfeatures=400).detect(cv2.cvtColor(rgb, code))
        rk = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in ref], np.float32)
        assert np.array_equal(_rows(out)[0].view(np.uint32), rk.view(np.uint32)) and len(ref) > 100
    flat = np.full((480, 640), 128, np.uint8)
    assert ctx.orb_detect(flat, 500)["octave"].size == 0                      # no corners at all
    assert ctx.orb_detect(g, 0)["octave"].size == 0                           # zero budget
    small = ctx.orb_detect(g[:100, :120].copy(), 100)                         # upper levels narrower than the border band
    ref = cv2.ORB_create(nfeatures=100).detect(g[:100, :120].copy())
    assert small["octave"].size == len(ref) and (small["octave"] <= 2).all()
    with pytest.raises(api.PslamError):
        ctx.orb_detect(g, 500, cap=10)                                        # more keypoints than the caller's buffers hold


def test_orb_describe_on_the_resident_frame(ctx):
    """image == NULL: describe on the frame the preceding detect uploaded (gray and colour; the colour frame is detected
    with RGB2GRAY and described with BGR2GRAY, like the reference's two calls), same answer as passing the image again"""
    import cv2
    from putslam_b200 import api
    rng = np.random.default_rng(21)
    g = cv2.GaussianBlur(rng.integers(0, 256, (480, 640), dtype=np.uint8), (0, 0), 1.5)
    rgb = np.stack([g, np.roll(g, 3, 1), 255 - np.roll(g, 2, 0)], 2).copy()
    for img, order_flag in ((g, 0), (rgb, 1)):
        det = ctx.orb_detect(img, 500, colour_order=order_flag)
        again = ctx.orb_describe(img, det["xy"], det["octave"], det["angle"])
        det = ctx.orb_detect(img, 500, colour_order=order_flag)                 # fresh upload, then describe without image
        res = ctx.orb_describe(None, det["xy"], det["octave"], det["angle"], resident_shape=img.shape)
        assert np.array_equal(res[0], again[0]) and np.array_equal(res[1], again[1]) and res[0].size > 200
        twice = ctx.orb_describe(None, det["xy"], det["octave"], det["angle"], resident_shape=img.shape)
        assert np.array_equal(twice[1], again[1])
        kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, int(o)) for (x, y), a, o in zip(det["xy"], det["angle"], det["octave"])]
        _, d2 = cv2.ORB_create().compute(img, kps)
        assert np.array_equal(res[1], d2)
    with pytest.raises(api.PslamError):                                          # resident frame has another shape
        ctx.orb_describe(None, det["xy"], det["octave"], det["angle"], resident_shape=(240, 320))
    fresh = api.Context(0)
    try:
        with pytest.raises(api.PslamError):                                      # nothing uploaded yet on this context
            fresh.orb_describe(None, det["xy"], det["octave"], det["angle"], resident_shape=rgb.shape)
    finally:
        fresh.close()


@pytest.mark.parametrize("W,H,threshold,sigma", [(640, 480, 10, 1.5), (640, 480, 20, 1.0), (333, 217, 10, 2.0), (1280, 720, 40, 1.2), (64, 48, 5, 0.8)])
def test_fast_detect_vs_live_cv2(ctx, W, H, threshold, sigma):
    """detector option "FAST": cv::FastFeatureDetector (9_16, non-max suppression): positions, raster order, scores"""
    import cv2
    from oracle import orb_oracle as OO
    rng = np.random.default_rng(W + H + threshold)
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W), dtype=np.uint8), (0, 0), sigma)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    xy, resp = ctx.fast_detect(img, threshold)
    ref = cv2.FastFeatureDetector_create(threshold, True).detect(img)
    assert len(ref) == resp.size and len(ref) > 10
    assert np.array_equal(xy, np.array([k.pt for k in ref], np.float32))
    assert np.array_equal(resp, np.array([k.response for k in ref], np.float32))
    if W * H <= 640 * 480:
        mine = OO.fast_nms(OO.fast_score_map(img, threshold))
        assert [(int(x), int(y), int(r)) for (x, y), r in zip(xy, resp)] == mine
    rgb = np.stack([img, np.roll(img, 1, 0), np.roll(img, 2, 1)], 2).copy()
    xy3, r3 = ctx.fast_detect(rgb, threshold, colour_order=1)
    ref3 = cv2.FastFeatureDetector_create(threshold, True).detect(cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
    assert np.array_equal(xy3, np.array([k.pt for k in ref3], np.float32).reshape(-1, 2))
    assert np.array_equal(r3, np.array([k.response for k in ref3], np.float32))


def test_pixels_to_matches_pipeline(ctx):
    """The front end from pixels: detectFeatures -> describeFeatures on two frames of one scene (the second shifted by a
    known amount), then performMatching.  Every stage equals OpenCV's output, and the matches recover the shift."""
    import cv2
    import bench
    a = bench.orb_bench_image(np.random.default_rng(5))
    dx, dy = 9, -6
    b = np.roll(np.roll(a, dy, 0), dx, 1)
    b = np.clip(b.astype(np.int32) + np.random.default_rng(6).integers(-3, 4, b.shape), 0, 255).astype(np.uint8)   # sensor noise
    frames = []
    for img in (a, b):
        det = ctx.orb_detect(img, 500)
        order, desc = ctx.orb_describe(None, det["xy"], det["octave"], det["angle"], resident_shape=img.shape)
        ref_kp = cv2.ORB_create().detect(img)
        ref_kp, ref_desc = cv2.ORB_create().compute(img, ref_kp)
        assert np.array_equal(desc, ref_desc) and len(ref_kp) == order.size
        assert all(float(k.pt[0]) == float(det["xy"][i, 0]) and float(k.pt[1]) == float(det["xy"][i, 1]) for k, i in zip(ref_kp, order))
        frames.append((det["xy"][order], desc))
    (xa, da), (xb, db) = frames
    mq, mt, md = ctx.match_bf_mutual(da, db)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(da, db)
    assert np.array_equal(mq, [m.queryIdx for m in ref]) and np.array_equal(mt, [m.trainIdx for m in ref])
    assert np.array_equal(md, np.array([m.distance for m in ref], np.float32))
    shift = xb[mt] - xa[mq]
    good = (np.abs(shift[:, 0] - dx) < 2.5) & (np.abs(shift[:, 1] - dy) < 2.5)
    assert mq.size > 150 and good.mean() > 0.6, (mq.size, good.mean())


@pytest.mark.parametrize("W,H", [(100, 100), (88, 72), (80, 60)])
def test_orb_and_fast_detect_small_image_first_call_on_a_fresh_context(W, H):
    """A small image as the FIRST call of a context: the candidate buffers hold fewer records than the fixed first
    read-back chunk, which must be clamped (a detectFeatures grid cell such as 80x60 takes this path)."""
    import cv2
    from putslam_b200 import api
    rng = np.random.default_rng(W * H)
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W), dtype=np.uint8), (0, 0), 1.2)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    for which in ("orb", "fast"):
        fresh = api.Context(0)
        try:
            if which == "orb":
                out = fresh.orb_detect(img, 200)
                ref = cv2.ORB_create(nfeatures=200).detect(img)
                kp, octave = _rows(out)
                rk = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in ref], np.float32).reshape(-1, 5)
                assert np.array_equal(octave, np.array([k.octave for k in ref], octave.dtype))
                assert np.array_equal(kp.view(np.uint32), rk.view(np.uint32))
            else:
                xy, resp = fresh.fast_detect(img, 10)
                ref = cv2.FastFeatureDetector_create(10, True).detect(img)
                assert np.array_equal(xy, np.array([k.pt for k in ref], np.float32).reshape(-1, 2))
                assert np.array_equal(resp, np.array([k.response for k in ref], np.float32))
        finally:
            fresh.close()
