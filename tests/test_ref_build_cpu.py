"""CPU tests (-m "not gpu"): the oracle against the REFERENCE'S OWN COMPILED CODE.

oracle/_ref/libref_frontend.so is the reference's RANSAC.cpp, RGBD.cpp, kabschEst.cpp, depthSensorModel.cpp, matcher.cpp and
dbscan.cpp compiled from /root/reference where they lie (`make -C oracle ref`) against the Eigen / OpenCV stand-ins of
oracle/ref_shim.  Everything the reference's scalar code decides -- the RANSAC driver (filter, adaptive bound, strict >,
refit + Euclidean recount, rejection), the guided gate of matchXYZ, back-projection, projection, the sensor model's
covariances, Kabsch, the list edits of trackKLT -- is therefore checked here against the reference itself; what the
stand-ins decide (the arithmetic order inside Eigen's own operations) is bounded by tools/umeyama_sensitivity.py ->
profiles/umeyama_sensitivity.json.  The sample stream is replayed through rand() (see ref_frontend_wrap.cpp)."""
import numpy as np
import pytest

from conftest import bits

from oracle import ref_build as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_frontend.so not built (needs /root/reference)")

CAM = None


def cam():
    from putslam_b200 import synth
    return (synth.FX, synth.FY, synth.CX, synth.CY)


def ransac_case(case, rng):
    """one seeded RANSAC problem: size, inlier fraction, noise and error version vary; a few invalid points"""
    from putslam_b200 import synth
    m = int(rng.choice([16, 20, 40, 100, 300, 600, 800]))      # <= 800: beyond, 1 inlier makes int(log/log) overflow (UB)
    frac = float(rng.choice([0.15, 0.25, 0.4, 0.6, 0.8]))
    ev = [0, 1, 2, 4][case % 4]
    mc = synth.matched_clouds(m=m, inlier_frac=frac, seed=1000 + case, sigma=float(rng.choice([0.002, 0.01, 0.02])))
    prev, cur = mc["prev"].copy(), mc["cur"].copy()
    k = rng.integers(0, m, 4)
    prev[k[0], 2] = 7.0; cur[k[1], 0] = np.nan; prev[k[2], 2] = 0.05; cur[k[3], 2] = 6.5
    return prev, cur, mc["mq"], mc["mt"], ev


def test_ransac_driver_equals_reference_build(O):
    """RANSAC::estimateTransformation (RANSAC.cpp:50-174) compiled from the reference == orc_ransac: final inlier index
    sets, the number of hypotheses drawn, bestInlierRatio as the reference prints it, the pose bit for bit."""
    rng = np.random.default_rng(0)
    n_id = 0
    for case in range(240):
        prev, cur, mq, mt, ev = ransac_case(case, rng)
        p = O.default_ransac_params(ev)
        if case % 7 == 3:
            p.minimal_inlier_ratio_threshold = 0.5                       # some runs end in the rejection branch (:161-164)
        o = O.ransac(prev, cur, mq, mt, params=p, seed=case)
        r = R.ransac(prev, cur, mq, mt, args=R.from_oracle_params(p), seed=case)
        assert np.array_equal(o["inliers"], r["inliers"]), case
        assert o["hyp_used"] == r["hyp_used"], case
        assert np.array_equal(bits(o["T"]), bits(r["T"])), case
        if r["best_ratio_pct"] >= 0:
            assert abs(o["best_ratio"] * 100 - r["best_ratio_pct"]) <= 1e-4 * max(1.0, r["best_ratio_pct"]), case
        n_id += int(len(o["inliers"]) == 0)
    assert 5 < n_id < 200            # both outcomes occur


def test_ransac_identity_and_empty_conventions_equal_reference_build(O):
    from putslam_b200 import synth
    mc = synth.matched_clouds(m=60, inlier_frac=0.7, seed=5)
    ident = np.eye(4, dtype=np.float32)
    # fewer valid matches than minimalNumberOfMatches (:77-80): identity, no inliers, no hypothesis drawn
    p = O.default_ransac_params(0); p.minimal_number_of_matches = 61
    o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=p, seed=1)
    r = R.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], args=R.from_oracle_params(p), seed=1)
    assert r["filtered"] == -1 and r["hyp_used"] == 0 == o["hyp_used"] and r["inliers"].size == 0 == o["inliers"].size
    assert np.array_equal(r["T"], ident) and np.array_equal(o["T"], ident)
    # all points invalid by depth: the filter leaves nothing
    far = mc["prev"].copy(); far[:, 2] = 9.0
    p = O.default_ransac_params(0)
    o = O.ransac(far, mc["cur"], mc["mq"], mc["mt"], params=p, seed=1)
    r = R.ransac(far, mc["cur"], mc["mq"], mc["mt"], args=R.from_oracle_params(p), seed=1)
    assert r["inliers"].size == 0 == o["inliers"].size and np.array_equal(r["T"], ident) and np.array_equal(o["T"], ident)
    # pure outliers: the best ratio stays below minimalInlierRatioThreshold -> identity, cleared inliers (:161-164)
    rng = np.random.default_rng(3)
    prev = rng.uniform(-1, 1, (200, 3)).astype(np.float32) + np.float32([0, 0, 3]); cur = rng.uniform(-1, 1, (200, 3)).astype(np.float32) + np.float32([0, 0, 3])
    idx = np.arange(200, dtype=np.int32)
    o = O.ransac(prev, cur, idx, idx, params=p, seed=2)
    r = R.ransac(prev, cur, idx, idx, args=R.from_oracle_params(p), seed=2)
    assert r["inliers"].size == 0 == o["inliers"].size and np.array_equal(r["T"], ident) and np.array_equal(o["T"], ident)
    assert r["hyp_used"] == o["hyp_used"] > 100
    # empty match list
    e = np.zeros(0, np.int32)
    r = R.ransac(prev, cur, e, e, args=R.from_oracle_params(p), seed=2)
    o = O.ransac(prev, cur, e, e, params=p, seed=2)
    assert r["inliers"].size == 0 == o["inliers"].size and np.array_equal(r["T"], ident) and np.array_equal(o["T"], ident)


def test_ransac_iteration_bound_equals_reference_build(O):
    """RANSAC::computeRANSACIteration / saveBetterModel's bound update, observed through the number of draws: with one
    exact model and a planted inlier ratio w the loop stops after min(iters(minRatio), iters(w)) hypotheses at the latest"""
    from putslam_b200 import synth
    for frac in (0.3, 0.5, 0.7, 0.9, 1.0):
        mc = synth.matched_clouds(m=200, inlier_frac=frac, seed=int(frac * 100), sigma=0.0)
        p = O.default_ransac_params(0)
        o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=p, seed=4)
        r = R.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], args=R.from_oracle_params(p), seed=4)
        assert o["hyp_used"] == r["hyp_used"] and np.array_equal(o["inliers"], r["inliers"])
        assert r["hyp_used"] <= max(1, O.ransac_iterations(o["best_ratio"])) + 487 * (o["best_ratio"] == 0)


def test_point_inlier_ratio_equals_reference_build(O):
    rng = np.random.default_rng(1)
    for _ in range(50):
        all_t = rng.integers(0, 40, 60).astype(np.int32)
        inl = all_t[rng.random(60) < 0.4]
        if np.unique(all_t).size:
            assert R.point_inlier_ratio(inl, all_t) == O.point_inlier_ratio(inl, all_t, 40)


def test_backprojection_equals_reference_build(O, golden):
    """RGBD::removeImageDistortion + keypoints2Dto3D + roundSize + the detDist expression, compiled from RGBD.cpp"""
    from putslam_b200 import synth
    for seed in range(6):
        fp = synth.frame_pair(n=400, seed=seed, distorted=bool(seed % 2))
        for uv, depth in ((fp["uv1"], fp["depth1"]), (fp["uv2"], fp["depth2"])):
            # away from the right / bottom border: roundSize returns `size` there and the reference reads out of bounds
            keep = (uv[:, 0] < 638.4) & (uv[:, 1] < 478.4)
            uv = uv[keep]
            xyz, dd = O.backproject(uv, depth, *cam(), synth.DEPTH_SCALE)
            used, rxyz, rdd = R.backproject(uv, depth)
            assert np.array_equal(bits(xyz), bits(rxyz)) and np.array_equal(bits(dd), bits(rdd))
            und = O.undistort(uv, *cam(), synth.DIST)
            keep2 = (und[:, 0] < 638.4) & (und[:, 1] < 478.4)
            xyz_u, dd_u = O.backproject(und[keep2], depth, *cam(), synth.DEPTH_SCALE)
            used, rxyz, rdd = R.backproject(uv[keep2], depth, dist5=synth.DIST)
            assert np.array_equal(bits(used), bits(und[keep2]))
            assert np.array_equal(bits(xyz_u), bits(rxyz)) and np.array_equal(bits(dd_u), bits(rdd))
    # the stand-in's undistortPoints is the algorithm pinned against the real OpenCV
    g = golden["undistort_cv2"]
    K = g["K"]
    depth = np.zeros((480, 640), np.uint16)
    used, _, _ = R.backproject(g["uv"], depth, k4=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]), dist5=g["dist"])
    assert np.array_equal(bits(used), bits(g["uv_undist"]))
    # roundSize (RGBD.cpp:10-16): negative -> 0, in range -> round half away from zero, beyond size-1 -> size (the quirk)
    assert [R.round_size(x, 640) for x in (-3.2, 0.49, 0.5, 1.5, 638.5, 639.0, 639.01, 700.0)] == [0, 0, 1, 2, 639, 639, 640, 640]


def test_projection_equals_reference_build(O):
    """RGBD::point3Dto2D (RGBD.cpp:92-98) through the reprojection error versions: covered by the RANSAC test (ev 1, 2);
    here the function alone against the expression the oracle uses"""
    rng = np.random.default_rng(2)
    p = np.stack([rng.uniform(-2, 2, 500), rng.uniform(-2, 2, 500), rng.uniform(0.5, 6, 500)], 1).astype(np.float32)
    fx, fy, cx, cy = (np.float32(v) for v in cam())
    uv = R.point3Dto2D(p)
    assert np.array_equal(bits(uv[:, 0]), bits(p[:, 0] * fx / p[:, 2] + cx))
    assert np.array_equal(bits(uv[:, 1]), bits(p[:, 1] * fy / p[:, 2] + cy))


def test_sensor_model_equals_reference_build(O):
    """DepthSensorModel::computeCov / informationMatrixFromImageCoordinates / inverseModel / uncertinatyFromNormal /
    uncertinatyFromRGBGradient compiled from depthSensorModel.cpp == the oracle (float64; bit for bit where the stand-in's
    3x3 product order equals the oracle's, 1e-12 otherwise)"""
    from putslam_b200 import synth
    rng = np.random.default_rng(8)
    s = R.sensor_args()
    for _ in range(300):
        u, v, z = rng.uniform(0, 639), rng.uniform(0, 479), rng.uniform(0.8, 6.0)
        cov = O.compute_cov(int(u), int(v), z, *cam(), synth.VAR_U, synth.VAR_V, synth.DIST_VAR_COEFS)
        rc = R.compute_cov(int(u), int(v), z, s)
        assert np.allclose(cov, rc, rtol=1e-13, atol=0)
        c2, info = O.information_matrix(u, v, z, *cam(), synth.VAR_U, synth.VAR_V, synth.DIST_VAR_COEFS)
        ri = R.information_matrix_uvz(u, v, z, s)
        assert np.allclose(info, ri, rtol=1e-11, atol=0)
    # inverseModel: projection with the validity box (0 <= u <= W, 0 <= v <= H, 0.8 <= z <= 6 else (-1,-1,-1))
    for x, y, z in [(0.1, -0.2, 2.0), (3.0, 0.0, 1.0), (0.0, 0.0, 0.5), (0.0, 0.0, 6.5), (-2.0, 0.1, 1.5), (0.0, 2.4, 1.0)]:
        r = R.inverse_model(x, y, z, s)
        uu, vv = synth.FX * x / z + synth.CX, synth.FY * y / z + synth.CY
        if uu < 0 or uu > 640 or vv < 0 or vv > 480 or z < 0.8 or z > 6.0:
            assert (r == -1).all()
        else:
            assert np.allclose(r, [uu, vv, z], rtol=1e-15)
    for _ in range(100):
        n = rng.standard_normal(3); n /= np.linalg.norm(n)
        assert np.allclose(O.uncertainty_from_normal(n, 0.8), R.uncertainty_from_normal(n, s), rtol=0, atol=1e-12)
        g = rng.standard_normal(3); g /= np.linalg.norm(g)
        assert np.allclose(O.uncertainty_from_gradient(g, 0.8), R.uncertainty_from_gradient(g, s), rtol=0, atol=1e-12)


def test_normals_and_gradients_equal_reference_build(O):
    """RGBD::computeNormal / computeRGBGradient compiled from RGBD.cpp == the oracle, bit for bit, on noisy depth with holes"""
    rng = np.random.default_rng(11)
    H, W = 480, 640
    uu, vv = np.meshgrid(np.arange(W), np.arange(H))
    z = 2.0 + 0.002 * uu - 0.001 * vv + rng.normal(0, 0.003, (H, W))
    depth = np.rint(z * 5000).astype(np.uint16)
    depth[rng.random((H, W)) < 0.1] = 0
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    for _ in range(300):
        u, v = int(rng.integers(2, W - 2)), int(rng.integers(2, H - 2))
        a = O.compute_normal(depth, u, v, *cam(), 5000.0); b = R.compute_normal(depth, u, v)
        assert np.array_equal(bits(a), bits(b)) or (np.isnan(a).all() and np.isnan(b).all()), (u, v, a, b)
        a = O.compute_rgb_gradient(rgb, depth, u, v, *cam(), 5000.0); b = R.compute_rgb_gradient(rgb, depth, u, v)
        assert np.array_equal(bits(a), bits(b)) or (np.isnan(a) == np.isnan(b)).all(), (u, v, a, b)
    for u, v in [(0, 5), (1, 5), (5, 1), (W - 1, 5), (5, H - 1)]:       # the border test of computeRGBGradient (:154)
        assert (R.compute_rgb_gradient(rgb, depth, u, v) == 1.0).all() and (O.compute_rgb_gradient(rgb, depth, u, v, *cam(), 5000.0) == 1.0).all()


def test_kabsch_equals_reference_build(O):
    """KabschEst::computeTransformation compiled from kabschEst.cpp == orc_kabsch (float64, bit for bit), N = 3..400"""
    from putslam_b200 import synth
    rng = np.random.default_rng(4)
    for n in (3, 4, 10, 100, 400):
        for _ in range(10):
            A = rng.uniform(-2, 2, (n, 3))
            Rm = synth.rot_from_rotvec(rng.standard_normal(3) * 0.4)
            B = A @ Rm.T + rng.uniform(-1, 1, 3) + rng.normal(0, 0.01, (n, 3))
            a, b = O.kabsch(A, B), R.kabsch(A, B)
            assert np.array_equal(bits(a), bits(b)), n
    assert np.array_equal(R.kabsch(np.zeros((0, 3)), np.zeros((0, 3))), np.eye(4)[:3])       # empty -> identity (:28)


def test_transform_uncertainty_equals_reference_build():
    """TransformEst::computeUncertainty / computeUncertaintyG2O compiled from transformEst.h (not the Python evaluation of its
    statements used for tests/golden/uncertainty_ref.npz) == the derived oracle, and == those golden vectors"""
    import os
    from oracle import uncertainty_oracle as U
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "uncertainty_ref.npz"))
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    for name in g["names"]:
        for mode_i, mode in enumerate(("euler", "quat")):
            A, B, CA, CB, T = (g[f"{name}_{k}"] for k in ("A", "B", "CA", "CB", "T"))
            ref = R.transform_uncertainty(A, B, CA, CB, T, mode_i)
            assert rel(ref, g[f"{name}_{mode}_U"]) < 1e-8, (name, mode)
            assert rel(U.compute_uncertainty(A, B, CA, CB, T, mode)[0], ref) < 1e-8, (name, mode)


def _xyz_inputs(seed, M=1500, N=400, n_reobs=260):
    from putslam_b200 import host, synth
    mf = synth.map_frame(M=M, N=N, n_reobs=n_reobs, seed=seed)
    if seed % 2:
        mf["cur_detdist"] = mf["cur_detdist"].copy(); mf["cur_detdist"][::5] *= 1.3
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    return mf, ml, cl


def test_match_xyz_equals_reference_build(O):
    """Matcher::matchXYZ compiled from matcher.cpp (gates, level prediction with the reference's own pow / log / ceil, the
    saturating-difference distance, accept ratio, RANSAC with errorVersionMap, pointInlierRatio) == the oracle chain
    guided_match -> ransac.  Retry numbers widen the gates (:619-622)."""
    for seed in range(6):
        mf, ml, cl = _xyz_inputs(seed)
        for comp in (1, 2, 5):
            radius = 0.12 + 0.02 * (comp - 1); ratio = max(0.1, 0.55 - 0.05 * (comp - 1))
            ev = (0, 4, 2)[seed % 3]
            q, t, d, perfect = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, radius, ratio, 0)
            p = O.default_ransac_params(ev)
            o = O.ransac(mf["map_xyz"].astype(np.float32), mf["cur_xyz"], q, t, params=p, seed=seed + comp)
            a = R.matcher_args(ransac=R.from_oracle_params(p))
            r = R.match_xyz(mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], mf["cur_xyz"], mf["cur_desc"],
                            mf["cur_octave"], mf["cur_detdist"], args=a, computation_number=comp, seed=seed + comp,
                            use_frame_ids=bool(seed % 2))
            assert r["n_matches"] == q.size and r["n_perfect"] == perfect
            assert np.array_equal(r["pairs"], np.stack([q[o["inliers"]], t[o["inliers"]]], 1))
            assert r["hyp_used"] == o["hyp_used"] and np.array_equal(bits(r["T"]), bits(o["T"]))
            assert r["ratio"] == O.point_inlier_ratio(t[o["inliers"]], t, mf["cur_xyz"].shape[0])
        # the match list itself (pairs and order): with an enormous threshold every match is an inlier of the first model
        big = R.matcher_args(ransac=R.ransac_args(thr_e=1e9, min_ratio=0.0, min_matches=1))
        q, t, d, _ = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0)
        r = R.match_xyz(mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], mf["cur_xyz"], mf["cur_desc"],
                        mf["cur_octave"], mf["cur_detdist"], args=big, seed=1)
        assert np.array_equal(r["pairs"], np.stack([q, t], 1)) and q.size > 100
    # no candidate inside the sphere -> -1.0 (:755-756)
    mf, ml, cl = _xyz_inputs(0, M=50, N=40, n_reobs=10)
    far = mf["cur_xyz"] + np.float32([10, 0, 0])
    r = R.match_xyz(mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], far, mf["cur_desc"], mf["cur_octave"],
                    mf["cur_detdist"], seed=1)
    assert r["ratio"] == -1.0 and r["pairs"].size == 0


def test_match_vo_equals_reference_build(O):
    """Matcher::match compiled from matcher.cpp (DBScan -> describe -> performMatching -> removeImageDistortion ->
    keypoints2Dto3D -> RANSAC with errorVersionVO -> pointInlierRatio; state swap) == the oracle chain"""
    from putslam_b200 import synth
    for seed in range(5):
        fp = synth.frame_pair(n=350, seed=20 + seed, distorted=bool(seed % 2))
        keep1 = (fp["uv1"][:, 0] < 630) & (fp["uv1"][:, 1] < 470); keep2 = (fp["uv2"][:, 0] < 630) & (fp["uv2"][:, 1] < 470)
        uv1, d1 = fp["uv1"][keep1], fp["desc1"][keep1]; uv2, d2 = fp["uv2"][keep2], fp["desc2"][keep2]
        dist = synth.DIST if seed % 2 else (0, 0, 0, 0, 0)
        und1 = O.undistort(uv1, *cam(), dist)      # the reference undistorts even with zero coefficients (not an identity in float)
        und2 = O.undistort(uv2, *cam(), dist)
        x1, _ = O.backproject(und1, fp["depth1"], *cam(), 5000.0)
        x2, _ = O.backproject(und2, fp["depth2"], *cam(), 5000.0)
        oq, ot, od = O.bf_mutual(d1, d2)
        ev = (0, 1, 2, 4, 0)[seed]
        p = O.default_ransac_params(ev)
        o = O.ransac(x1, x2, oq, ot, params=p, seed=seed)
        a = R.matcher_args(ransac=R.from_oracle_params(p), dist=dist, dbscan_eps=0.0)      # eps 0: DBScan keeps every key point
        r = R.match_vo(d1, x1, uv2, np.zeros(len(uv2), np.int32), d2, fp["depth2"], args=a, seed=seed)
        assert np.array_equal(r["kept"], np.arange(len(uv2))) and np.array_equal(bits(r["xyz"]), bits(x2))
        assert np.array_equal(bits(r["uv"]), bits(np.ascontiguousarray(und2, np.float32)))
        assert np.array_equal(r["inliers"], np.stack([oq[o["inliers"]], ot[o["inliers"]]], 1))
        assert r["hyp_used"] == o["hyp_used"] and np.array_equal(bits(r["T"]), bits(o["T"]))
        assert r["ratio"] == O.point_inlier_ratio(ot[o["inliers"]], ot, len(uv2))


def test_loop_closure_equals_reference_build(O):
    """Matcher::matchFeatureLoopClosure compiled from matcher.cpp == bf_mutual -> ransac, and its return conventions"""
    from putslam_b200 import synth
    for seed, n in ((3, 300), (4, 35), (5, 120)):
        fp = synth.frame_pair(n=n, seed=seed)
        x1, _ = O.backproject(fp["uv1"], fp["depth1"], *cam(), 5000.0)
        x2, _ = O.backproject(fp["uv2"], fp["depth2"], *cam(), 5000.0)
        oq, ot, od = O.bf_mutual(fp["desc1"], fp["desc2"])
        p = O.default_ransac_params(0)
        o = O.ransac(x1, x2, oq, ot, params=p, seed=9)
        r = R.loop_closure(fp["desc1"], x1.astype(np.float64), fp["desc2"], x2.astype(np.float64),
                           args=R.matcher_args(ransac=R.from_oracle_params(p)), seed=9)
        assert np.array_equal(r["pairs"], np.stack([oq[o["inliers"]], ot[o["inliers"]]], 1))
        assert r["hyp_used"] == o["hyp_used"] and np.array_equal(bits(r["T"]), bits(o["T"]))
        assert r["ret"] == O.point_inlier_ratio(ot[o["inliers"]], ot, n)
    # fewer than 10 features on one side -> 0 (:830-834)
    r = R.loop_closure(fp["desc1"][:9], x1[:9].astype(np.float64), fp["desc2"], x2.astype(np.float64), seed=1)
    assert r["ret"] == 0 and r["pairs"].size == 0


def test_reference_build_variants_are_what_they_say():
    for v, want in (("", "fixed_redux_tree=1 product_coeff_seq=0 umeyama_scale_lhs=0 jacobi_threshold32=0 jacobi_sweep_alt=0 umeyama_f64=0"),
                    ("_seq", "product_coeff_seq=1"), ("_scalelhs", "umeyama_scale_lhs=1"), ("_jac32", "jacobi_threshold32=1"),
                    ("_sweep", "jacobi_sweep_alt=1"), ("_f64", "umeyama_f64=1"), ("_allseq", "fixed_redux_tree=0")):
        if R.available(v):
            assert want in R.shim_model(v)


def test_tracking_list_edits_equal_reference_build():
    """Matcher::removeTooCloseFeatures / mergeTrackedFeatures compiled from matcher.cpp == the numpy restatements
    (oracle/klt_oracle.py) that tests/test_abi_cpu.py holds the adapter's grid-accelerated versions to"""
    from oracle import klt_oracle as K
    rng = np.random.default_rng(41)
    for n, spread, e3, r2 in ((400, 120.0, 0.01, 3.0), (300, 60.0, 0.05, 1.0), (200, 500.0, 0.0, 0.0), (50, 30.0, 1e9, 2.0),
                              (0, 10.0, 0.01, 3.0)):
        und = rng.uniform(0, spread, (n, 2)).astype(np.float32)
        if n > 20:
            und[5] = und[3]; und[9] = und[3] + np.float32(0.5); und[n - 1] = und[0]
        z = rng.uniform(0.8, 5.0, n)
        xyz = np.stack([(und[:, 0] - spread / 2) / 500.0 * z, (und[:, 1] - spread / 2) / 500.0 * z, z], 1).astype(np.float32)
        mq = rng.integers(0, 1000, n).astype(np.int32); mt = rng.permutation(n).astype(np.int32)
        gone = K.remove_too_close(und, xyz, e3, r2)
        kept, rq, rt = R.remove_too_close(und, und, xyz, mq, mt, e3, r2)
        assert np.array_equal(kept, np.setdiff1d(np.arange(n), gone)), n
        sel = ~np.isin(mt, gone)
        assert np.array_equal(rq, mq[sel]) and np.array_equal(rt, mt[sel])
    for n, m, spread, thr in ((150, 500, 200.0, 3.0), (0, 300, 60.0, 5.0), (100, 0, 50.0, 3.0), (80, 200, 100.0, 0.0), (60, 400, 30.0, 2.5)):
        und = rng.uniform(0, spread, (n, 2)).astype(np.float32); s_und = rng.uniform(0, spread, (m, 2)).astype(np.float32)
        if m > 10 and n > 10:
            s_und[2] = und[1]; s_und[4] = s_und[3] + np.float32(0.25)
        assert np.array_equal(R.merge_tracked(und, s_und, thr), K.merge_tracked(und, s_und, thr)), (n, m)
