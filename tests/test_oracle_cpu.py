"""CPU tests (-m "not gpu"): the oracle against the cv2 golden vectors, live cv2, numpy float64."""
import numpy as np
import pytest

from conftest import bits


def test_bf_mutual_and_knn2_match_cv2_golden(O, golden):
    g = golden["bf_cv2"]
    for name in g["names"]:
        q, t = g[f"{name}_q"], g[f"{name}_t"]
        oq, ot, od = O.bf_mutual(q, t)
        assert np.array_equal(oq, g[f"{name}_mq"]), name
        assert np.array_equal(ot, g[f"{name}_mt"]), name
        assert np.array_equal(od, g[f"{name}_md"]), name
        oq2, ot2, od2 = O.bf_mutual(q, t, threads=3)
        assert np.array_equal(oq, oq2) and np.array_equal(ot, ot2) and np.array_equal(od, od2)
        if f"{name}_k_idx" in g:
            idx, dist = O.knn2(q, t)
            assert np.array_equal(idx, g[f"{name}_k_idx"]), name
            assert np.array_equal(dist.astype(np.float32), g[f"{name}_k_dist"]), name


def test_bf_mutual_live_cv2(O):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for nq, nt in [(257, 511), (640, 129)]:
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        ms = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t)
        oq, ot, od = O.bf_mutual(q, t)
        assert [m.queryIdx for m in ms] == oq.tolist()
        assert [m.trainIdx for m in ms] == ot.tolist()
        assert [m.distance for m in ms] == od.tolist()


def test_empty_inputs(O):
    e = np.zeros((0, 32), np.uint8)
    q = np.zeros((5, 32), np.uint8)
    assert O.bf_mutual(e, q)[0].size == 0 and O.bf_mutual(q, e)[0].size == 0
    idx, dist = O.knn2(q, np.zeros((1, 32), np.uint8))
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all()


def test_satsub_quirk_matches_cv2_golden(O, golden):
    g = golden["satsub_cv2"]
    for a, b, quirk, xor in zip(g["a"], g["b"], g["quirk"], g["xor"]):
        assert O.hamming_satsub(a, b) == int(quirk)
        assert O.hamming_xor(a, b) == int(xor)
    # the quirk is not symmetric and not the Hamming distance (SURVEY finding 4)
    assert any(O.hamming_satsub(a, b) != O.hamming_satsub(b, a) for a, b in zip(g["a"], g["b"]))


def test_undistort_matches_cv2_golden(O, golden):
    g = golden["undistort_cv2"]
    K = g["K"]
    out = O.undistort(g["uv"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], g["dist"])
    assert np.array_equal(bits(out), bits(g["uv_undist"]))


def test_backproject_definition(O):
    from putslam_b200 import synth
    fp = synth.frame_pair(n=200, seed=1)
    xyz, dd = O.backproject(fp["uv1"], fp["depth1"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DEPTH_SCALE)
    uv = fp["uv1"]
    uR = np.clip(np.rint(uv[:, 0]), 0, 639).astype(int); vR = np.clip(np.rint(uv[:, 1]), 0, 479).astype(int)
    Z = (fp["depth1"][vR, uR].astype(np.float64) / 5000.0).astype(np.float32)
    X = ((uv[:, 0] - np.float32(synth.CX)) / np.float32(synth.FX)) * Z
    assert np.array_equal(bits(xyz[:, 2]), bits(Z)) and np.array_equal(bits(xyz[:, 0]), bits(X.astype(np.float32)))
    assert np.allclose(dd, np.linalg.norm(xyz.astype(np.float64), axis=1), rtol=1e-6)
    # zero depth -> origin (later dropped by RANSAC's z < 0.1 filter)
    z = np.zeros((480, 640), np.uint16)
    xyz0, _ = O.backproject(uv[:3], z, synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
    assert (xyz0 == 0).all()


def test_cov_is_J_R_Jt(O):
    from putslam_b200 import synth
    c = synth.DIST_VAR_COEFS
    cov = O.compute_cov(321, 200, 2.5, synth.FX, synth.FY, synth.CX, synth.CY, synth.VAR_U, synth.VAR_V, c)
    d = 2.5
    J = np.array([[d / synth.FX, 0, 321 / synth.FX - synth.CX / synth.FX], [0, d / synth.FY, 200 / synth.FY - synth.CY / synth.FY], [0, 0, 1]])
    R = np.diag([synth.VAR_U, synth.VAR_V, c[0] * d ** 3 + c[1] * d ** 2 + c[2] * d + c[3]])
    assert np.allclose(cov, J @ R @ J.T, rtol=1e-13, atol=0)


def test_information_matrix_is_inverse(O):
    from putslam_b200 import synth
    cov, info = O.information_matrix(321.7, 200.2, 2.5, synth.FX, synth.FY, synth.CX, synth.CY, synth.VAR_U, synth.VAR_V,
                                     synth.DIST_VAR_COEFS)
    ref = O.compute_cov(321, 200, 2.5, synth.FX, synth.FY, synth.CX, synth.CY, synth.VAR_U, synth.VAR_V, synth.DIST_VAR_COEFS)
    assert np.array_equal(cov, ref)
    assert np.allclose(info, np.linalg.inv(cov), rtol=1e-10)


def test_normal_and_uncertainty_from_normal(O):
    from putslam_b200 import synth
    uu, vv = np.meshgrid(np.arange(640), np.arange(480))
    z = 2.0 + 0.002 * uu                       # plane tilted about the image y axis
    depth = np.rint(z * 5000).astype(np.uint16)
    n = O.compute_normal(depth, 300, 200, synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
    assert abs(np.linalg.norm(n) - 1) < 1e-12 and abs(n[1]) < 0.05 and abs(n[2]) > 0.5
    cov = O.uncertainty_from_normal(n, 0.8)
    w = np.linalg.eigvals(cov).real
    assert np.allclose(sorted(w), [0.64, 1.0, 1.0], atol=1e-9)       # S^2 in the frame of the normal
    assert np.allclose(cov @ n, 0.64 * n, atol=1e-9)
    depth[:] = 0
    assert np.isnan(O.compute_normal(depth, 300, 200, synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)).all()


def _py_rgb_gradient_dir(rgb, u, v):
    """independent restatement of the direction part of RGBD::computeRGBGradient (src/RGBD/RGBD.cpp:154-166)"""
    import math
    row = lambda r: rgb[v - 1 + r].reshape(-1)[3 * (u - 1):3 * (u - 1) + 6].view("<u2").astype(np.int64)
    p = np.stack([row(0), row(1), row(2)])
    gx = -3 * p[0, 0] - 10 * p[0, 1] - 3 * p[0, 2] + 3 * p[2, 0] + 10 * p[2, 1] + 3 * p[2, 2]
    gy = -3 * p[0, 0] - 10 * p[1, 0] - 3 * p[2, 0] + 3 * p[0, 2] + 10 * p[1, 2] + 3 * p[2, 2]
    a = math.atan2(float(gy), float(gx)) + math.pi / 2.0
    c1 = (int(math.sqrt(2) * math.sin(a)), int(math.sqrt(2) * math.cos(a)))
    c2 = (int(math.sqrt(2) * math.sin(a + math.pi)), int(math.sqrt(2) * math.cos(a + math.pi)))
    return int(gx), int(gy), c1, c2


def test_rgb_gradient_and_uncertainty_from_gradient(O):
    """Uncertainty model 2 (src/RGBD/RGBD.cpp:147-187, src/Grabber/depthSensorModel.cpp:79-95)."""
    from putslam_b200 import synth
    rng = np.random.default_rng(21)
    H, W = 480, 640
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    depth = np.full((H, W), 10000, np.uint16)                  # fronto-parallel wall at 2 m
    cam = (synth.FX, synth.FY, synth.CX, synth.CY)
    for _ in range(200):
        u, v = int(rng.integers(2, W - 1)), int(rng.integers(2, H - 1))
        gx, gy, c1, c2 = _py_rgb_gradient_dir(rgb, u, v)
        g = O.compute_rgb_gradient(rgb, depth, u, v, *cam, 5000.0)
        # on a wall both end points have depth: grad is the normalised 3-D difference of pixels u+c1 and u+c2
        d = np.array([(c1[0] - c2[0]) / synth.FX, (c1[1] - c2[1]) / synth.FY, 0.0]) * 2.0
        assert np.allclose(g, d / np.linalg.norm(d), atol=1e-5), (u, v, gx, gy, c1, c2, g)
        assert abs(np.linalg.norm(g) - 1) < 1e-12
    # border test (:154): u-1 > 0 ... else the un-normalised (1,1,1)
    for u, v in [(0, 5), (1, 5), (5, 1), (W - 1, 5), (5, H - 1)]:
        assert (O.compute_rgb_gradient(rgb, depth, u, v, *cam, 5000.0) == 1.0).all()
    # no depth anywhere -> (coord1, 0) normalised
    g = O.compute_rgb_gradient(rgb, np.zeros_like(depth), 100, 100, *cam, 5000.0)
    _, _, c1, _ = _py_rgb_gradient_dir(rgb, 100, 100)
    assert np.allclose(g, np.array([c1[0], c1[1], 0.0]) / np.hypot(*c1))
    # the libm-sensitive diagonals: the table equals the direct evaluation, for any magnitude
    tab = O.gradient_diag_table()
    for k in (1, 7, 255, 65535):
        for q, (r, c) in enumerate([(2, 2), (0, 2), (2, 0), (0, 0)]):      # q = (gx<0) + 2 (gy<0)
            img = np.zeros((H, W, 3), np.uint8)
            base = 3 * (50 - 1) + 2 * c
            img[60 - 1 + r].reshape(-1)[base:base + 2] = [k & 255, k >> 8]
            gx, gy, c1, c2 = _py_rgb_gradient_dir(img, 50, 60)
            assert abs(gx) == abs(gy) == 3 * k and (gx < 0) + 2 * (gy < 0) == q
            assert tuple(tab[q]) == c1 + c2
    # cov = R S^2 R^-1 with S = diag(1, s, 1): eigenvalue s^2 along y = z x grad, 1 along grad
    g = np.array([0.6, 0.0, 0.8])
    cov = O.uncertainty_from_gradient(g, 0.8)
    y = np.cross([0, 0, 1.0], g); y /= np.linalg.norm(y)
    assert np.allclose(cov @ y, 0.64 * y, atol=1e-12) and np.allclose(cov @ g, g, atol=1e-12)
    assert np.allclose(O.inverse3d(cov) @ cov, np.eye(3), atol=1e-12)


def test_map_prepare_against_numpy(O):
    from putslam_b200 import synth
    rng = np.random.default_rng(3)
    pose = np.eye(4); pose[:3, :3] = synth.rot_from_rotvec([0.3, -0.2, 0.1]); pose[:3, 3] = [0.5, 0.2, -1.0]
    M = 300
    local = np.stack([rng.uniform(-2, 2, M), rng.uniform(-1, 1, M), rng.uniform(0.5, 6.5, M)], 1)
    glob = local @ pose[:3, :3].T + pose[:3, 3]
    axes = np.stack([synth.rot_from_rotvec(rng.normal(0, 0.5, 3))[:, 2] for _ in range(M)]).astype(np.float32)
    kept, xl, uv, ang = O.map_prepare(glob, axes, pose, synth.FX, synth.FY, synth.CX, synth.CY, 640, 480, 0.6, 5.0)
    zc = pose[:3, 2]
    a = np.arccos(np.clip((axes.astype(np.float64) @ zc) / np.linalg.norm(axes.astype(np.float64), axis=1), -1, 1))
    exp = np.nonzero((a <= 0.6) & (local[:, 2] <= 5.0))[0]
    border = np.abs(a - 0.6) < 1e-6
    assert set(kept.tolist()) ^ set(exp.tolist()) <= set(np.nonzero(border)[0].tolist())
    assert np.abs(xl - local[kept]).max() < 1e-12 and np.abs(ang - a[kept]).max() < 1e-6
    u = synth.FX * local[kept, 0] / local[kept, 2] + synth.CX
    ok = uv[:, 0] >= 0
    assert np.abs(uv[ok, 0] - u[ok]).max() < 1e-9


def test_philox_known_answers(O):
    # Random123 kat_vectors, philox4x32-10
    assert O.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sample3_distinct_and_in_range(O):
    for m in (3, 4, 15, 1000):
        for h in range(50):
            s = O.sample3(12345, h, m)
            assert len(set(s.tolist())) == 3 and s.min() >= 0 and s.max() < m


def test_svd_against_numpy(O):
    rng = np.random.default_rng(0)
    for _ in range(50):
        A = rng.standard_normal((3, 3)).astype(np.float32)
        U, S, V = O.svd3f(A)
        assert np.abs(U @ np.diag(S) @ V.T - A).max() < 5e-6
        assert np.allclose(S, np.linalg.svd(A.astype(np.float64))[1], atol=5e-6)
        assert S[0] >= S[1] >= S[2] >= 0
        assert np.abs(U.T @ U - np.eye(3)).max() < 5e-6 and np.abs(V.T @ V - np.eye(3)).max() < 5e-6
    Ud, Sd, Vd = O.svd3d(np.diag([3.0, 0.0, 1.0]))
    assert np.allclose(Sd, [3, 1, 0])
    # rank-2 (three centred points are coplanar) and rank-1 inputs stay finite
    B = np.outer([1, 2, 3], [4, 5, 6]).astype(np.float32)
    U, S, V = O.svd3f(B)
    assert np.isfinite(U).all() and np.isfinite(V).all() and S[1] < 1e-4 * S[0]


def _umeyama64(src, dst):
    sm, dm = src.mean(0), dst.mean(0)
    sig = (dst - dm).T @ (src - sm) / len(src)
    U, S, Vt = np.linalg.svd(sig)
    D = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        D[2, 2] = -1
    R = U @ D @ Vt
    return R, dm - R @ sm


def test_umeyama_and_kabsch_against_float64(O):
    from putslam_b200 import synth
    rng = np.random.default_rng(1)
    for n in (3, 4, 10, 100, 1000):
        src = rng.uniform(-1.5, 1.5, (n, 3))
        R = synth.rot_from_rotvec(rng.standard_normal(3) * 0.3)
        t = np.array([0.1, 0.2, -0.3])
        dst = src @ R.T + t + rng.normal(0, 0.01, (n, 3))
        ok, T = O.umeyama(src.astype(np.float32), dst.astype(np.float32))
        R64, t64 = _umeyama64(src.astype(np.float32).astype(np.float64), dst.astype(np.float32).astype(np.float64))
        assert ok == 1
        assert np.abs(T[:3, :3] - R64).max() < 1e-5 and np.abs(T[:3, 3] - t64).max() < 1e-5
        if n >= 4:
            Tk = O.kabsch(src, dst)
            Rk, tk = _umeyama64(src, dst)
            assert np.abs(Tk[:, :3] - Rk).max() < 1e-10 and np.abs(Tk[:, 3] - tk).max() < 1e-10
    # reflection case: planar mirrored configuration must still give det(R) = +1
    src = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    dst = np.array([[0, 0, 0], [1, 0, 0], [0, -1, 0]], np.float32)
    ok, T = O.umeyama(src, dst)
    assert ok == 1 and abs(np.linalg.det(T[:3, :3].astype(np.float64)) - 1) < 1e-5
    assert np.array_equal(O.kabsch(np.zeros((0, 3)), np.zeros((0, 3))), np.eye(3, 4))


def test_inverse4_against_numpy(O):
    from putslam_b200 import synth
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = synth.rot_from_rotvec([0.2, -0.1, 0.3]); T[:3, 3] = [0.3, -0.2, 0.1]
    assert np.abs(O.inverse4(T) - np.linalg.inv(T.astype(np.float64))).max() < 1e-6


def test_ransac_iterations(O):
    assert O.ransac_iterations(0.2) == 487            # reference RANSAC.cpp:30
    assert O.ransac_iterations(1.0) == 0              # log(0) = -inf
    assert O.ransac_iterations(0.0005) == 2**31 - 1   # out of int range: saturating (UB in the reference)
    assert O.ransac_iterations(0.5) == int(np.log(0.02) / np.log(1 - 0.125))


def test_usac_standard_stopping(O):
    import math
    assert O.lib().orc_usac_stopping(0, 100, 850000) == 850000 and O.lib().orc_usac_stopping(2, 100, 850000) == 850000
    assert O.lib().orc_usac_stopping(100, 100, 850000) == 1
    p = (50 * 49 * 48) / (100 * 99 * 98)
    assert O.lib().orc_usac_stopping(50, 100, 850000) == math.ceil(math.log(0.01) / math.log(1 - p))


def test_ransac_recovers_planted_transform(O):
    from putslam_b200 import synth
    mc = synth.matched_clouds(m=600, inlier_frac=0.6, seed=2)
    for ev in (0, 4, 1, 2):
        p = O.default_ransac_params(ev)
        r = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=p, seed=9, num_hyp=0, want_counts=True)
        assert len(r["inliers"]) > 250, ev
        assert np.abs(r["T"] - mc["T_gt"]).max() < 0.02, ev
        assert 0 < r["hyp_used"] <= 487
        assert np.all(np.diff(r["inliers"]) > 0)
    r1 = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=9, num_hyp=256)
    r2 = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=9, num_hyp=256)
    assert np.array_equal(r1["inliers"], r2["inliers"]) and r1["hyp_used"] == 256


def test_ransac_failure_conventions(O):
    from putslam_b200 import synth
    mc = synth.matched_clouds(m=600, inlier_frac=0.6, seed=2)
    # fewer than minimalNumberOfMatches -> identity, no inliers (RANSAC.cpp:77-80)
    r = O.ransac(mc["prev"], mc["cur"], mc["mq"][:10], mc["mt"][:10], seed=1)
    assert np.array_equal(r["T"], np.eye(4)) and r["inliers"].size == 0
    # pure outliers -> ratio below 0.2 -> identity (RANSAC.cpp:161-164)
    rng = np.random.default_rng(0)
    prev = rng.uniform(0.5, 4, (300, 3)).astype(np.float32); cur = rng.uniform(0.5, 4, (300, 3)).astype(np.float32)
    r = O.ransac(prev, cur, np.arange(300), np.arange(300), seed=1)
    assert np.array_equal(r["T"], np.eye(4)) and r["inliers"].size == 0
    # NaN / out-of-range depth are filtered before sampling (RANSAC.cpp:65-74)
    prev2 = mc["prev"].copy(); prev2[::7, 2] = np.nan; prev2[1::7, 2] = 7.0; prev2[2::7, 2] = 0.05
    r = O.ransac(prev2, mc["cur"], mc["mq"], mc["mt"], seed=3)
    bad = np.isnan(prev2[mc["mq"], 2]) | (prev2[mc["mq"], 2] > 6) | (prev2[mc["mq"], 2] < 0.1)
    assert not bad[r["inliers"]].any()


def test_guided_match_against_python_restatement(O):
    from putslam_b200 import host, synth
    mf = synth.map_frame(M=300, N=120, n_reobs=80, seed=4)
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    for mode in (0, 1):
        q, t, d, perfect = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, mode)
        exp = []
        mx = mf["map_xyz"].astype(np.float32)
        for j in range(300):
            diff = mx[j] - mf["cur_xyz"]
            nrm = np.sqrt((diff[:, 0] ** 2 + (diff[:, 1] ** 2 + diff[:, 2] ** 2)).astype(np.float32))
            cand = [i for i in range(120) if nrm[i] < 0.12 and abs(int(cl[i]) - int(ml[j])) <= 1]
            if mode == 0:
                vals = [int(np.unpackbits(np.clip(mf["map_desc"][j].astype(int) - mf["cur_desc"][i].astype(int), 0, 255).astype(np.uint8)).sum()) for i in cand]
            else:
                vals = [int(np.unpackbits(mf["map_desc"][j] ^ mf["cur_desc"][i]).sum()) for i in cand]
            if cand:
                best = min(vals)
                exp += [(j, i, v) for i, v in zip(cand, vals) if 0.55 * v <= best]
        assert list(zip(q.tolist(), t.tolist(), d.astype(int).tolist())) == exp
        assert len(exp) > 40


def test_levels_host_equals_oracle(O):
    from putslam_b200 import host
    rng = np.random.default_rng(3)
    for _ in range(500):
        o = int(rng.integers(0, 8)); dd = float(rng.uniform(0.5, 6)); cd = float(dd * rng.uniform(0.6, 1.6))
        assert host.predicted_level(o, dd, cd) == O.pred_level(o, dd, cd)


def test_lc_scores_and_topk(O):
    from putslam_b200 import synth
    db = synth.keyframe_db(n_kf=40, per_kf=120, n_query=100, n_planted=4, shared=50, seed=3, ragged=True)
    s = O.lc_scores(db["query"], db["db"], db["kf_off"], tau=64, threads=2)
    top, sc = O.topk(s, 4)
    assert sorted(top.tolist()) == db["planted"].tolist()
    assert (sc >= 30).all() and (np.delete(s, db["planted"]) < 10).all()
    order = sorted(range(40), key=lambda i: (-s[i], i))[:8]
    assert O.topk(s, 8)[0].tolist() == order


def test_point_inlier_ratio(O):
    assert O.point_inlier_ratio([1, 1, 2], [1, 2, 3, 3, 4], 10) == 2 / 4


# ---------------------------------------------------------------- ORB descriptors (describeFeatures seam)
def test_orb_oracle_matches_cv2_golden(golden):
    """oracle/orb_oracle.py against cv::ORB::compute outputs recorded from cv2 4.13.0: kept/reordered keypoints and
    every descriptor bit."""
    from oracle import orb_oracle as OO
    g = golden["orb_cv2"]
    for name in g["names"]:
        order, desc = OO.describe(g[f"{name}_img"], g[f"{name}_xy"], g[f"{name}_octave"], g[f"{name}_angle"])
        assert np.array_equal(order, g[f"{name}_order"]), name
        assert np.array_equal(desc, g[f"{name}_desc"]), name
        assert 0.5 * len(g[f"{name}_xy"]) < order.size < len(g[f"{name}_xy"])        # the border filter was exercised


def test_orb_stages_against_live_cv2():
    """each stage of the restatement against the library call it stands for, on fresh inputs"""
    import cv2
    from oracle import orb_oracle as OO
    rng = np.random.default_rng(41)
    img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    smooth = cv2.GaussianBlur(img, (0, 0), 2.0)
    # 1. gray conversion
    bgr = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    assert np.array_equal(OO.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    # 3. the resize cascade: every level size ORB uses for 640x480, plus odd shapes
    prev = img
    for l in range(1, 8):
        w, h = OO.level_size(640, 480, l)
        ref = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(OO.resize_linear_exact(prev, w, h), ref), l
        prev = ref
    assert [OO.level_size(640, 480, l) for l in (1, 2, 7)] == [(533, 400), (444, 333), (179, 134)]
    for (sw, sh, dw, dh) in [(101, 77, 84, 64), (64, 64, 53, 53), (33, 200, 28, 167)]:
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        assert np.array_equal(OO.resize_linear_exact(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))
    # 4. the blur ORB applies to a pyramid level (sub-matrix => float separable filter): same kernel, same pixels
    kf = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    assert np.array_equal(kf, OO.GAUSS7)
    for im in (img, smooth):
        ref = cv2.sepFilter2D(im, cv2.CV_8U, kf, kf, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(OO.gaussian7_submatrix(im), ref)
    # 2 + 5. whole path, full-size frame, 600 keypoints over all octaves
    n = 600
    xy = np.stack([rng.uniform(10, 630, n), rng.uniform(10, 470, n)], 1).astype(np.float32)
    octave = rng.integers(0, 8, n).astype(np.int32); angle = rng.uniform(0, 360, n).astype(np.float32)
    kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, int(o)) for (x, y), a, o in zip(xy, angle, octave)]
    k2, d2 = cv2.ORB_create().compute(smooth, kps)
    order, desc = OO.describe(smooth, xy, octave, angle)
    assert order.size == len(k2)
    assert all(float(k.pt[0]) == float(xy[i, 0]) and k.octave == octave[i] for k, i in zip(k2, order))
    assert np.array_equal(desc, d2)
    # keypoints already sorted by level keep their order; an empty list gives an empty result
    srt = np.argsort(octave, kind="stable")
    o2, d3 = OO.describe(smooth, xy[srt], octave[srt], angle[srt])
    assert np.array_equal(srt[o2], order) and np.array_equal(d3, desc)
    assert OO.describe(smooth, np.zeros((0, 2)), np.zeros(0, np.int32), np.zeros(0))[0].size == 0


def test_orb_pattern_tables_agree():
    """the device's copy of the sampling pattern (putslam_b200/csrc/orb_pattern.inc) equals the oracle's"""
    import os
    import re
    from oracle import orb_oracle as OO
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "putslam_b200", "csrc", "orb_pattern.inc")).read()
    nums = [int(x) for line in txt.splitlines() if not line.lstrip().startswith("//") for x in re.findall(r"-?\d+", line)]
    assert len(nums) == 1024 and np.array_equal(np.array(nums).reshape(256, 4), OO.PATTERN)
    assert OO.PATTERN[0].tolist() == [8, -3, 9, 5] and np.abs(OO.PATTERN).max() == 13


def _kp_set(kp, octave):
    return {(float(r[0]), float(r[1]), int(o)): (float(r[2]), float(r[3]), float(r[4])) for r, o in zip(kp, octave)}


def test_orb_detect_oracle_matches_cv2_golden(golden):
    """cv::ORB::detect recorded from cv2 4.13.0: the oracle yields the same keypoints (position, octave) with the same
    size, angle and Harris response, level by level (OpenCV's order inside a level is an artefact of nth_element)."""
    from oracle import orb_oracle as OO
    g = golden["orb_detect_cv2"]
    for name in g["names"]:
        mine = OO.detect(g[f"{name}_img"], int(g[f"{name}_nfeatures"]))
        ref = _kp_set(g[f"{name}_kp"], g[f"{name}_octave"])
        got = {(float(m[0]), float(m[1]), int(m[5])): (float(m[2]), float(m[3]), float(m[4])) for m in mine}
        assert got == ref, name
        assert [m[5] for m in mine] == sorted(m[5] for m in mine)                       # grouped by level
        assert np.array_equal(np.bincount(g[f"{name}_octave"]), np.bincount([m[5] for m in mine]))


def test_orb_detect_stages_against_live_cv2():
    import cv2
    from oracle import orb_oracle as OO
    rng = np.random.default_rng(53)
    img = cv2.GaussianBlur(rng.integers(0, 256, (240, 320), dtype=np.uint8), (0, 0), 1.4)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    # FAST-9/16 with non-maximum suppression: positions, raster order and corner scores
    ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in cv2.FastFeatureDetector_create(20, True).detect(img)]
    assert OO.fast_nms(OO.fast_score_map(img, 20)) == ref and len(ref) > 200
    # fastAtan2
    for _ in range(20000):
        y = float(rng.integers(-300000, 300000)); x = float(rng.integers(-300000, 300000))
        assert float(OO.fast_atan2(y, x)) == cv2.fastAtan2(y, x)
    assert float(OO.fast_atan2(0.0, 0.0)) == cv2.fastAtan2(0.0, 0.0)
    # the split of nfeatures over the levels and the circular patch
    assert OO.features_per_level(500, 8) == [109, 90, 75, 63, 52, 44, 36, 31] and sum(OO.features_per_level(1000, 8)) == 1000
    assert OO.UMAX[:16] == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    # whole detector, several feature budgets
    for nf in (60, 500, 1500):
        kps = cv2.ORB_create(nfeatures=nf).detect(img)
        ref = _kp_set(np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kps], np.float32), [k.octave for k in kps])
        got = {(float(m[0]), float(m[1]), int(m[5])): (float(m[2]), float(m[3]), float(m[4])) for m in OO.detect(img, nf)}
        assert got == ref, nf


# ---------------------------------------------------------------- KLT tracking (performTracking seam)
def test_klt_oracle_matches_cv2_golden(golden):
    """oracle/klt_oracle.py against cv::calcOpticalFlowPyrLK recorded from cv2 4.13.0: tracked positions, status and
    err bit for bit -- colour and gray frames, minimum-eigenvalue error, initial flow"""
    from oracle import klt_oracle as K
    g = golden["klt_cv2"]
    a, b, pts = g["a"], g["b"], g["pts"]
    cases = {"colour": (a, b, {}), "gray": (a[..., 0], b[..., 0], {}), "mineig": (a, b, {"min_eig_err": True}),
             "initflow": (a, b, {"init": g["init"]})}
    for name in g["names"]:
        ia, ib, kw = cases[str(name)]
        nxt, st, err = K.lk_pyr(ia, ib, pts, **kw)
        assert np.array_equal(st, g[f"{name}_status"]), name
        ok = st == 1
        assert 60 < ok.sum() < len(pts)                                              # some points are lost at the borders
        assert np.array_equal(nxt[ok].view(np.uint32), g[f"{name}_next"][ok].view(np.uint32)), name
        assert np.array_equal(err[ok].view(np.uint32), g[f"{name}_err"][ok].view(np.uint32)), name


def test_klt_stages_and_extreme_contrast_against_live_cv2():
    import cv2
    from oracle import klt_oracle as K
    rng = np.random.default_rng(67)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    ref = img
    mine = img
    for _ in range(3):                                                               # pyrDown cascade, odd sizes
        ref = cv2.pyrDown(ref); mine = K.pyr_down(mine)
        assert np.array_equal(ref, mine)
    # blocks of 0 / 255: the float accumulation order of the window sums matters here (it is the scalar raster order)
    a = cv2.resize(rng.integers(0, 2, (60, 80), dtype=np.uint8) * 255, (160, 120), interpolation=cv2.INTER_NEAREST)
    b = cv2.warpAffine(a, np.float32([[1, 0, 0.3], [0, 1, 0.2]]), (160, 120), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = np.stack([rng.uniform(3, 157, 80), rng.uniform(3, 117, 80)], 1).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01)
    p1, st, er = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(7, 7), maxLevel=3, criteria=crit)
    nxt, ms, me = K.lk_pyr(a, b, pts)
    ok = st.ravel() == 1
    assert np.array_equal(ms, st.ravel()) and ok.sum() > 20
    assert np.array_equal(nxt[ok].view(np.uint32), p1.reshape(-1, 2)[ok].view(np.uint32))
    assert np.array_equal(me[ok].view(np.uint32), er.ravel()[ok].view(np.uint32))
    # colour frames with hard edges, several window sizes (27 / 15 / 21 values per window row): the lane order of the sums
    for win, lev in ((9, 2), (5, 1), (7, 3)):
        ca = cv2.resize(rng.integers(0, 2, (60, 80, 3), dtype=np.uint8) * 255, (160, 120), interpolation=cv2.INTER_NEAREST)
        cb = cv2.warpAffine(ca, np.float32([[1, 0, 0.4], [0, 1, -0.3]]), (160, 120), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        cp = pts[:40]
        p1, st, er = cv2.calcOpticalFlowPyrLK(ca, cb, cp.reshape(-1, 1, 2), None, winSize=(win, win), maxLevel=lev, criteria=crit)
        nxt, ms, me = K.lk_pyr(ca, cb, cp, win=win, max_level=lev)
        ok = st.ravel() == 1
        assert np.array_equal(ms, st.ravel()) and ok.sum() > 10
        assert np.array_equal(nxt[ok].view(np.uint32), p1.reshape(-1, 2)[ok].view(np.uint32)), win
        assert np.array_equal(me[ok].view(np.uint32), er.ravel()[ok].view(np.uint32)), win
    # another window size and pyramid depth
    g = cv2.GaussianBlur(rng.integers(0, 256, (120, 160), dtype=np.uint8), (0, 0), 1.5)
    g2 = cv2.warpAffine(g, np.float32([[1, 0, 1.7], [0, 1, -1.1]]), (160, 120), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    p1, st, er = cv2.calcOpticalFlowPyrLK(g, g2, pts.reshape(-1, 1, 2), None, winSize=(11, 11), maxLevel=2, criteria=crit)
    nxt, ms, me = K.lk_pyr(g, g2, pts, win=11, max_level=2)
    ok = st.ravel() == 1
    assert np.array_equal(ms, st.ravel()) and np.array_equal(nxt[ok].view(np.uint32), p1.reshape(-1, 2)[ok].view(np.uint32))
    assert np.array_equal(me[ok].view(np.uint32), er.ravel()[ok].view(np.uint32))


def test_klt_perform_tracking_postprocessing():
    """error threshold, pairwise too-close removal (the larger err goes, the second on ties, every feature takes part
    whatever its status), ordered survivors"""
    from oracle import klt_oracle as K
    pts = np.array([[10, 10], [10.5, 10], [50, 50], [50, 50.4], [90, 90], [10.2, 10.1]], np.float32)
    err = np.array([1.0, 2.0, 3.0, 3.0, 9.0, 0.5], np.float32)
    status = np.array([1, 1, 1, 1, 1, 0], np.uint8)
    kept = K.perform_tracking(err, status, pts, error_threshold=5.0, min_distance=1.0)
    # 0-1 close: 1 goes; 0-5 close: 0 goes (err 1.0 > 0.5); 1-5 close: 1 goes; 2-3 tie: 3 goes; 4 over the threshold; 5 lost
    assert kept.tolist() == [2]
