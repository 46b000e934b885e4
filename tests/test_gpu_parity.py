"""GPU parity tests (-m gpu): every kernel, called through the C ABI, against the oracle and the cv2
golden vectors.  Integer / index / inlier-set results are bit-exact; poses within 1e-5 m and 1e-5 rad
(north_star tolerance), float32 back-projection bit-exact, float64 covariance within 1e-12 relative."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-5      # north_star: 1e-5 m translation
POSE_TOL_RAD = 1e-5    # north_star: 1e-5 rad rotation


def rot_angle(Ra, Rb):
    R = Ra.astype(np.float64).T @ Rb.astype(np.float64)
    return float(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))


def assert_pose_close(Ta, Tb):
    assert np.abs(Ta[:3, 3].astype(np.float64) - Tb[:3, 3]).max() <= POSE_TOL_M
    # arccos near 1 has sqrt(eps) resolution; compare matrices, which bounds the angle
    assert np.abs(Ta[:3, :3].astype(np.float64) - Tb[:3, :3]).max() <= POSE_TOL_RAD


# ---------------------------------------------------------------- stage 2: brute force + cross-check
def test_bf_mutual_golden_cv2(ctx, golden):
    g = golden["bf_cv2"]
    for name in g["names"]:
        oq, ot, od = ctx.match_bf_mutual(g[f"{name}_q"], g[f"{name}_t"])
        assert np.array_equal(oq, g[f"{name}_mq"]), name
        assert np.array_equal(ot, g[f"{name}_mt"]), name
        assert np.array_equal(od, g[f"{name}_md"]), name


@pytest.mark.parametrize("nq,nt", [(500, 500), (1000, 1000), (1000, 5000), (5000, 1000), (1, 1), (33, 2047),
                                    (257, 255), (1025, 300), (2500, 2500)])
def test_bf_mutual_vs_oracle(ctx, O, nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    n = min(nq, nt) // 2
    if n:  # plant near-duplicates so that many mutual matches exist
        from putslam_b200 import synth
        t[rng.choice(nt, n, replace=False)] = synth.flip_bits(rng, q[rng.choice(nq, n, replace=False)], 0.05)
    a = ctx.match_bf_mutual(q, t)
    b = O.bf_mutual(q, t, threads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_bf_mutual_heavy_ties_and_empty(ctx, O):
    rng = np.random.default_rng(1)
    q = np.zeros((700, 32), np.uint8); t = np.zeros((900, 32), np.uint8)
    q[:, 0] = rng.integers(0, 4, 700); t[:, 0] = rng.integers(0, 4, 900)
    for x, y in zip(ctx.match_bf_mutual(q, t), O.bf_mutual(q, t)):
        assert np.array_equal(x, y)
    same = np.tile(rng.integers(0, 256, (1, 32), dtype=np.uint8), (300, 1))
    oq, ot, od = ctx.match_bf_mutual(same, same)   # all distances 0: only (0, 0) is mutual
    assert oq.tolist() == [0] and ot.tolist() == [0] and od.tolist() == [0.0]
    e = np.zeros((0, 32), np.uint8)
    assert ctx.match_bf_mutual(e, t)[0].size == 0 and ctx.match_bf_mutual(q, e)[0].size == 0


def test_bf_rejects_unsupported(ctx):
    from putslam_b200 import api
    with pytest.raises(api.PslamError) as ei:
        ctx.match_bf_mutual(np.zeros((4, 64), np.uint8), np.zeros((4, 64), np.uint8))
    assert ei.value.code == api.ERR_UNSUPPORTED


def test_knn2_golden_and_oracle(ctx, O, golden):
    g = golden["bf_cv2"]
    for name in g["names"]:
        if f"{name}_k_idx" not in g:
            continue
        idx, dist = ctx.match_knn2(g[f"{name}_q"], g[f"{name}_t"])
        assert np.array_equal(idx, g[f"{name}_k_idx"]), name
        assert np.array_equal(dist, g[f"{name}_k_dist"]), name
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (1000, 32), dtype=np.uint8); t = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
    idx, dist = ctx.match_knn2(q, t)
    oi, od = O.knn2(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od.astype(np.float32))
    idx, dist = ctx.match_knn2(q[:5], t[:1])
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all() and (dist[:, 1] == -1).all()


# ---------------------------------------------------------------- stage 1: back-projection
@pytest.mark.parametrize("distorted", [False, True])
def test_backproject_bit_exact(ctx, O, distorted):
    from putslam_b200 import api, synth
    fp = synth.frame_pair(n=1000, seed=5, distorted=distorted)
    cam = api.make_camera()
    out = ctx.backproject(fp["uv1"], fp["depth1"], cam=cam, undistort=distorted)
    uv = fp["uv1"]
    if distorted:
        uv = O.undistort(uv, synth.FX, synth.FY, synth.CX, synth.CY, synth.DIST)
        assert np.array_equal(bits(out["uv_undist"]), bits(uv))
    xyz, dd = O.backproject(uv, fp["depth1"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DEPTH_SCALE)
    assert np.array_equal(bits(out["xyz"]), bits(xyz))
    assert np.array_equal(bits(out["det_dist"]), bits(dd))


def test_undistort_golden_cv2(ctx, golden):
    from putslam_b200 import api
    g = golden["undistort_cv2"]
    K = g["K"]
    cam = api.make_camera(K[0, 0], K[1, 1], K[0, 2], K[1, 2], tuple(float(x) for x in g["dist"]))
    out = ctx.backproject(g["uv"], np.full((480, 640), 5000, np.uint16), cam=cam, undistort=True)
    assert np.array_equal(bits(out["uv_undist"]), bits(g["uv_undist"]))


def test_backproject_cov_and_edges(ctx, O):
    from putslam_b200 import api, synth
    fp = synth.frame_pair(n=300, seed=6)
    cp = api.CovParams(synth.FX, synth.FY, synth.CX, synth.CY, synth.VAR_U, synth.VAR_V,
                       (api.C.c_double * 4)(*synth.DIST_VAR_COEFS))
    uv = fp["uv1"].copy()
    uv[0] = [-3.5, 10.2]; uv[1] = [639.4, 479.4]; uv[2] = [0.49, 0.51]   # border clamps (no OOB read)
    out = ctx.backproject(uv, fp["depth1"], cov=cp)
    xyz, _ = O.backproject(uv, fp["depth1"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DEPTH_SCALE)
    assert np.array_equal(bits(out["xyz"]), bits(xyz))
    for i in range(0, 300, 7):
        if uv[i, 0] < 0:
            continue
        ref = O.compute_cov(int(uv[i, 0]), int(uv[i, 1]), float(xyz[i, 2]), synth.FX, synth.FY, synth.CX, synth.CY,
                            synth.VAR_U, synth.VAR_V, synth.DIST_VAR_COEFS)
        assert np.allclose(out["cov"][i], ref, rtol=1e-12, atol=0), i   # float64 tolerance 1e-12 relative
    assert ctx.backproject(np.zeros((0, 2), np.float32), fp["depth1"])["xyz"].shape == (0, 3)


# ---------------------------------------------------------------- stage 2: guided map matching
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("M,N,seed", [(5000, 1000, 0), (5000, 1000, 1), (800, 333, 2), (40, 7, 3)])
def test_guided_match_vs_oracle(ctx, O, M, N, seed, mode):
    from putslam_b200 import host, synth
    mf = synth.map_frame(M=M, N=N, n_reobs=int(0.7 * N), seed=seed)
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    for cn in (1, 4):   # retry ladder widens the gates (matcher.cpp:617-622)
        radius, ratio = host.retry_gates(0.12, 0.55, cn)
        out = ctx.match_guided_xyz(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, radius, ratio, mode)
        q, t, d, perfect = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, radius, ratio, mode)
        assert out["total"] == q.size and out["perfect"] == perfect
        assert np.array_equal(out["q"], q) and np.array_equal(out["t"], t) and np.array_equal(out["d"], d)


def test_guided_match_many_candidates(ctx, O):
    """More gated candidates per map feature than the kernel's per-feature cache: the recompute path."""
    from putslam_b200 import host, synth
    mf = synth.map_frame(M=700, N=400, n_reobs=300, seed=12)
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    for mode in (0, 1):
        for radius, ratio in ((10.0, 0.1), (1.0, 0.9), (10.0, 1.0)):
            out = ctx.match_guided_xyz(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, radius, ratio,
                                       mode, cap=400000)
            q, t, d, perfect = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, radius, ratio, mode)
            assert out["total"] == q.size and out["perfect"] == perfect
            assert np.array_equal(out["q"], q) and np.array_equal(out["t"], t) and np.array_equal(out["d"], d)
    assert q.size > 700


def test_guided_match_depth_bins_adversarial(ctx, O):
    """The matcher sorts the current keypoints into depth bins (0.125 m, z in [0, 8)) and visits only the bins a
    feature's sphere reaches: coordinates outside the binned range, NaN / infinite depths, every keypoint in one bin,
    features exactly on bin edges, radii from one bin to the whole range and a NaN radius must all give the
    reference's answer."""
    rng = np.random.default_rng(77)
    M, N = 900, 700
    zs = np.concatenate([rng.uniform(-1.0, 0.2, 80), rng.uniform(7.5, 12.0, 80), np.full(150, 2.0),
                         np.arange(60) * 0.125, rng.uniform(0.5, 6.0, N - 370)])
    cur = np.stack([rng.uniform(-1, 1, N), rng.uniform(-1, 1, N), zs], 1).astype(np.float32)
    cur[5, 2] = np.nan; cur[6, 2] = np.inf; cur[7, 2] = -np.inf; cur[8, 0] = np.nan
    src = rng.integers(0, N, M)
    mxyz = cur[src].astype(np.float64) + rng.normal(0, 0.03, (M, 3))
    mxyz[::7] = cur[src[::7]]                                   # exact coincidences, incl. the bin-edge depths
    mxyz[3] = [0.0, 0.0, np.nan]; mxyz[4] = [0.0, 0.0, 1e9]; mxyz[9] = [0.0, 0.0, -1e9]
    cur_desc = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    map_desc = cur_desc[src] ^ (rng.random((M, 32)) < 0.02).astype(np.uint8)
    ml = rng.integers(0, 8, M).astype(np.int32); cl = np.clip(ml[rng.integers(0, M, N)] + rng.integers(-1, 2, N), 0, 7).astype(np.int32)
    n_total = 0
    with np.errstate(invalid="ignore"):
        for mode in (0, 1):
            for radius, ratio in ((0.05, 0.55), (0.12, 0.55), (0.3, 0.3), (1.5, 0.8), (20.0, 0.9), (float("nan"), 0.5)):
                out = ctx.match_guided_xyz(mxyz, map_desc, ml, cur, cur_desc, cl, radius, ratio, mode, cap=700000)
                q, t, d, perfect = O.guided_match(mxyz, map_desc, ml, cur, cur_desc, cl, radius, ratio, mode)
                assert out["total"] == q.size and out["perfect"] == perfect, (mode, radius)
                assert np.array_equal(out["q"], q) and np.array_equal(out["t"], t) and np.array_equal(out["d"], d), (mode, radius)
                n_total += q.size
    assert n_total > 5000


def test_guided_match_capacity_and_empty(ctx, O):
    from putslam_b200 import host, synth
    mf = synth.map_frame(M=600, N=300, n_reobs=200, seed=9)
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    full = ctx.match_guided_xyz(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0)
    cut = ctx.match_guided_xyz(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0, cap=10)
    assert cut["truncated"] and cut["total"] == full["total"] and np.array_equal(cut["q"], full["q"][:10])
    none = ctx.match_guided_xyz(mf["map_xyz"] + 100.0, mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0)
    assert none["total"] == 0


# ---------------------------------------------------------------- stage 3: RANSAC / Umeyama / Kabsch
@pytest.mark.parametrize("ev", [0, 4, 1, 2])
@pytest.mark.parametrize("num_hyp", [0, 256, 4096])
def test_ransac_vs_oracle(ctx, O, ev, num_hyp):
    from putslam_b200 import api, synth
    for seed in range(3):
        mc = synth.matched_clouds(m=1000, inlier_frac=0.55, seed=seed)
        r = ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=api.default_ransac_params(ev),
                                seed=100 + seed, num_hyp=num_hyp, want_counts=True)
        o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=O.default_ransac_params(ev), seed=100 + seed,
                     num_hyp=num_hyp, want_counts=True)
        n = o["hyp_used"]
        if num_hyp:  # every hypothesis is evaluated in fixed-H mode: all counts must agree
            assert np.array_equal(r["counts"][:num_hyp], o["counts"][:num_hyp])
        else:
            assert np.array_equal(r["counts"][:n], o["counts"][:n])
        assert r["hyp_used"] == o["hyp_used"]
        assert r["best_ratio"] == o["best_ratio"]
        assert np.array_equal(r["inliers"], o["inliers"])
        assert_pose_close(r["T"], o["T"].astype(np.float64))
        assert len(o["inliers"]) > 300 and np.abs(r["T"] - mc["T_gt"]).max() < 0.02


def test_ransac_models_bit_exact(ctx, O):
    """Hypothesis models are reproducible bit for bit: identical counts over 65536 hypotheses."""
    from putslam_b200 import synth
    mc = synth.matched_clouds(m=500, inlier_frac=0.4, seed=11)
    r = ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=5, num_hyp=65536, want_counts=True)
    o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=5, num_hyp=65536, want_counts=True)
    assert np.array_equal(r["counts"], o["counts"])
    assert np.array_equal(r["inliers"], o["inliers"]) and np.array_equal(bits(r["T"]), bits(o["T"]))


def test_ransac_failure_conventions(ctx, O):
    from putslam_b200 import api, synth
    mc = synth.matched_clouds(m=600, inlier_frac=0.6, seed=2)
    r = ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"][:10], mc["mt"][:10], seed=1)
    assert np.array_equal(r["T"], np.eye(4)) and r["inliers"].size == 0
    rng = np.random.default_rng(0)
    prev = rng.uniform(0.5, 4, (300, 3)).astype(np.float32); cur = rng.uniform(0.5, 4, (300, 3)).astype(np.float32)
    r = ctx.ransac_estimate(prev, cur, np.arange(300), np.arange(300), seed=1)
    o = O.ransac(prev, cur, np.arange(300), np.arange(300), seed=1)
    assert np.array_equal(r["T"], np.eye(4)) and r["inliers"].size == 0 and r["best_ratio"] == o["best_ratio"]
    prev2 = mc["prev"].copy(); prev2[::7, 2] = np.nan; prev2[1::7, 2] = 7.0; prev2[2::7, 2] = 0.05
    r = ctx.ransac_estimate(prev2, mc["cur"], mc["mq"], mc["mt"], seed=3)
    o = O.ransac(prev2, mc["cur"], mc["mq"], mc["mt"], seed=3)
    assert np.array_equal(r["inliers"], o["inliers"]) and r["hyp_used"] == o["hyp_used"]
    # filtered count falls below minimalNumberOfMatches on the device
    prev3 = mc["prev"].copy(); prev3[:, 2] = 9.0
    r = ctx.ransac_estimate(prev3, mc["cur"], mc["mq"], mc["mt"], seed=3)
    assert np.array_equal(r["T"], np.eye(4)) and r["inliers"].size == 0
    with pytest.raises(api.PslamError):
        ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=api.default_ransac_params(3))


def test_kabsch_batch_vs_oracle(ctx, O):
    from putslam_b200 import synth
    rng = np.random.default_rng(4)
    As, Bs = [], []
    for n in (100, 4, 1000, 0, 57):
        A = rng.uniform(-1.5, 1.5, (n, 3))
        R = synth.rot_from_rotvec(rng.standard_normal(3) * 0.4)
        B = A @ R.T + np.array([0.1, 0.2, -0.3]) + rng.normal(0, [0.01, 0.02, 0.03], (n, 3))
        As.append(A); Bs.append(B)
    Ts = ctx.kabsch_batch(As, Bs)
    for A, B, T in zip(As, Bs, Ts):
        ref = O.kabsch(A, B)
        assert np.array_equal(bits(T), bits(ref))     # same operation order -> same bits


# ---------------------------------------------------------------- fused pipelines
def test_frame_to_frame_pipeline(ctx, O):
    from putslam_b200 import api, synth
    for seed, n, distorted in [(0, 500, False), (1, 500, True), (2, 1000, False)]:
        fp = synth.frame_pair(n=n, seed=seed, distorted=distorted)
        cam = api.make_camera()
        f1 = ctx.frame_to_frame(None, None, fp["desc1"], fp["uv1"], fp["depth1"], cam=cam, undistort=distorted)
        out = ctx.frame_to_frame(fp["desc1"], f1["xyz"], fp["desc2"], fp["uv2"], fp["depth2"], cam=cam,
                                 undistort=distorted, seed=seed + 50)
        uv1, uv2 = fp["uv1"], fp["uv2"]
        if distorted:
            uv1 = O.undistort(uv1, synth.FX, synth.FY, synth.CX, synth.CY, synth.DIST)
            uv2 = O.undistort(uv2, synth.FX, synth.FY, synth.CX, synth.CY, synth.DIST)
        x1, _ = O.backproject(uv1, fp["depth1"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DEPTH_SCALE)
        x2, dd2 = O.backproject(uv2, fp["depth2"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DEPTH_SCALE)
        oq, ot, od = O.bf_mutual(fp["desc1"], fp["desc2"])
        ref = O.ransac(x1, x2, oq, ot, seed=seed + 50)
        assert np.array_equal(bits(f1["xyz"]), bits(x1)) and np.array_equal(bits(out["xyz"]), bits(x2))
        assert np.array_equal(bits(out["det_dist"]), bits(dd2))
        assert np.array_equal(out["mq"], oq) and np.array_equal(out["mt"], ot) and np.array_equal(out["md"], od)
        assert np.array_equal(out["inliers"], ref["inliers"]) and out["hyp_used"] == ref["hyp_used"]
        assert_pose_close(out["T"], ref["T"].astype(np.float64))
        assert out["inlier_ratio"] == O.point_inlier_ratio(ot[ref["inliers"]], ot, n)
        # the planted motion is recovered
        assert np.abs(out["T"] - fp["T_gt"]).max() < 0.01


@pytest.mark.parametrize("mode", [0, 1])
def test_frame_to_map_pipeline(ctx, O, mode):
    from putslam_b200 import host, synth
    for seed in range(2):
        mf = synth.map_frame(M=5000, N=1000, seed=seed)
        ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
        cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
        for num_hyp in (0, 4096):
            out = ctx.frame_to_map(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55,
                                   mode, seed=7, num_hyp=num_hyp)
            q, t, d, _ = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, mode)
            ref = O.ransac(mf["map_xyz"].astype(np.float32), mf["cur_xyz"], q, t, seed=7, num_hyp=num_hyp)
            assert np.array_equal(out["mq"], q) and np.array_equal(out["mt"], t) and np.array_equal(out["md"], d)
            assert np.array_equal(out["inliers"], ref["inliers"]) and out["hyp_used"] == ref["hyp_used"]
            assert_pose_close(out["T"], ref["T"].astype(np.float64))
            assert out["inlier_ratio"] == O.point_inlier_ratio(t[ref["inliers"]], t, 1000)
            assert len(ref["inliers"]) > 300 and np.abs(out["T"] - mf["T_gt"]).max() < 0.01
        # resident replay leaves the answer unchanged
        ctx.frame_to_map_resident(); ctx.sync()


def test_information_matrices(ctx, O):
    """DepthSensorModel::informationMatrixFromImageCoordinates batch (float64, tolerance 1e-12 relative vs the oracle,
    1e-9 vs numpy's LU inverse)."""
    from putslam_b200 import api, synth
    rng = np.random.default_rng(8)
    uvz = np.stack([rng.uniform(0, 639, 500), rng.uniform(0, 479, 500), rng.uniform(0.8, 6.0, 500)], 1)
    cp = api.CovParams(synth.FX, synth.FY, synth.CX, synth.CY, synth.VAR_U, synth.VAR_V,
                       (api.C.c_double * 4)(*synth.DIST_VAR_COEFS))
    info, cov = ctx.information_matrices(uvz, cp)
    for i in range(0, 500, 11):
        c, f = O.information_matrix(uvz[i, 0], uvz[i, 1], uvz[i, 2], synth.FX, synth.FY, synth.CX, synth.CY, synth.VAR_U,
                                    synth.VAR_V, synth.DIST_VAR_COEFS)
        assert np.allclose(cov[i], c, rtol=1e-12, atol=0) and np.allclose(info[i], f, rtol=1e-11, atol=0)
        assert np.allclose(info[i] @ cov[i], np.eye(3), atol=1e-9)
    assert ctx.information_matrices(np.zeros((0, 3)), cp)[0].shape == (0, 3, 3)


def test_loop_closure_pair(ctx, O):
    """Matcher::matchFeatureLoopClosure core: performMatching + RANSAC (map error version) in one submission."""
    from putslam_b200 import synth
    for seed, n in ((3, 400), (4, 35)):
        fp = synth.frame_pair(n=n, seed=seed)
        x1, _ = O.backproject(fp["uv1"], fp["depth1"], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
        x2, _ = O.backproject(fp["uv2"], fp["depth2"], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
        out = ctx.loop_closure_pair(fp["desc1"], x1, fp["desc2"], x2, seed=9)
        oq, ot, od = O.bf_mutual(fp["desc1"], fp["desc2"])
        ref = O.ransac(x1, x2, oq, ot, seed=9)
        assert np.array_equal(out["mq"], oq) and np.array_equal(out["mt"], ot) and np.array_equal(out["md"], od)
        assert np.array_equal(out["inliers"], ref["inliers"]) and np.abs(out["T"] - ref["T"]).max() <= 1e-5
        exp = O.point_inlier_ratio(ot[ref["inliers"]], ot, n)
        assert out["inlier_ratio"] == exp or (np.isnan(out["inlier_ratio"]) and np.isnan(exp))
    small = ctx.loop_closure_pair(fp["desc1"][:9], x1[:9], fp["desc2"], x2)      # fewer than 10 features -> 0
    assert small["inlier_ratio"] == 0 and small["mq"].size == 0


def test_normal_uncertainty(ctx, O):
    """Uncertainty model 1: surface normal from the 8 depth neighbours + covariance aligned with it (float64;
    1e-12 relative vs the oracle; NaN where the reference yields NaN)."""
    from putslam_b200 import synth
    rng = np.random.default_rng(10)
    # a tilted plane z = 1.5 + 0.001*u + 0.0005*v (metres), with holes and a depth edge
    uu, vv = np.meshgrid(np.arange(640), np.arange(480))
    z = 1.5 + 0.001 * uu + 0.0005 * vv
    z[:, 400:] += 0.7
    depth = np.rint(z * 5000).astype(np.uint16)
    depth[rng.random(depth.shape) < 0.05] = 0
    depth[100:110, 100:110] = 0
    px = np.stack([rng.integers(0, 640, 400), rng.integers(0, 480, 400)], 1).astype(np.int32)
    px[:4] = [[0, 0], [639, 479], [105, 105], [399, 200]]
    nrm, cov = ctx.normal_uncertainty(px, depth, scale=0.8)
    n_nan = 0
    for i in range(400):
        rn = O.compute_normal(depth, px[i, 0], px[i, 1], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
        rc = O.uncertainty_from_normal(rn, 0.8)
        if np.isnan(rn).any():
            assert np.isnan(nrm[i]).any(); n_nan += 1
            continue
        assert np.allclose(nrm[i], rn, rtol=1e-12, atol=1e-15), i
        assert np.allclose(cov[i], rc, rtol=1e-10, atol=1e-14), i
    assert n_nan >= 1 and n_nan < 100
    # on the clean plane the normal is the plane normal (up to sign) and cov shrinks along it by 0.8^2
    good = np.nonzero((px[:, 0] > 5) & (px[:, 0] < 390) & ~np.isnan(nrm).any(1))[0]
    assert abs(np.abs(nrm[good[0]] @ nrm[good[1]]) - 1) < 1e-2


def test_gradient_uncertainty(ctx, O):
    """Uncertainty model 2: colour-gradient direction (uint16 reads of the 8-bit patch, like the reference) ->
    3-D gradient -> covariance and information matrix.  Directions are integer decisions (bit-exact incl. the
    libm-sensitive diagonals); float64 results within 1e-12 relative of the oracle."""
    from putslam_b200 import synth
    rng = np.random.default_rng(12)
    H, W = 480, 640
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    rgb[200:260] = rng.integers(0, 2, (60, W, 3), dtype=np.uint8) * 255       # saturated block: many equal sums
    uu, vv = np.meshgrid(np.arange(W), np.arange(H))
    depth = np.rint((1.5 + 0.001 * uu + 0.0005 * vv) * 5000).astype(np.uint16)
    depth[rng.random(depth.shape) < 0.25] = 0                                    # every fallback branch of :171-184
    n = 600
    px = np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], 1).astype(np.int32)
    px[:6] = [[0, 0], [1, 7], [W - 1, 9], [9, H - 1], [2, 2], [W - 2, H - 2]]
    # planted diagonals |gx| == |gy| in all four quadrants and three magnitudes
    k = 6
    for mag in (1, 300, 65535):
        for (r, c) in [(2, 2), (0, 2), (2, 0), (0, 0)]:
            u, v = 20 + 7 * k, 300 + k
            rgb[v - 1:v + 2, u - 1:u + 3] = 0
            base = 3 * (u - 1) + 2 * c
            rgb[v - 1 + r].reshape(-1)[base:base + 2] = [mag & 255, mag >> 8]
            px[k] = [u, v]; k += 1
    g, cov, info = ctx.gradient_uncertainty(px, rgb, depth, scale=0.7)
    n_border = n_nan = 0
    for i in range(n):
        rg = O.compute_rgb_gradient(rgb, depth, px[i, 0], px[i, 1], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
        if np.isnan(rg).any():
            assert np.isnan(g[i]).any(); n_nan += 1
            continue
        n_border += int((rg == 1.0).all())
        assert np.allclose(g[i], rg, rtol=1e-12, atol=1e-15), (i, px[i], g[i], rg)
        rc = O.uncertainty_from_gradient(rg, 0.7)
        if np.isnan(rc).any():
            assert np.isnan(cov[i]).any()
            continue
        assert np.allclose(cov[i], rc, rtol=1e-9, atol=1e-13), i
        assert np.allclose(info[i], O.inverse3d(rc), rtol=1e-8, atol=1e-12), i
    assert n_border >= 4 and n_nan < n // 4
    # model 1 also hands back the information matrix
    nrm, ncov, ninfo = ctx.normal_uncertainty(px[:50], depth, scale=0.8, with_info=True)
    ok = ~np.isnan(ncov).any((1, 2))
    assert ok.sum() > 5
    for i in np.nonzero(ok)[0]:
        assert np.allclose(ninfo[i], O.inverse3d(ncov[i]), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("min_ratio,inlier_frac", [(0.1, 0.55), (0.1, 0.12), (0.05, 0.08), (0.3, 0.55), (0.2, 0.03)])
def test_ransac_adaptive_bound_follows_min_ratio(ctx, O, min_ratio, inlier_frac):
    """Adaptive mode with minimalInlierRatioThreshold != 0.2: the loop starts at 487 iterations and after the first
    improvement is bounded by computeRANSACIteration(minimalInlierRatioThreshold) -- 3910 at 0.1, 31k at 0.05 -- so with
    few inliers it runs far beyond 487 hypotheses.  hyp_used (finished on the host), the counts of every hypothesis the
    loop visited, the inlier set and the pose follow the oracle."""
    from putslam_b200 import api, synth
    for seed in range(3):
        mc = synth.matched_clouds(m=600, inlier_frac=inlier_frac, seed=50 + seed)
        pa = api.default_ransac_params(0); pa.minimal_inlier_ratio_threshold = min_ratio
        po = O.default_ransac_params(0); po.minimal_inlier_ratio_threshold = min_ratio
        r = ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=pa, seed=7 + seed, num_hyp=0, want_counts=True,
                                counts_cap=40000)
        o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], params=po, seed=7 + seed, num_hyp=0, want_counts=True,
                     counts_cap=40000)
        n = o["hyp_used"]
        assert r["hyp_used"] == n, (r["hyp_used"], n)
        assert np.array_equal(r["counts"][:n], o["counts"][:n])
        assert r["best_ratio"] == o["best_ratio"] and np.array_equal(r["inliers"], o["inliers"])
        assert_pose_close(r["T"], o["T"].astype(np.float64))


def test_ransac_usac_standard_stopping(ctx, O):
    """Alternative termination rule: USAC<T>::updateStandardStopping replayed over the scored hypotheses."""
    from putslam_b200 import synth
    try:
        ctx.ransac_set_stopping(1, 0.99); O.set_stopping(1, 0.99)
        for seed in range(3):
            mc = synth.matched_clouds(m=800, inlier_frac=0.5, seed=20 + seed)
            for num_hyp in (0, 2048):
                r = ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=seed, num_hyp=num_hyp)
                o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=seed, num_hyp=num_hyp)
                assert r["hyp_used"] == o["hyp_used"] and r["hyp_used"] < 200    # stops long before the budget
                assert np.array_equal(r["inliers"], o["inliers"]) and np.abs(r["T"] - o["T"]).max() <= 1e-5
    finally:
        ctx.ransac_set_stopping(0, 0.99); O.set_stopping(0, 0.99)


def test_map_prepare(ctx, O):
    """Map-side preparation: view-angle filter, global->local, projection, far-feature removal.
    Kept indices identical; float64 positions within 1e-12 relative, angles within 1e-6 rad (acos of a float ratio)."""
    from putslam_b200 import api, synth
    rng = np.random.default_rng(17)
    M = 5000
    pose = np.eye(4); pose[:3, :3] = synth.rot_from_rotvec([0.1, 0.7, -0.2]); pose[:3, 3] = [1.0, -0.5, 2.0]
    local = np.stack([rng.uniform(-3, 3, M), rng.uniform(-2, 2, M), rng.uniform(0.3, 7.0, M)], 1)
    glob = local @ pose[:3, :3].T + pose[:3, 3]
    axes = np.stack([synth.rot_from_rotvec(rng.normal(0, 0.5, 3))[:, 2] for _ in range(M)]) @ pose[:3, :3].T
    axes = (axes * rng.uniform(0.9, 1.1, (M, 1))).astype(np.float32)
    axes[5] = np.nan
    prm = api.MapPrepareParams(synth.FX, synth.FY, synth.CX, synth.CY, 640, 480, 0.6, 5.0)
    kept, xl, uv, ang = ctx.map_prepare(glob, axes, pose, prm)
    ek, exl, euv, eang = O.map_prepare(glob, axes, pose, synth.FX, synth.FY, synth.CX, synth.CY, 640, 480, 0.6, 5.0)
    assert np.array_equal(kept, ek) and 500 < kept.size < M and 5 not in kept
    assert np.allclose(xl, exl, rtol=1e-12, atol=1e-13) and np.allclose(uv, euv, rtol=1e-12, atol=1e-10)
    assert np.abs(ang - eang).max() < 1e-6
    assert np.abs(xl - local[kept]).max() < 1e-9 and (xl[:, 2] <= 5.0).all() and (ang <= 0.6).all()
    inside = uv[:, 0] >= 0
    assert inside.any() and (~inside).any() and (xl[inside, 2] >= 0.8).all()
    assert ctx.map_prepare(np.zeros((0, 3)), np.zeros((0, 3), np.float32), pose, prm)[0].size == 0


def test_frame_to_map_device_levels(ctx, O):
    """pslam_frame_to_map_features (levels predicted on the device) == pslam_frame_to_map with host-libm levels,
    including current keypoints whose detDist equals their current distance exactly (the ceil() boundary case)."""
    from putslam_b200 import host, synth
    n_exact = 0
    for seed in range(4):
        mf = synth.map_frame(M=3000, N=800, n_reobs=500, seed=40 + seed)
        cur_det = mf["cur_detdist"].copy()
        x, y, z = mf["cur_xyz"][:, 0], mf["cur_xyz"][:, 1], mf["cur_xyz"][:, 2]
        eig = np.sqrt((x * x + (y * y + z * z)).astype(np.float32)).astype(np.float64)
        n_exact += int((cur_det == eig).sum())
        cur_det[::5] *= 1.3                      # some keypoints described at another distance
        ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
        cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], cur_det)
        a = ctx.frame_to_map(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0, seed=seed)
        b = ctx.frame_to_map_features(mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], mf["cur_xyz"],
                                      mf["cur_desc"], mf["cur_octave"], cur_det, 0.12, 0.55, 0, seed=seed)
        for k in ("mq", "mt", "md", "inliers"):
            assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["T"], b["T"]) and a["mq"].size > 300
    assert n_exact > 100      # the exact-ratio case was really exercised


def test_two_contexts_run_concurrently(O):
    """SURVEY 8b threading: the tracking Matcher and the loop-closure Matcher are separate instances driven from two
    threads; here two pslam_ctx (own stream, own arenas) work concurrently and both stay bit-exact."""
    import threading
    from putslam_b200 import api, host, synth
    mf = synth.map_frame(M=3000, N=800, n_reobs=500, seed=5)
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    q, t, d, _ = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0)
    ref_a = O.ransac(mf["map_xyz"].astype(np.float32), mf["cur_xyz"], q, t, seed=3)
    db = synth.keyframe_db(n_kf=150, per_kf=500, n_query=600, n_planted=6, shared=200, seed=9)
    ref_b = O.topk(O.lc_scores(db["query"], db["db"], db["kf_off"], tau=64, threads=4), 8)
    errors = []

    def tracking():
        try:
            c = api.Context(0)
            for _ in range(30):
                r = c.frame_to_map(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0, seed=3)
                assert np.array_equal(r["mq"], q) and np.array_equal(r["inliers"], ref_a["inliers"])
            c.close()
        except Exception as e:  # noqa: BLE001
            errors.append(("tracking", repr(e)))

    def loop_closure():
        try:
            c = api.Context(0)
            c.lc_append(db["db"], db["kf_off"])
            for _ in range(30):
                ids, sc = c.lc_query(db["query"], tau=64, k=8)
                assert np.array_equal(ids, ref_b[0]) and np.array_equal(sc, ref_b[1])
            c.close()
        except Exception as e:  # noqa: BLE001
            errors.append(("loop_closure", repr(e)))

    th = [threading.Thread(target=tracking), threading.Thread(target=loop_closure)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors
