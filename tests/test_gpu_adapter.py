"""GPU test (-m gpu) of the C++ adapter: PUTSLAM-shaped classes (MatcherB200::performMatching, RGBD::*,
RANSAC::estimateTransformation, matchXYZ core, KabschEst) driven by adapter/adapter_selftest the way the
reference's call sites drive them, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "adapter", "adapter_selftest")


def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


def _rd(d, name, dt):
    return np.fromfile(os.path.join(d, name), dtype=dt)


def test_adapter_end_to_end(tmp_path, O):
    from putslam_b200 import host, synth
    assert os.path.exists(EXE), "adapter_selftest not built (run __graft_entry__.build())"
    d = str(tmp_path)
    fp = synth.frame_pair(n=500, seed=21, distorted=True)
    mf = synth.map_frame(M=2000, N=600, n_reobs=400, seed=22)
    rng = np.random.default_rng(23)
    A = rng.uniform(-1.5, 1.5, (100, 3))
    B = A @ synth.rot_from_rotvec([0.2, 0.1, -0.3]).T + [0.1, 0.2, -0.3] + rng.normal(0, [0.01, 0.02, 0.03], (100, 3))
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orb_cv2.npz"))
    np.array(g["bgr_img"].shape[:2], np.int32).tofile(os.path.join(d, "orb_dims.bin"))
    for name, arr, dt in [("orb_bgr", g["bgr_img"], np.uint8), ("orb_xy", g["bgr_xy"], np.float32),
                          ("orb_octave", g["bgr_octave"], np.int32), ("orb_angle", g["bgr_angle"], np.float32)]:
        np.ascontiguousarray(arr, dt).tofile(os.path.join(d, name + ".bin"))
    import cv2
    det_rng = np.random.default_rng(61)
    det_gray = cv2.GaussianBlur(det_rng.integers(0, 256, (480, 640), dtype=np.uint8), (0, 0), 1.5)
    det_rgb = np.stack([det_gray, np.roll(det_gray, 3, 1), 255 - np.roll(det_gray, 2, 0)], 2).copy()
    np.array(det_rgb.shape[:2], np.int32).tofile(os.path.join(d, "det_dims.bin"))
    det_rgb.tofile(os.path.join(d, "det_rgb.bin"))
    for name, arr, dt in [("desc1", fp["desc1"], np.uint8), ("desc2", fp["desc2"], np.uint8), ("uv1", fp["uv1"], np.float32),
                          ("uv2", fp["uv2"], np.float32), ("depth1", fp["depth1"], np.uint16), ("depth2", fp["depth2"], np.uint16),
                          ("map_xyz", mf["map_xyz"], np.float64), ("map_desc", mf["map_desc"], np.uint8),
                          ("map_octave", mf["map_octave"], np.int32), ("map_detdist", mf["map_detdist"], np.float64),
                          ("cur_xyz", mf["cur_xyz"], np.float32), ("cur_desc", mf["cur_desc"], np.uint8),
                          ("cur_octave", mf["cur_octave"], np.int32), ("cur_detdist", mf["cur_detdist"], np.float64),
                          ("kabsch_A", A, np.float64), ("kabsch_B", B, np.float64)]:
        np.ascontiguousarray(arr, dt).tofile(os.path.join(d, name + ".bin"))
    out = subprocess.run([EXE, d], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr

    # ---- VO path ----
    oq, ot, od = O.bf_mutual(fp["desc1"], fp["desc2"])
    assert np.array_equal(_rd(d, "vo_matches_q.bin", np.int32), oq)
    assert np.array_equal(_rd(d, "vo_matches_t.bin", np.int32), ot)
    assert np.array_equal(_rd(d, "vo_matches_d.bin", np.float32), od)
    assert (_rd(d, "vo_matches_img.bin", np.int32) == 0).all()          # BFMatcher sets imgIdx = 0
    u1 = O.undistort(fp["uv1"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DIST)
    u2 = O.undistort(fp["uv2"], synth.FX, synth.FY, synth.CX, synth.CY, synth.DIST)
    x1, _ = O.backproject(u1, fp["depth1"], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
    x2, _ = O.backproject(u2, fp["depth2"], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
    assert np.array_equal(bits(_rd(d, "vo_und2.bin", np.float32)), bits(u2.ravel()))
    assert np.array_equal(bits(_rd(d, "vo_xyz2.bin", np.float32)), bits(x2.ravel()))
    ref = O.ransac(x1, x2, oq, ot, seed=4242)
    assert np.array_equal(_rd(d, "vo_inliers_q.bin", np.int32), oq[ref["inliers"]])
    assert np.array_equal(_rd(d, "vo_inliers_t.bin", np.int32), ot[ref["inliers"]])
    T = _rd(d, "vo_T.bin", np.float32).reshape(4, 4).T                  # column-major Eigen layout
    assert np.abs(T - ref["T"]).max() <= 1e-5                           # 1e-5 m / 1e-5 rad (north_star)
    ratio, used, best = _rd(d, "vo_ratio.bin", np.float64)
    assert ratio == O.point_inlier_ratio(ot[ref["inliers"]], ot, 500) and int(used) == ref["hyp_used"] and best == ref["best_ratio"]

    # ---- fused VO step (MatcherB200::matchCore) gives the same answer as the three separate calls ----
    assert np.array_equal(_rd(d, "vof_matches_q.bin", np.int32), oq) and np.array_equal(_rd(d, "vof_matches_t.bin", np.int32), ot)
    assert np.array_equal(_rd(d, "vof_inliers_q.bin", np.int32), oq[ref["inliers"]])
    assert np.array_equal(bits(_rd(d, "vof_xyz2.bin", np.float32)), bits(x2.ravel()))
    assert np.abs(_rd(d, "vof_T.bin", np.float32).reshape(4, 4).T - ref["T"]).max() <= 1e-5
    assert _rd(d, "vof_ratio.bin", np.float64)[0] == ratio

    # ---- matchXYZ path, first call and first retry (wider gates) ----
    ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
    cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
    for cn in (1, 2):
        radius, ratio_g = host.retry_gates(0.12, 0.55, cn)
        q, t, dd, _ = O.guided_match(mf["map_xyz"], mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, radius, ratio_g, 0)
        tag = f"map{cn}"
        assert np.array_equal(_rd(d, tag + "_matches_q.bin", np.int32), q)
        assert np.array_equal(_rd(d, tag + "_matches_t.bin", np.int32), t)
        assert np.array_equal(_rd(d, tag + "_matches_d.bin", np.float32), dd)
        r = O.ransac(mf["map_xyz"].astype(np.float32), mf["cur_xyz"], q, t, seed=77)
        assert np.array_equal(_rd(d, tag + "_inliers_q.bin", np.int32), q[r["inliers"]])
        Tm = _rd(d, tag + "_T.bin", np.float32).reshape(4, 4).T
        assert np.abs(Tm - r["T"]).max() <= 1e-5
        assert _rd(d, tag + "_ratio.bin", np.float64)[0] == O.point_inlier_ratio(t[r["inliers"]], t, 600)

    # ---- resident map (uploadMapFeatures x2 + matchXYZResident, identity pose) == the host-buffer call ----
    assert np.array_equal(_rd(d, "mapr_kept.bin", np.int32), np.arange(2000))
    for f in ("matches_q", "matches_t", "inliers_q", "inliers_t"):
        assert np.array_equal(_rd(d, f"mapr_{f}.bin", np.int32), _rd(d, f"map1_{f}.bin", np.int32)), f
    assert np.array_equal(bits(_rd(d, "mapr_matches_d.bin", np.float32)), bits(_rd(d, "map1_matches_d.bin", np.float32)))
    assert np.array_equal(bits(_rd(d, "mapr_T.bin", np.float32)), bits(_rd(d, "map1_T.bin", np.float32)))
    assert _rd(d, "mapr_ratio.bin", np.float64)[0] == _rd(d, "map1_ratio.bin", np.float64)[0]

    # ---- describeFeatures: the cv2 golden vectors (colour image), features reordered like cv::ORB::compute does ----
    assert np.array_equal(_rd(d, "orb_order.bin", np.int32), g["bgr_order"])
    assert np.array_equal(_rd(d, "orb_desc.bin", np.uint8).reshape(-1, 32), g["bgr_desc"])

    # ---- detectFeatures: the reference's wrapper restated with cv2 (responses are distinct, so the sorts are unambiguous) ----
    gray = cv2.cvtColor(det_rgb, cv2.COLOR_RGB2GRAY)
    for gridn in (1, 2):
        per_roi = 500 * 3 // (gridn * gridn)
        raw = []
        for k in range(gridn):
            for i in range(gridn):
                x0, y0, w, h = k * 640 // gridn, i * 480 // gridn, 640 // gridn, 480 // gridn
                kps = cv2.ORB_create().detect(np.ascontiguousarray(gray[y0:y0 + h, x0:x0 + w]))
                kps = sorted(kps, key=lambda q: -q.response)[:per_roi]
                raw += [(np.float32(q.pt[0]) + np.float32(x0), np.float32(q.pt[1]) + np.float32(y0), q.size, q.angle, q.response,
                         q.octave) for q in kps]
        raw = sorted(raw, key=lambda q: -q[4])[:500]
        assert len({q[4] for q in raw}) == len(raw)
        exp = np.array([q[:5] for q in raw], np.float32)
        got = _rd(d, f"det{gridn}_kp.bin", np.float32).reshape(-1, 5)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), gridn
        assert np.array_equal(_rd(d, f"det{gridn}_octave.bin", np.int32), [q[5] for q in raw])

    # detector == "FAST": the 500 strongest cv::FAST corners (scores are integers, so the cut falls inside a group of ties
    # whose order is std::sort's; everything above the cut must be there, sorted, and every entry must be a cv2 corner)
    fk = _rd(d, "detfast_kp.bin", np.float32).reshape(-1, 6)
    ref = {(k.pt[0], k.pt[1]): k.response for k in cv2.FastFeatureDetector_create(10, True).detect(gray)}
    assert fk.shape[0] == 500 and (np.diff(fk[:, 4]) <= 0).all()
    assert all(ref.get((r[0], r[1])) == r[4] for r in fk) and len({(r[0], r[1]) for r in fk}) == 500
    assert (fk[:, 2] == 7).all() and (fk[:, 3] == -1).all() and (fk[:, 5] == 0).all()
    cut = fk[-1, 4]
    assert {p for p, v in ref.items() if v > cut} <= {(r[0], r[1]) for r in fk}

    # ---- Kabsch ----
    Tk = _rd(d, "kabsch_T.bin", np.float64).reshape(4, 4).T
    assert np.array_equal(bits(Tk[:3]), bits(O.kabsch(A, B)))
    assert np.array_equal(Tk[3], [0, 0, 0, 1])
    assert np.array_equal(_rd(d, "kabsch_T_empty.bin", np.float64).reshape(4, 4).T, np.eye(4))


def test_adapter_perform_tracking(tmp_path, golden):
    """MatcherB200::performTracking (reference MatcherOpenCV::performTracking, src/Matcher/matcherOpenCV.cpp:209-300) driven
    like Matcher::trackKLT drives it over a three-frame sequence -- second call on the resident previous frame, third with
    useInitialFlow + minimum-eigenvalue error and another window -- against the oracle: matches, compacted features,
    key points (attributes carried over, positions replaced) and detDists."""
    import cv2
    from oracle import klt_oracle as K
    assert os.path.exists(EXE), "adapter_selftest not built (run __graft_entry__.build())"
    d = str(tmp_path)
    g = golden["klt_cv2"]
    f0, f1 = g["a"], g["b"]
    H, W = f0.shape[:2]
    f2 = cv2.warpAffine(f1, np.float32([[1, 0, 1.6], [0, 1, 0.9]]), (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    pts = g["pts"].copy()
    pts = np.concatenate([pts, pts[:30] + np.float32(0.75), pts[:10]])          # near-duplicates and exact duplicates
    np.array([H, W, 3], np.int32).tofile(os.path.join(d, "klt_dims.bin"))
    for name, arr in (("klt_f0", f0), ("klt_f1", f1), ("klt_f2", f2)):
        np.ascontiguousarray(arr, np.uint8).tofile(os.path.join(d, name + ".bin"))
    pts.tofile(os.path.join(d, "klt_xy.bin"))
    # inputs of the fused frame (trackKLTCore): depth of frame 1, 3-D points of frame 0, intrinsics of a 200 x 150 camera
    from putslam_b200 import synth
    O = oracle_mod()
    drng = np.random.default_rng(3)
    fx, fy, cx, cy = 160.0, 160.0, 99.5, 74.5
    depth0 = (9000 + drng.integers(-3, 4, (H, W))).astype(np.uint16)
    depth1 = (9000 + drng.integers(-3, 4, (H, W))).astype(np.uint16)
    prev_xyz, _ = O.backproject(O.undistort(pts, fx, fy, cx, cy, synth.DIST), depth0, fx, fy, cx, cy, 5000.0)
    depth1.tofile(os.path.join(d, "klt_depth1.bin")); prev_xyz.tofile(os.path.join(d, "klt_prev_xyz.bin"))
    np.array([fx, fy, cx, cy], np.float32).tofile(os.path.join(d, "klt_cam.bin"))
    out = subprocess.run([EXE, d, "klt"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr

    def check(tag, nxt, kept, ids):
        assert np.array_equal(_rd(d, tag + "_q.bin", np.int32), kept)
        assert np.array_equal(_rd(d, tag + "_t.bin", np.int32), np.arange(len(kept)))
        assert (_rd(d, tag + "_d.bin", np.float32) == 0).all() and (_rd(d, tag + "_img.bin", np.int32) == -1).all()
        xy = _rd(d, tag + "_xy.bin", np.float32).reshape(-1, 4)
        assert np.array_equal(bits(xy[:, :2]), bits(nxt[kept])) and np.array_equal(bits(xy[:, 2:]), bits(nxt[kept]))
        assert np.array_equal(_rd(d, tag + "_id.bin", np.int32), ids)
        assert np.array_equal(_rd(d, tag + "_oct.bin", np.int32), ids % 5)
        assert np.array_equal(_rd(d, tag + "_det.bin", np.float64), 0.25 * ids)

    n01, s01, e01 = K.lk_pyr(f0, f1, pts, min_eig_thr=0.0)
    k01 = K.perform_tracking(e01, s01, n01, 25.0, 3.0)
    assert 40 < len(k01) < len(pts)
    check("klt01", n01, k01, k01)
    p1 = n01[k01]
    n12, s12, e12 = K.lk_pyr(f1, f2, p1, min_eig_thr=0.0)
    k12 = K.perform_tracking(e12, s12, n12, 25.0, 3.0)
    assert 30 < len(k12) <= len(k01)
    check("klt12", n12, k12, k01[k12])
    n02, s02, e02 = K.lk_pyr(f0, f2, pts, win=9, max_level=2, min_eig_thr=1e-3, init=pts, min_eig_err=True)
    k02 = K.perform_tracking(e02, s02, n02, 1e9, 1.5)
    assert 30 < len(k02) < len(pts)
    check("klt02", n02, k02, k02)
    # fused frame 0 -> 1: the same survivors, then undistort / back-project / RANSAC as the oracle composes them
    check("kltf", n01, k01, k01)
    und1 = O.undistort(n01[k01], fx, fy, cx, cy, synth.DIST)
    xyz1, _ = O.backproject(und1, depth1, fx, fy, cx, cy, 5000.0)
    assert np.array_equal(bits(_rd(d, "kltf_und.bin", np.float32)), bits(np.ascontiguousarray(und1, np.float32).ravel()))
    assert np.array_equal(bits(_rd(d, "kltf_xyz.bin", np.float32)), bits(xyz1.ravel()))
    ref = O.ransac(prev_xyz, xyz1, k01.astype(np.int32), np.arange(len(k01), dtype=np.int32), seed=99)
    assert np.array_equal(_rd(d, "kltf_inliers_q.bin", np.int32), k01[ref["inliers"]])
    assert np.array_equal(_rd(d, "kltf_inliers_t.bin", np.int32), ref["inliers"])
    assert np.abs(_rd(d, "kltf_T.bin", np.float32).reshape(4, 4).T - ref["T"]).max() <= 1e-5
    assert _rd(d, "kltf_ratio.bin", np.float64)[0] == len(ref["inliers"]) / len(k01)


def test_adapter_compute_uncertainty(tmp_path):
    """KabschEst::computeTransformation followed by TransformEst::computeUncertainty / computeUncertaintyG2O, the call
    sequence of demos/demoKabsch.cpp:1020-1028, against the oracle evaluated at the transformation the adapter returned"""
    from oracle import uncertainty_oracle as Uo
    from putslam_b200 import synth
    assert os.path.exists(EXE), "adapter_selftest not built (run __graft_entry__.build())"
    d = str(tmp_path)
    rng = np.random.default_rng(29)
    n = 100
    A = rng.uniform(-1.5, 1.5, (n, 3))
    B = A @ synth.rot_from_rotvec([0.2, 0.1, -0.3]).T + [0.1, 0.2, -0.3] + rng.normal(0, [0.01, 0.02, 0.03], (n, 3))
    L = rng.normal(0, 0.01, (2, n, 3, 3))
    CA = L[0] @ L[0].transpose(0, 2, 1); CB = L[1] @ L[1].transpose(0, 2, 1)
    for name, arr in (("kabsch_A", A), ("kabsch_B", B), ("unc_CA", CA), ("unc_CB", CB)):
        np.ascontiguousarray(arr, np.float64).tofile(os.path.join(d, name + ".bin"))
    out = subprocess.run([EXE, d, "unc"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    T = _rd(d, "unc_T.bin", np.float64).reshape(4, 4).T                      # column-major Mat34, B ~ R A + t
    assert np.abs(B - (A @ T[:3, :3].T + T[:3, 3])).max() < 0.2
    for mode in ("euler", "quat"):
        U = _rd(d, f"unc_{mode}.bin", np.float64).reshape(6, 6).T
        ref, _, _ = Uo.compute_uncertainty(B, A, CB, CA, T, mode)            # setA := B, setB := A (B ~ R A + t)
        assert np.abs(U - ref).max() < 1e-9 * np.abs(ref).max(), mode
