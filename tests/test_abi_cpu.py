"""CPU tests: the C-ABI library loads, exports every symbol include/pslam_b200.h declares, fails loudly
without a GPU (no CPU fallback), and its host-only entry points agree with the oracle."""
import ctypes as C
import os
import subprocess
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pslam_b200.h")).read()
    return sorted(set(re.findall(r"PSLAM_API\s+[\w\s\*]+?\b(pslam_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from putslam_b200 import api
    lib = api.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    assert sorted(api.ABI_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name


def test_no_torch_types_or_cpp_in_header():
    text = open(os.path.join(ROOT, "include", "pslam_b200.h")).read()
    assert 'extern "C"' in text and "torch" not in text and "std::" not in text


def test_ctx_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from putslam_b200 import api
    with pytest.raises(api.PslamError):
        api.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing shipped may import, include, link or dlopen it."""
    pat = re.compile(r"(^|\s)(import|from)\s+oracle|oracle/|liboracle|orc_\w+\(|oracle\.py")
    for sub in ("putslam_b200", "adapter", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not pat.search(text), os.path.join(dirpath, f)


def test_host_sampler_mirrors_oracle(O):
    from putslam_b200 import api
    for m in (3, 16, 999, 5000):
        for h in (0, 1, 77, 4095):
            assert api.ransac_sample(0xDEADBEEFCAFE, h, m).tolist() == O.sample3(0xDEADBEEFCAFE, h, m).tolist()


def test_host_point_inlier_ratio(O):
    from putslam_b200 import api
    rng = np.random.default_rng(0)
    allt = rng.integers(0, 50, 200); inl = allt[rng.random(200) < 0.3]
    assert api.point_inlier_ratio(inl, allt) == O.point_inlier_ratio(inl, allt, 50)


def test_host_merge_and_shards():
    from putslam_b200 import host
    assert host.shard_keyframes(10, 4) == [0, 2, 5, 7, 10]
    ids, sc = host.merge_topk([(5, 3), (9, 7), (5, 1), (-1, -1), (9, 2)], 4)
    assert ids.tolist() == [2, 7, 1, 3] and sc.tolist() == [9, 9, 5, 5]
    assert host.retry_gates(0.12, 0.55, 3) == (0.12 + 0.04, 0.55 - 0.1)


def test_dbscan_against_the_reference_build(tmp_path):
    """putslam_b200::DBScan (adapter, host C++) against the REFERENCE's own DBScan, compiled from the reference tree into
    oracle/_ref (src/Matcher/dbscan.cpp with a three-symbol OpenCV stand-in, oracle/Makefile target `ref`): same
    survivors in the same order for random and clustered keypoint lists, several eps / minPts / per-cluster settings."""
    import ctypes as C
    import subprocess
    import numpy as np
    from oracle import oracle
    oracle.build()
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_dbscan.so")
    cli = os.path.join(ROOT, "adapter", "dbscan_cli")
    assert os.path.exists(ref_so), "oracle/_ref/libref_dbscan.so missing (make -C oracle ref needs /root/reference)"
    assert os.path.exists(cli), "adapter/dbscan_cli not built (run __graft_entry__.build())"
    ref = C.CDLL(ref_so)
    rng = np.random.default_rng(99)
    n_removed = 0
    for trial in range(40):
        n = int(rng.integers(0, 600))
        kind = trial % 4
        if kind == 0:                                   # uniform
            xy = rng.uniform(0, 640, (n, 2))
        elif kind == 1:                                 # tight clusters + background (multi-octave duplicates)
            centres = rng.uniform(20, 620, (max(1, n // 6), 2))
            xy = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 0.6, (n, 2))
        elif kind == 2:                                 # chains: points along lines with spacing around eps
            t = np.sort(rng.uniform(0, 300, n)); xy = np.stack([t, 0.3 * t + rng.normal(0, 0.4, n)], 1)
        else:                                           # exact duplicates and integer grid (distances exactly eps)
            xy = rng.integers(0, 25, (n, 2)).astype(np.float64)
        xy = np.ascontiguousarray(xy, np.float32)
        rng.shuffle(xy)
        for eps, min_pts, per in ((1.0, 2, 1), (3.0, 2, 1), (2.0, 3, 2), (10.0, 2, 1), (1.0, 1, 1)):
            kept = np.empty(max(1, n), np.int32)
            m = ref.orc_ref_dbscan(xy.ctypes.data_as(C.POINTER(C.c_float)), n, C.c_double(eps), min_pts, per,
                                   kept.ctypes.data_as(C.POINTER(C.c_int)))
            fin, fout = str(tmp_path / "xy.bin"), str(tmp_path / "kept.bin")
            xy.tofile(fin)
            subprocess.check_call([cli, fin, repr(eps), str(min_pts), str(per), fout])
            got = np.fromfile(fout, np.int32)
            assert np.array_equal(got, kept[:m]), (trial, eps, min_pts, per)
            n_removed += n - m
    assert n_removed > 1000          # the de-clustering really removed keypoints


def _tracking_cli(tmp_path, arrays, *args):
    cli = os.path.join(ROOT, "adapter", "tracking_cli")
    assert os.path.exists(cli), "adapter/tracking_cli not built (run __graft_entry__.build())"
    d = str(tmp_path)
    for name, (arr, dt) in arrays.items():
        np.ascontiguousarray(arr, dt).tofile(os.path.join(d, name + ".bin"))
    subprocess.check_call([cli, d] + [str(a) for a in args])
    return lambda name, dt: np.fromfile(os.path.join(d, name + ".bin"), dt)


def _tracked_lists(rng, n, spread):
    und = rng.uniform(0, spread, (n, 2)).astype(np.float32)
    if n > 20:
        und[5] = und[3]; und[9] = und[3] + np.float32(0.5); und[n - 1] = und[0]          # exact and near duplicates
    dist = (und + rng.normal(0, 0.3, und.shape)).astype(np.float32)
    z = rng.uniform(0.8, 5.0, n)
    xyz = np.stack([(und[:, 0] - spread / 2) / 500.0 * z, (und[:, 1] - spread / 2) / 500.0 * z, z], 1).astype(np.float32)
    octv = rng.integers(0, 8, n).astype(np.int32)
    det = rng.uniform(0.5, 6.0, n)
    return und, dist, xyz, octv, det


def test_tracking_remove_too_close_matches_the_reference_rule(tmp_path):
    """putslam_b200::tracking::removeTooCloseFeatures (grid-accelerated and brute-force) against the numpy restatement of
    Matcher::removeTooCloseFeatures (src/Matcher/matcher.cpp:886-974): same removed set, same compacted lists, matches
    erased by trainIdx without renumbering"""
    from oracle import klt_oracle as K
    rng = np.random.default_rng(41)
    for n, spread, e3, r2 in ((400, 120.0, 0.01, 3.0), (300, 60.0, 0.05, 1.0), (200, 500.0, 0.0, 0.0), (50, 30.0, 1e9, 2.0),
                              (0, 10.0, 0.01, 3.0), (120, 40.0, 0.02, float("inf"))):
        und, dist, xyz, octv, det = _tracked_lists(rng, n, spread)
        matches = np.stack([rng.integers(0, 1000, n), rng.permutation(n)], 1).astype(np.int32) if n else np.zeros((0, 2), np.int32)
        gone = K.remove_too_close(und, xyz, e3, r2)
        keep = np.setdiff1d(np.arange(n), gone)
        for brute in (0, 1):
            rd = _tracking_cli(tmp_path, {"und": (und, np.float32), "dist": (dist, np.float32), "xyz": (xyz, np.float32),
                                          "oct": (octv, np.int32), "det": (det, np.float64), "matches": (matches, np.int32)},
                               "remove", repr(e3), repr(r2), brute)
            assert np.array_equal(rd("removed", np.int32), gone), (n, brute)
            assert np.array_equal(rd("id", np.int32), keep), (n, brute)
            rows = rd("rows", np.float32).reshape(-1, 8)
            assert np.array_equal(rows[:, :2], und[keep]) and np.array_equal(rows[:, 2:4], dist[keep]) and np.array_equal(rows[:, 4:7], xyz[keep])
            assert np.array_equal(rd("matches_out", np.int32).reshape(-1, 2), matches[~np.isin(matches[:, 1], gone)]), (n, brute)
        if n >= 200 and r2 > 0:
            assert 0 < len(gone) < n


def test_tracking_merge_matches_the_reference_rule(tmp_path):
    """mergeTrackedFeatures (matcher.cpp:96-131): greedy, order-dependent -- a rejected candidate does not block later ones,
    an accepted one does"""
    from oracle import klt_oracle as K
    rng = np.random.default_rng(43)
    for n, m, spread, thr in ((150, 500, 200.0, 3.0), (0, 300, 60.0, 5.0), (100, 0, 50.0, 3.0), (80, 200, 100.0, 0.0), (60, 400, 30.0, 2.5),
                              (40, 60, 50.0, float("inf"))):
        und, dist, xyz, octv, det = _tracked_lists(rng, n, spread)
        s_und, s_dist, s_xyz, s_oct, s_det = _tracked_lists(rng, m, spread)
        if m > 10 and n > 10:
            s_und[2] = und[1]                                                          # on top of a tracked feature
            s_und[4] = s_und[3] + np.float32(0.25)                                     # next to an earlier candidate
        added = K.merge_tracked(und, s_und, thr)
        for brute in (0, 1):
            rd = _tracking_cli(tmp_path, {"und": (und, np.float32), "dist": (dist, np.float32), "xyz": (xyz, np.float32),
                                          "oct": (octv, np.int32), "det": (det, np.float64), "s_und": (s_und, np.float32),
                                          "s_dist": (s_dist, np.float32), "s_xyz": (s_xyz, np.float32), "s_oct": (s_oct, np.int32),
                                          "s_det": (s_det, np.float64)}, "merge", repr(thr), brute)
            assert np.array_equal(rd("id", np.int32), np.concatenate([np.arange(n), n + added])), (n, m, brute)
            rows = rd("rows", np.float32).reshape(-1, 8)
            assert np.array_equal(rows[n:, :2], s_und[added]) and np.array_equal(rows[n:, 4:7], s_xyz[added])
        if m >= 200 and thr > 0:
            assert 0 < len(added) < m
        if thr == float("inf"):
            assert len(added) == 0


def test_tracking_description_levels_match_the_reference_rule(tmp_path):
    """matcher.cpp:281-322: predicted pyramid level (host libm) and the stable regrouping by level"""
    from oracle import klt_oracle as K
    rng = np.random.default_rng(44)
    und, dist, xyz, octv, det = _tracked_lists(rng, 500, 300.0)
    det[:50] = np.sqrt((xyz[:50].astype(np.float64) ** 2).sum(1))                      # ratio 1: ceil(log(1.2^k)/log(1.2)) is ulp-sensitive
    det[50] = 0.0; xyz[51] = 0.0                                                       # log(0), division by zero
    lv, order = K.predict_description_levels(xyz, octv, det)
    rd = _tracking_cli(tmp_path, {"und": (und, np.float32), "dist": (dist, np.float32), "xyz": (xyz, np.float32),
                                  "oct": (octv, np.int32), "det": (det, np.float64)}, "levels")
    assert np.array_equal(rd("desc_oct", np.int32), lv)
    assert np.array_equal(rd("id", np.int32), order)
    assert len(np.unique(lv)) == 8


def test_uncertainty_strasdat_host_rule(tmp_path):
    """TransformEst::computeUncertaintyStrasdat (include/putslam/TransformEst/transformEst.h:343-356): identity with
    (t_k / mean point distance over both sets)^2 on the translation diagonal"""
    rng = np.random.default_rng(47)
    A = rng.uniform(-2, 2, (60, 3)); B = rng.uniform(-2, 2, (60, 3))
    T = np.eye(4); T[:3, 3] = [0.3, -0.2, 0.05]
    rd = _tracking_cli(tmp_path, {"A": (A, np.float64), "B": (B, np.float64), "T": (T, np.float64)}, "strasdat")
    U = rd("U", np.float64).reshape(6, 6)
    depth = (np.sqrt((A * A).sum(1)).sum() + np.sqrt((B * B).sum(1)).sum()) / (2 * len(A))
    ref = np.eye(6); ref[[0, 1, 2], [0, 1, 2]] = (T[:3, 3] / depth) ** 2
    assert np.allclose(U, ref, rtol=1e-13, atol=0)


def test_tracking_drop_undescribed_matches_the_reference_walk(tmp_path):
    """matcher.cpp:341-380: after cv::ORB::compute dropped border key points the feature lists keep the features whose
    position is still present -- a sequential two-pointer walk, so a dropped duplicate position keeps the FIRST of the two"""
    rng = np.random.default_rng(48)
    und, dist, xyz, octv, det = _tracked_lists(rng, 300, 200.0)
    dist[7] = dist[6]                                                                  # same key-point position twice
    kept = np.sort(rng.choice(300, 240, replace=False))
    kept = kept[kept != 7]                                                             # the second of the pair was dropped
    if 6 not in kept: kept = np.sort(np.append(kept, 6))
    rd = _tracking_cli(tmp_path, {"und": (und, np.float32), "dist": (dist, np.float32), "xyz": (xyz, np.float32),
                                  "oct": (octv, np.int32), "det": (det, np.float64), "desc_xy": (dist[kept], np.float32)}, "undescribed")
    # restatement of the walk
    exp, j = [], 0
    for i in range(300):
        if j == len(kept): break
        d = dist[i].astype(np.float32) - dist[kept[j]].astype(np.float32)
        if np.sqrt(float(d[0]) ** 2 + float(d[1]) ** 2) < 0.0001:
            exp.append(i); j += 1
    assert np.array_equal(rd("id", np.int32), exp) and np.array_equal(exp, kept)
    # nothing dropped: untouched
    rd = _tracking_cli(tmp_path, {"desc_xy": (dist, np.float32)}, "undescribed")
    assert np.array_equal(rd("id", np.int32), np.arange(300))
