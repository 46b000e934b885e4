"""CPU tests: the C-ABI library loads, exports every symbol include/pslam_b200.h declares, fails loudly
without a GPU (no CPU fallback), and its host-only entry points agree with the oracle."""
import ctypes as C
import os
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
