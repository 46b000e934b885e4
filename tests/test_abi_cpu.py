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
