"""CPU tests (-m "not gpu"): the drop-in boundary against the reference's own headers.

1. Token level: the declarations the adapter mirrors (adapter/pslam_adapter.h, adapter/putslam_tree/Matcher/matcherB200.h) are
   compared token for token with the reference's headers where they lie -- the four virtuals of Matcher (matcher.h:405-422),
   RANSAC::parameters / RANSAC(...) / estimateTransformation / pointInlierRatio (RANSAC.h:23-66), the RGBD free functions
   (RGBD.h:38-73), TransformEst::computeTransformation and the Kabsch factory (transformEst.h:23, kabschEst.h:16).
2. Compiler level: adapter/putslam_tree/src/matcherB200.cpp (MatcherB200 : public MatcherOpenCV with `override` on every
   virtual, plus the two factories) and adapter/pslam_adapter.cpp with -DPSLAM_USE_REAL_HEADERS are compiled against the
   reference's real matcher.h / matcherOpenCV.h and the Eigen / OpenCV API stand-ins the reference's own sources compile
   against (oracle/ref_shim) -- a signature that drifted from the reference fails to compile.
Needs /root/reference (present where the CPU suite runs); skipped otherwise."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "include", "putslam")), reason="reference tree not present")


def text(path):
    s = open(path).read()
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def tokens(decl):
    """C++ tokens of a declaration, without virtual / override / = 0 / default arguments' spacing differences"""
    decl = re.sub(r"\b(virtual|override|inline)\b", " ", decl)
    decl = re.sub(r"=\s*0\s*$", " ", decl.strip().rstrip(";").strip())
    return re.findall(r"[A-Za-z_][A-Za-z_0-9]*|::|[-+]?\d+\.?\d*|[^\sA-Za-z_0-9]", decl)


def params(decl):
    """parameter TYPES of a function declaration (names and default values dropped), normalised"""
    inner = decl[decl.index("(") + 1: decl.rindex(")")]
    out, depth, cur = [], 0, ""
    for ch in inner:
        if ch in "<(":
            depth += 1
        if ch in ">)":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    res = []
    for p in out:
        p = p.split("=")[0].strip()
        if p == "void":
            continue
        p = re.sub(r"\bputslam::", "", p)
        m = re.match(r"^(.*?)([A-Za-z_][A-Za-z_0-9]*)$", p)          # strip the parameter name (last identifier), if any
        typ = m.group(1) if m and re.search(r"[\s&*>]$", m.group(1)) else p
        res.append("".join(tokens(typ)))
    return res


def find_decl(src, name, nth=0, before=None):
    """the nth declaration `... name(...)` in src (up to the closing parenthesis of its parameter list)"""
    pos = -1
    for _ in range(nth + 1):
        pos = src.index(name + "(", pos + 1)
    start = max(src.rfind(";", 0, pos), src.rfind("}", 0, pos), src.rfind("{", 0, pos)) + 1
    for spec in ("public:", "private:", "protected:"):
        k = src.rfind(spec, 0, pos)
        if k >= 0:
            start = max(start, k + len(spec))
    depth, i = 0, src.index("(", pos)
    while True:
        depth += src[i] == "("
        depth -= src[i] == ")"
        if depth == 0:
            break
        i += 1
    return src[start:i + 1].strip()


def ret_type(decl, name):
    return "".join(tokens(re.sub(r"\bputslam::", "", decl[:decl.index(name + "(")])))


def test_matcher_virtuals_match_the_reference_header():
    ref = text(os.path.join(REF, "include/putslam/Matcher/matcher.h"))
    mine = text(os.path.join(ROOT, "adapter/putslam_tree/Matcher/matcherB200.h"))
    core = text(os.path.join(ROOT, "adapter/pslam_adapter.h"))
    for name in ("detectFeatures", "describeFeatures", "performMatching", "performTracking"):
        r = find_decl(ref, name)
        assert "virtual" in r, name                                    # these are the reference's dispatch points
        m = find_decl(mine, name)
        assert params(m) == params(r) and ret_type(m, name) == ret_type(r, name), (name, params(m), params(r))
    # the adapter core behind them takes the same image / list arguments (plus explicit parameters where the member read fields)
    for name in ("performMatching", "describeFeatures", "performTracking"):
        assert params(find_decl(core, name)) == params(find_decl(ref, name)), name


def test_ransac_interface_matches_the_reference_header():
    ref = text(os.path.join(REF, "include/putslam/TransformEst/RANSAC.h"))
    mine = text(os.path.join(ROOT, "adapter/pslam_adapter.h"))
    body = lambda s: re.search(r"struct\s+parameters\s*\{(.*?)\}\s*;", s, re.S).group(1)
    assert tokens(body(mine)) == tokens(body(ref))                      # field for field, type for type, in order
    enum = lambda s: tokens(re.search(r"enum\s+ERROR_VERSION\s*\{(.*?)\}", s, re.S).group(1))
    assert enum(mine) == enum(ref)
    r_ctor = find_decl(ref, "RANSAC", nth=0); m_ctor = find_decl(mine[mine.index("class RANSAC"):], "RANSAC", nth=0)
    assert [p.replace("RANSAC::", "") for p in params(r_ctor)] == params(m_ctor)
    for name in ("estimateTransformation", "pointInlierRatio"):
        r = find_decl(ref, name); m = find_decl(mine, name)
        assert params(m) == params(r) and ret_type(m, name).replace("static", "") == ret_type(r, name).replace("static", ""), name


def test_rgbd_and_transform_est_interfaces_match_the_reference_headers():
    ref = text(os.path.join(REF, "include/putslam/RGBD/RGBD.h"))
    mine = text(os.path.join(ROOT, "adapter/pslam_adapter.h"))
    ref_k2d = [find_decl(ref, "keypoints2Dto3D", nth=i) for i in range(2)]
    m = find_decl(mine, "keypoints2Dto3D")
    assert params(m) in [params(r) for r in ref_k2d] and ret_type(m, "keypoints2Dto3D") == ret_type(ref_k2d[1], "keypoints2Dto3D")
    ref_rid = sorted(params(find_decl(ref, "removeImageDistortion", nth=i)) for i in range(2))
    mine_rid = sorted(params(find_decl(mine, "removeImageDistortion", nth=i)) for i in range(2))
    assert ref_rid == mine_rid
    te = text(os.path.join(REF, "include/putslam/TransformEst/transformEst.h"))
    r = find_decl(te, "computeTransformation"); m = find_decl(mine, "computeTransformation")
    assert params(m) == params(r) and ret_type(m, "computeTransformation") == ret_type(r, "computeTransformation")
    ke = text(os.path.join(REF, "include/putslam/TransformEst/kabschEst.h"))
    assert ret_type(find_decl(ke, "createKabschEstimator"), "createKabschEstimator") == ret_type(find_decl(mine, "createKabschEstimator"), "createKabschEstimator")
    # factories of the tree file have the reference's signatures
    mo = text(os.path.join(REF, "include/putslam/Matcher/matcherOpenCV.h"))
    tree = text(os.path.join(ROOT, "adapter/putslam_tree/Matcher/matcherB200.h"))
    assert params(find_decl(tree, "createMatcherB200")) == params(find_decl(mo, "createMatcherOpenCV", nth=1))
    assert params(find_decl(tree, "createloopClosingMatcherB200")) == params(find_decl(mo, "createloopClosingMatcherOpenCV"))


def test_tree_files_compile_against_the_reference_headers():
    """`override` on every virtual + the real matcher.h / matcherOpenCV.h: the compiler is the conformance checker"""
    flags = ["-std=c++14", "-fsyntax-only", "-w", "-DPSLAM_USE_REAL_HEADERS", "-I", os.path.join(ROOT, "oracle/ref_shim"),
             "-I", os.path.join(REF, "include/putslam"), "-I", os.path.join(ROOT, "adapter/putslam_tree"), "-I", os.path.join(ROOT, "adapter"),
             "-I", os.path.join(ROOT, "include")]
    for src in ("adapter/putslam_tree/src/matcherB200.cpp", "adapter/pslam_adapter.cpp"):
        out = subprocess.run(["g++"] + flags + [os.path.join(ROOT, src)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, (src, out.stderr[-1500:])
    # and a drifted signature is caught: drop one reference from performTracking's parameter list
    bad = open(os.path.join(ROOT, "adapter/putslam_tree/Matcher/matcherB200.h")).read().replace(
        "std::vector<double>& prevDetDists,", "std::vector<double> prevDetDists,")
    probe = os.path.join(ROOT, "tests", "_build", "drift")
    os.makedirs(os.path.join(probe, "Matcher"), exist_ok=True)
    open(os.path.join(probe, "Matcher", "matcherB200.h"), "w").write(bad)
    open(os.path.join(probe, "probe.cpp"), "w").write('#include "Matcher/matcherB200.h"\n')
    flags2 = [f if f != os.path.join(ROOT, "adapter/putslam_tree") else probe for f in flags]
    out = subprocess.run(["g++"] + flags2 + [os.path.join(probe, "probe.cpp")], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "override" in out.stderr
