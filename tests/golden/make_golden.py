"""Generates tests/golden/*.npz from the REAL OpenCV (cv2) -- the third-party code that carries the
arithmetic of the reference's matching and undistortion calls:

  cv::BFMatcher(NORM_HAMMING, crossCheck=true).match   src/Matcher/matcherOpenCV.cpp:105,203
  cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)            (north_star ratio-test extension)
  cv::Mat a - b (saturating) + cv::norm(NORM_HAMMING)  src/Matcher/matcher.cpp:719-721
  cv::undistortPoints                                  src/RGBD/RGBD.cpp:268,298
  cv::ORB::compute (provided keypoints)                src/Matcher/matcherOpenCV.cpp:181-195
  cv::ORB::detect                                      src/Matcher/matcherOpenCV.cpp:118-176
  cv::calcOpticalFlowPyrLK                             src/Matcher/matcherOpenCV.cpp:209-236

Run in the build container (cv2 4.13.0):  python tests/golden/make_golden.py
The vectors are small on purpose; the oracle (oracle/oracle.c) and the CUDA path are both checked
against them, bit for bit.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def bf_cases():
    rng = np.random.default_rng(20261017)
    cases = {}
    shapes = [("sq500", 500, 500, 32, 256), ("q_lt_t", 200, 700, 32, 256), ("q_gt_t", 700, 200, 32, 256),
              ("ties4", 300, 300, 32, 4), ("dups", 128, 128, 32, 256), ("one", 1, 1, 32, 256),
              ("onerow", 1, 50, 32, 256)]
    for name, nq, nt, nb, alphabet in shapes:
        if alphabet == 256:
            q = rng.integers(0, 256, (nq, nb), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, nb), dtype=np.uint8)
        else:  # heavy ties: descriptors drawn from a tiny alphabet in a few bytes, zeros elsewhere
            q = np.zeros((nq, nb), np.uint8); t = np.zeros((nt, nb), np.uint8)
            q[:, :2] = rng.integers(0, alphabet, (nq, 2)); t[:, :2] = rng.integers(0, alphabet, (nt, 2))
        if name == "dups":
            t[::2] = t[0]; q[::3] = t[0]; q[5] = t[7]
        ms = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(q, t)
        cases[name + "_q"] = q; cases[name + "_t"] = t
        cases[name + "_mq"] = np.array([m.queryIdx for m in ms], np.int32)
        cases[name + "_mt"] = np.array([m.trainIdx for m in ms], np.int32)
        cases[name + "_md"] = np.array([m.distance for m in ms], np.float32)
        if nt >= 2:
            kn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
            cases[name + "_k_idx"] = np.array([[m[0].trainIdx, m[1].trainIdx] for m in kn], np.int32)
            cases[name + "_k_dist"] = np.array([[m[0].distance, m[1].distance] for m in kn], np.float32)
    cases["names"] = np.array([s[0] for s in shapes])
    return cases


def satsub_cases():
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    a[0] = b[0]; a[1] = 255; b[1] = 0; a[2] = 0; b[2] = 255
    quirk = np.array([cv2.norm(cv2.subtract(x.reshape(1, -1), y.reshape(1, -1)), cv2.NORM_HAMMING) for x, y in zip(a, b)], np.float32)
    xor = np.array([cv2.norm(x, y, cv2.NORM_HAMMING) for x, y in zip(a, b)], np.float32)
    return dict(a=a, b=b, quirk=quirk, xor=xor)


def undistort_cases():
    rng = np.random.default_rng(11)
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    d = np.array([-0.0410, 0.3286, 0.0087, 0.0051, -0.5643], np.float32)
    uv = np.stack([rng.uniform(0, 639, 400), rng.uniform(0, 479, 400)], 1).astype(np.float32)
    und = cv2.undistortPoints(uv.reshape(-1, 1, 2), K, d).reshape(-1, 2)
    # RGBD::removeImageDistortion re-projection, float32 (src/RGBD/RGBD.cpp:274-279)
    out = np.stack([und[:, 0] * K[0, 0] + K[0, 2], und[:, 1] * K[1, 1] + K[1, 2]], 1).astype(np.float32)
    return dict(K=K, dist=d, uv=uv, normalized=und.astype(np.float32), uv_undist=out)


def orb_scene(rng, H, W, colour=False):
    """deterministic test image with structure at several scales (blobs, bars, texture), uint8"""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    img = 96 + 40 * np.sin(xx / 9.0) * np.cos(yy / 13.0) + 30 * np.sin((xx + 2 * yy) / 37.0)
    for _ in range(60):
        cx, cy, r, a = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(2, 18), rng.uniform(-70, 70)
        img += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * r * r))
    img += rng.normal(0, 6, (H, W))
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    if not colour:
        return g
    return np.stack([g, np.roll(g, 3, 1), 255 - np.roll(g, 5, 0)], 2).copy()


def orb_cases():
    """cv::ORB::compute with caller-provided keypoints (MatcherOpenCV::describeFeatures,
    src/Matcher/matcherOpenCV.cpp:181-195): kept/reordered keypoint indices and descriptors from cv2."""
    rng = np.random.default_rng(31)
    orb = cv2.ORB_create()
    out = {}
    for name, H, W, colour, n in (("gray", 240, 320, False, 400), ("bgr", 150, 200, True, 150)):
        img = orb_scene(rng, H, W, colour)
        xy = np.stack([rng.uniform(20, W - 20, n), rng.uniform(20, H - 20, n)], 1).astype(np.float32)
        # rounding boundary of the border filter (30.5 -> 30, 31.5 -> 32) and exact edges
        xy[:6] = [[30.5, 100.0], [31.5, 100.0], [W - 31.5, 80.0], [W - 30.5, 80.0], [100.0, 30.5], [100.0, H - 31.5]]
        octave = rng.integers(0, 8, n).astype(np.int32)
        angle = rng.uniform(0, 360, n).astype(np.float32)
        angle[6:10] = [0.0, 90.0, 180.0, 359.99]
        kps = [cv2.KeyPoint(float(x), float(y), 31.0, float(a), 1.0, int(o)) for (x, y), a, o in zip(xy, angle, octave)]
        k2, d2 = orb.compute(img, kps)
        # identify the surviving keypoints: (x, y, octave, angle) is unique per keypoint
        key = {(float(k.pt[0]), float(k.pt[1]), k.octave, float(k.angle)): i for i, k in enumerate(kps)}
        order = np.array([key[(float(k.pt[0]), float(k.pt[1]), k.octave, float(k.angle))] for k in k2], np.int32)
        out.update({name + "_img": img, name + "_xy": xy, name + "_octave": octave, name + "_angle": angle,
                    name + "_order": order, name + "_desc": d2})
    out["names"] = np.array(["gray", "bgr"])
    return out


def orb_detect_cases():
    """cv::ORB::detect (the call inside MatcherOpenCV::detectFeatures, src/Matcher/matcherOpenCV.cpp:118-176), keypoints
    in OpenCV's own output order: x, y, size, angle, response (float32) and octave."""
    rng = np.random.default_rng(47)
    out = {}
    names = []
    for name, H, W, nf in (("d500", 240, 320, 500), ("d150", 200, 260, 150)):
        img = orb_scene(rng, H, W)
        img = np.clip(img.astype(np.int32) + rng.integers(-25, 26, img.shape), 0, 255).astype(np.uint8)   # more corners
        kps = cv2.ORB_create(nfeatures=nf).detect(img)
        out[name + "_img"] = img
        out[name + "_nfeatures"] = np.int32(nf)
        out[name + "_kp"] = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kps], np.float32)
        out[name + "_octave"] = np.array([k.octave for k in kps], np.int32)
        names.append(name)
    out["names"] = np.array(names)
    return out


def klt_cases():
    """cv::calcOpticalFlowPyrLK as MatcherOpenCV::performTracking calls it (src/Matcher/matcherOpenCV.cpp:209-236):
    colour frames, winSize 7, 3 pyramid levels above the base, 30 iterations / eps 0.01; plus a gray case and the
    two flag variants."""
    rng = np.random.default_rng(59)
    crit = (cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 30, 0.01)
    H, W = 150, 200
    g = orb_scene(rng, H, W)
    a = np.stack([g, np.roll(g, 4, 1), 255 - np.roll(g, 3, 0)], 2).copy()
    M = np.float32([[0.9985, 0.02, 2.6], [-0.02, 0.9985, -1.9]])
    b = cv2.warpAffine(a, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    b = np.clip(b.astype(np.int32) + rng.integers(-2, 3, b.shape), 0, 255).astype(np.uint8)
    n = 120
    pts = np.stack([rng.uniform(1, W - 1, n), rng.uniform(1, H - 1, n)], 1).astype(np.float32)
    pts[:4] = [[0.2, 0.3], [W - 1.2, H - 1.4], [3.5, H - 2.0], [W / 2, H / 2]]
    init = (pts + rng.normal(0, 1.2, pts.shape)).astype(np.float32)
    out = {"a": a, "b": b, "pts": pts, "init": init}
    for name, ia, ib, kw, nxt in (("colour", a, b, {}, None), ("gray", a[..., 0].copy(), b[..., 0].copy(), {}, None),
                                  ("mineig", a, b, {"flags": cv2.OPTFLOW_LK_GET_MIN_EIGENVALS}, None),
                                  ("initflow", a, b, {"flags": cv2.OPTFLOW_USE_INITIAL_FLOW}, init)):
        p1, st, er = cv2.calcOpticalFlowPyrLK(ia, ib, pts.reshape(-1, 1, 2), None if nxt is None else nxt.reshape(-1, 1, 2).copy(),
                                              winSize=(7, 7), maxLevel=3, criteria=crit, **kw)
        out[name + "_next"] = p1.reshape(-1, 2); out[name + "_status"] = st.ravel(); out[name + "_err"] = er.ravel()
    out["names"] = np.array(["colour", "gray", "mineig", "initflow"])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "bf_cv2.npz"), **bf_cases())
    np.savez_compressed(os.path.join(HERE, "satsub_cv2.npz"), **satsub_cases())
    np.savez_compressed(os.path.join(HERE, "undistort_cv2.npz"), **undistort_cases())
    np.savez_compressed(os.path.join(HERE, "orb_cv2.npz"), **orb_cases())
    np.savez_compressed(os.path.join(HERE, "orb_detect_cv2.npz"), **orb_detect_cases())
    np.savez_compressed(os.path.join(HERE, "klt_cv2.npz"), **klt_cases())
    print("cv2", cv2.__version__, "golden vectors written to", HERE)
