"""CPU restatement (numpy) of the ORB descriptor path -- TEST INFRASTRUCTURE ONLY, never imported by the product.

Reference seam: MatcherOpenCV::describeFeatures (src/Matcher/matcherOpenCV.cpp:181-195), i.e.
`descriptorExtractor->compute(rgbImage, features, descriptors)` with `cv::ORB::create()` defaults (:83-84):
8 levels, scale 1.2, patch 31, edge threshold 31, WTA_K 2.  The arithmetic lives in OpenCV (un-vendored, un-pinned
dependency; `FIND_PACKAGE(OpenCV REQUIRED)`, CMakeLists.txt:118) -- this file restates what cv::ORB::compute does with
caller-provided keypoints in OpenCV 4.13.0 and is pinned bit for bit against that library (tests/golden/orb_cv2.npz,
tests/test_oracle_cpu.py::test_orb_*):

  1. colour input -> gray, COLOR_BGR2GRAY fixed point: (B*3735 + G*19235 + R*9798 + 2^14) >> 15
  2. keypoints: drop those whose rounded position lies outside [31, W-31) x [31, H-31) (KeyPointsFilter::
     runByImageBorder, Point2f -> Point by round-half-even); regroup by octave (stable) -- ORB::compute reorders and
     shortens the caller's keypoint vector this way, so the descriptor rows follow that order
  3. pyramid: level 0 = the image; level l = resize(level l-1, (round(W/s_l), round(H/s_l)), INTER_LINEAR_EXACT) with
     s_l = float(1.2f ** l): 8.8 fixed-point coefficients round((frac) * 256) from double-precision source positions,
     horizontal pass to 8.8, vertical pass to 16.16, round half up
  4. every level gets a 32-pixel BORDER_REFLECT_101 frame of the UNBLURRED level; the interior is then blurred in place
     with GaussianBlur(7x7, sigma 2).  Inside ORB the level is a sub-matrix of the pyramid buffer, for which OpenCV
     takes the float separable filter, not its fixed-point one: float32 kernel getGaussianKernel(7, 2), row pass
     s = k0*x0, s = fma(k_i, x_i, s) in tap order, column pass s = k3*c, s = fma(k_{3+d}, (below + above), s), then
     round-half-even and saturate (the FMA contraction is what OpenCV's AVX2 build executes; verified against
     cv2.sepFilter2D: 0 differing pixels in 1.8 M, 2-5 without the FMAs)
  5. descriptor: for test t, point (px, py) -> (round(px*a - py*b), round(px*b + py*a)) with a = float(cos(angle)),
     b = float(sin(angle)), angle = float(deg) * float(pi/180), float32 products and sums, round-half-even; bit t =
     I(p0) < I(p1) sampled around (round(x / s_l), round(y / s_l)) in the level's framed image; 8 tests per byte,
     first test in bit 0.
"""
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN = np.loadtxt(os.path.join(_HERE, "orb_pattern_31.txt"), dtype=np.int32).reshape(256, 4)
EDGE_THRESHOLD = 31
BORDER = 32            # max(edgeThreshold, ceil(15 * sqrt 2), 9 / 2) + 1
# getGaussianKernel(7, 2, CV_32F): exp(-i^2 / 8) normalised in double, stored as float32
_g = np.exp(-(np.arange(7, dtype=np.float64) - 3.0) ** 2 / 8.0)
GAUSS7 = (_g / _g.sum()).astype(np.float32)
f32 = np.float32


def bgr2gray(img):
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def level_scale(level):
    """getScale(): (float)pow((double)1.2f, level)"""
    return f32(math.pow(float(f32(1.2)), float(level)))


def level_size(W, H, level):
    s = level_scale(level)
    return int(np.rint(f32(W) / s)), int(np.rint(f32(H) / s))


def linear_exact_coeffs(src, dst):
    """interpolationLinear<uchar>::getCoeffs for every destination index: source offset and the two 8.8 weights"""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))
    fval = scale * (np.arange(dst, dtype=np.float64) + 0.5) - 0.5
    ival = np.floor(fval).astype(np.int64)
    off = np.zeros(dst, np.int64); c0 = np.full(dst, 256, np.int64); c1 = np.zeros(dst, np.int64)
    if src > 1:
        mid = (ival >= 0) & (ival < src - 1)
        off[mid] = ival[mid]
        c1[mid] = np.rint((fval[mid] - ival[mid]) * 256.0).astype(np.int64)
        c0[mid] = 256 - c1[mid]
        off[ival >= src - 1] = src - 1          # right / bottom edge: the last source sample
    return off, c0, c1


def resize_linear_exact(img, dw, dh):
    sh, sw = img.shape
    ox, x0, x1 = linear_exact_coeffs(sw, dw)
    oy, y0, y1 = linear_exact_coeffs(sh, dh)
    s = img.astype(np.int64)
    h = x0[None, :] * s[:, ox] + x1[None, :] * s[:, np.minimum(ox + 1, sw - 1)]
    v = y0[:, None] * h[oy, :] + y1[:, None] * h[np.minimum(oy + 1, sh - 1), :]
    return ((v + 32768) >> 16).astype(np.uint8)


def _fma(a, b, c):
    # float32 fused multiply-add emulated through double: the product of two float32 is exact in double, the sum is
    # rounded to double and then to float32; this differs from a true fma only by double rounding (~2^-29 per
    # operation) and is pinned empirically against cv2 (0 differing pixels in the golden and live comparisons)
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def gaussian7_submatrix(img):
    """GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) as executed on a sub-matrix (float separable filter with FMA)"""
    H, W = img.shape
    k = GAUSS7
    p = np.pad(img, 3, mode="reflect").astype(np.float32)
    s = k[0] * p[:, 0:W]
    for i in range(1, 7):
        x = p[:, i:i + W]
        s = _fma(np.full_like(x, k[i]), x, s)
    v = k[3] * s[3:3 + H, :] + f32(0)
    for d in (1, 2, 3):
        t = s[3 + d:3 + d + H, :] + s[3 - d:3 - d + H, :]
        v = _fma(np.full_like(t, k[3 + d]), t, v)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def pyramid(gray, nlevels):
    """-> list of (scale float32, framed level image: 32 px reflect-101 frame of the unblurred level, blurred interior)"""
    H, W = gray.shape
    out = []
    prev = gray
    for l in range(nlevels):
        if l > 0:
            w, h = level_size(W, H, l)
            prev = resize_linear_exact(prev, w, h)
        ext = np.pad(prev, BORDER, mode="reflect")
        ext[BORDER:-BORDER, BORDER:-BORDER] = gaussian7_submatrix(prev)
        out.append((level_scale(l), ext))
    return out


def filter_and_order(xy, octave, W, H):
    """indices of the keypoints ORB::compute keeps, in its output order"""
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    octave = np.asarray(octave, np.int64)
    ix = np.rint(xy[:, 0]).astype(np.int64); iy = np.rint(xy[:, 1]).astype(np.int64)
    keep = np.nonzero((ix >= EDGE_THRESHOLD) & (ix < W - EDGE_THRESHOLD) & (iy >= EDGE_THRESHOLD) & (iy < H - EDGE_THRESHOLD))[0]
    return keep[np.argsort(octave[keep], kind="stable")]


def rotation_terms(angle_deg):
    """a = (float)cos(angle), b = (float)sin(angle), angle = float(deg) * (float)(CV_PI / 180.f)"""
    ang = np.asarray(angle_deg, np.float32) * f32(math.pi / 180.0)
    a = np.array([math.cos(float(t)) for t in ang.ravel()], np.float32).reshape(ang.shape)
    b = np.array([math.sin(float(t)) for t in ang.ravel()], np.float32).reshape(ang.shape)
    return a, b


def describe(image, xy, octave, angle_deg):
    """cv::ORB::compute(image, keypoints) -> (order int64[n_out] into the input keypoints, descriptors uint8[n_out, 32])"""
    image = np.asarray(image)
    gray = bgr2gray(image) if image.ndim == 3 else image
    H, W = gray.shape
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    octave = np.asarray(octave, np.int64).reshape(-1)
    if xy.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros((0, 32), np.uint8)
    order = filter_and_order(xy, octave, W, H)
    levels = pyramid(gray, int(octave.max()) + 1)      # level count from ALL provided keypoints (before the filter)
    a, b = rotation_terms(np.asarray(angle_deg, np.float32).reshape(-1))
    pat = PATTERN.astype(np.float32)
    desc = np.zeros((order.size, 32), np.uint8)
    for j, i in enumerate(order):
        scale, ext = levels[octave[i]]
        inv = f32(1.0) / scale
        cx = int(np.rint(xy[i, 0] * inv)) + BORDER
        cy = int(np.rint(xy[i, 1] * inv)) + BORDER
        x0 = np.rint(pat[:, 0] * a[i] - pat[:, 1] * b[i]).astype(np.int64)
        y0 = np.rint(pat[:, 0] * b[i] + pat[:, 1] * a[i]).astype(np.int64)
        x1 = np.rint(pat[:, 2] * a[i] - pat[:, 3] * b[i]).astype(np.int64)
        y1 = np.rint(pat[:, 2] * b[i] + pat[:, 3] * a[i]).astype(np.int64)
        bits = (ext[cy + y0, cx + x0] < ext[cy + y1, cx + x1]).astype(np.uint8)
        desc[j] = np.packbits(bits, bitorder="little")
    return order, desc


# ======================================================================================================================
# Detection: what cv::ORB::create()->detect(gray) does (OpenCV 4.13 features2d, computeKeyPoints), the call inside
# MatcherOpenCV::detectFeatures (reference src/Matcher/matcherOpenCV.cpp:118-176, `featureDetector->detect(roi, kps)`).
#   per pyramid level (unblurred): FAST-9/16, threshold 20, 3x3 non-maximum suppression on the corner score
#   -> drop keypoints closer than 31 px to the level's border -> retainBest(2 n_l) by FAST score (ties with the n-th
#   kept) -> Harris response (7x7 block of Sobel-like derivatives, k = 0.04, float32) -> retainBest(n_l) by Harris
#   -> intensity-centroid angle (circular patch of radius 15, fastAtan2) -> position * level scale, size = 31 * scale.
# n_l: nfeatures split over the levels by the factor 1/1.2 (float32 arithmetic, cvRound), remainder on the last level.
# The functions below return the keypoints of a level in RASTER order; OpenCV's own order inside a level is whatever
# std::nth_element / std::partition leave behind (the product reproduces it by running the same libstdc++ algorithms
# on the host; tests compare that order with cv2 directly, and sets / values with this restatement).
# ======================================================================================================================
FAST_RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1),
             (-3, 0), (-3, 1), (-2, 2), (-1, 3)]          # (dx, dy), OpenCV's makeOffsets order for patternSize 16


def fast_score_map(img, threshold=20):
    """corner score of every pixel (0 = not a FAST-9 corner at `threshold`): the largest t for which 9 contiguous ring
    pixels are all brighter than v + t or all darker than v - t (cv::cornerScore<16>)"""
    H, W = img.shape
    out = np.zeros((H, W), np.int32)
    if H < 7 or W < 7:
        return out
    I = img.astype(np.int32)
    c = I[3:H - 3, 3:W - 3]
    d = np.stack([c - I[3 + dy:H - 3 + dy, 3 + dx:W - 3 + dx] for (dx, dy) in FAST_RING], 0)
    ext = np.concatenate([d, d[:8]], 0)
    amin = np.full_like(c, -(1 << 20)); bmax = np.full_like(c, 1 << 20)
    for k in range(16):
        seg = ext[k:k + 9]
        amin = np.maximum(amin, seg.min(0))
        bmax = np.minimum(bmax, seg.max(0))
    corner = (amin > threshold) | (bmax < -threshold)
    out[3:H - 3, 3:W - 3] = np.where(corner, np.maximum(amin, -bmax) - 1, 0)
    return out


def fast_nms(score):
    """(x, y, score) of the corners that beat all 8 neighbours strictly, raster order (cv::FAST with nonmaxSuppression)"""
    H, W = score.shape
    p = np.pad(score, 1)
    nb = np.max(np.stack([p[1 + dy:H + 1 + dy, 1 + dx:W + 1 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dx, dy) != (0, 0)], 0), 0)
    ys, xs = np.nonzero((score > 0) & (score > nb))
    return [(int(x), int(y), int(score[y, x])) for y, x in zip(ys, xs)]


def features_per_level(nfeatures=500, nlevels=8):
    factor = f32(1.0 / float(f32(1.2)))
    nd = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    out, s = [], 0
    for _ in range(nlevels - 1):
        n = int(np.rint(nd)); out.append(n); s += n
        nd = f32(nd * factor)
    out.append(max(nfeatures - s, 0))
    return out


def umax_table(half=15):
    umax = [0] * (half + 2)
    vmax = int(math.floor(half * math.sqrt(2.0) / 2 + 1)); vmin = int(math.ceil(half * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(half * half - v * v)))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0; v0 += 1
    return umax


UMAX = umax_table()


def retain_best_set(kps, n, key):
    """KeyPointsFilter::retainBest as a set: the n best plus everything tied with the n-th (input order kept)"""
    if n < 0 or len(kps) <= n:
        return list(kps)
    if n == 0:
        return []
    r = sorted((key(k) for k in kps), reverse=True)[n - 1]
    return [k for k in kps if key(k) >= r]


def harris_response(level, x, y):
    """HarrisResponses(blockSize 7, k 0.04): int sums of Ix^2, Iy^2, IxIy over the 7x7 block, then float32"""
    p = level[y - 4:y + 5, x - 4:x + 5].astype(np.int64)
    Ix = (p[1:-1, 2:] - p[1:-1, :-2]) * 2 + (p[:-2, 2:] - p[:-2, :-2]) + (p[2:, 2:] - p[2:, :-2])
    Iy = (p[2:, 1:-1] - p[:-2, 1:-1]) * 2 + (p[2:, :-2] - p[:-2, :-2]) + (p[2:, 2:] - p[:-2, 2:])
    a, b, c = f32(int((Ix * Ix).sum())), f32(int((Iy * Iy).sum())), f32(int((Ix * Iy).sum()))
    scale = f32(1.0) / (f32(4 * 7) * f32(255.0))
    s4 = f32(f32(f32(scale * scale) * scale) * scale)
    apb = f32(a + b)
    return f32(f32(f32(f32(a * b) - f32(c * c)) - f32(f32(f32(0.04) * apb) * apb)) * s4)


_P1 = f32(0.9997878412794807) * f32(180 / math.pi); _P3 = f32(-0.3258083974640975) * f32(180 / math.pi)
_P5 = f32(0.1555786518463281) * f32(180 / math.pi); _P7 = f32(-0.04432655554792128) * f32(180 / math.pi)
_EPS = f32(2.220446049250313e-16)


def fast_atan2(y, x):
    """cv::fastAtan2 (degrees, 7th-order odd polynomial), float32 without fused operations"""
    y = f32(y); x = f32(x)
    ax, ay = abs(x), abs(y)
    if ax >= ay:
        c = f32(ay / f32(ax + _EPS)); c2 = f32(c * c)
        a = f32(f32(f32(f32(f32(f32(f32(_P7 * c2) + _P5) * c2) + _P3) * c2) + _P1) * c)
    else:
        c = f32(ax / f32(ay + _EPS)); c2 = f32(c * c)
        a = f32(f32(90.0) - f32(f32(f32(f32(f32(f32(f32(_P7 * c2) + _P5) * c2) + _P3) * c2) + _P1) * c))
    if x < 0:
        a = f32(f32(180.0) - a)
    if y < 0:
        a = f32(f32(360.0) - a)
    return a


def ic_moments(level, x, y):
    """m_01, m_10 of the circular patch of radius 15 (ICAngles)"""
    I = level.astype(np.int64)
    u = np.arange(-15, 16)
    m10 = int((u * I[y, x - 15:x + 16]).sum()); m01 = 0
    for v in range(1, 16):
        d = UMAX[v]
        plus = I[y + v, x - d:x + d + 1]; minus = I[y - v, x - d:x + d + 1]
        m01 += v * int((plus - minus).sum())
        m10 += int((np.arange(-d, d + 1) * (plus + minus)).sum())
    return m01, m10


def detect_levels(gray, nlevels=8):
    """the unblurred pyramid levels detection works on"""
    H, W = gray.shape
    levels = [gray]
    for l in range(1, nlevels):
        w, h = level_size(W, H, l)
        levels.append(resize_linear_exact(levels[-1], w, h))
    return levels


def fast_candidates(level, threshold=20, edge=EDGE_THRESHOLD):
    """FAST corners of one level after non-max suppression and the border filter, raster order: (x, y, score)"""
    h, w = level.shape
    return [(x, y, s) for (x, y, s) in fast_nms(fast_score_map(level, threshold)) if edge <= x < w - edge and edge <= y < h - edge]


def detect_candidates(gray, nlevels=8):
    """every level's candidates with all the values OpenCV may need later (what the device hands to the host):
    (level, x, y, fast_score, harris float32, angle float32), raster order inside a level"""
    out = []
    for l, lev in enumerate(detect_levels(gray, nlevels)):
        for (x, y, s) in fast_candidates(lev):
            m01, m10 = ic_moments(lev, x, y)
            out.append((l, x, y, s, harris_response(lev, x, y), fast_atan2(m01, m10)))
    return out


def detect(gray, nfeatures=500, nlevels=8):
    """cv::ORB::detect as a set per level (raster order inside a level):
    -> list of (x float32, y float32, size float32, angle float32, response float32, octave)"""
    npl = features_per_level(nfeatures, nlevels)
    out = []
    for l, lev in enumerate(detect_levels(gray, nlevels)):
        kps = retain_best_set(fast_candidates(lev), 2 * npl[l], key=lambda c: c[2])
        kps = [(x, y, harris_response(lev, x, y)) for (x, y, s) in kps]
        kps = retain_best_set(kps, npl[l], key=lambda c: float(c[2]))
        sc = level_scale(l)
        for (x, y, hr) in kps:
            m01, m10 = ic_moments(lev, x, y)
            out.append((f32(f32(x) * sc), f32(f32(y) * sc), f32(f32(31.0) * sc), fast_atan2(m01, m10), hr, l))
    return out
