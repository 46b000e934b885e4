"""CPU restatement (numpy) of the KLT tracking path -- TEST INFRASTRUCTURE ONLY, never imported by the product.

Reference seam: MatcherOpenCV::performTracking (virtual, include/putslam/Matcher/matcher.h:405-422; src/Matcher/
matcherOpenCV.cpp:209-300): cv::calcOpticalFlowPyrLK(prevImg, img, prevFeatures, features, status, err, winSize,
maxLevels, termcrit, flags, minEigThreshold) on the COLOUR frames (src/Matcher/matcher.cpp:151), then the error
threshold, the pairwise "too close" removal and the compaction into DMatch(i, j, 0).  The arithmetic is OpenCV's
(un-vendored dependency); restated from OpenCV 4.13's video/lkpyramid.cpp behaviour and pinned bit for bit against cv2
(tests/golden/klt_cv2.npz, tests/test_oracle_cpu.py::test_klt_*): positions, status and err, gray and 3-channel frames,
OPTFLOW_USE_INITIAL_FLOW and OPTFLOW_LK_GET_MIN_EIGENVALS.

  pyramid   level l+1 = pyrDown(level l): 5x5 kernel [1 4 6 4 1]^2, BORDER_REFLECT_101, (sum + 128) >> 8, size (n+1)/2
  gradient  Scharr on the level (3,10,3 smoothing x central difference), int, reflect-101 inside, zero outside the image
  per point, coarse to fine: window (win x win x channels) of I, Ix, Iy at the sub-pixel position by bilinear weights in
            2^14 fixed point (cvRound of float32 products), rounded shifts by 9 (image, keeps 5 fractional bits) / 14
            (gradient); A11, A12, A22 = float32 sums of the integer products in the lane order of facc() below, x 2^-20;
            minEig test;
            Newton steps delta = A^-1 b with b from the same interpolation of J, until |delta|^2 <= eps^2, 30
            iterations, or the oscillation rule (then half a step back); err = mean |J - I| / 32 over the window
"""
import numpy as np

f32 = np.float32
def descale(x, n): return (x + (1 << (n - 1))) >> n
def pyr_down(img):                         # H x W x C uint8
    H, W = img.shape[:2]
    oh, ow = (H + 1) // 2, (W + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int64)
    p = np.pad(img.astype(np.int64), ((2, 2), (2, 2), (0, 0)), mode="reflect")
    cols = 2 * np.arange(ow); rows = 2 * np.arange(oh)
    h = sum(k[i] * p[:, cols + i] for i in range(5))
    v = sum(k[i] * h[rows + i] for i in range(5))
    return ((v + 128) >> 8).astype(np.uint8)
def scharr_deriv(img):                     # -> dx, dy int32 H x W x C
    I = np.pad(img.astype(np.int32), ((1, 1), (1, 1), (0, 0)), mode="reflect")
    t0 = 3 * (I[:-2] + I[2:]) + 10 * I[1:-1]
    t1 = I[2:] - I[:-2]
    dx = t0[:, 2:] - t0[:, :-2]
    dy = 3 * (t1[:, :-2] + t1[:, 2:]) + 10 * t1[:, 1:-1]
    return dx, dy
def weights(a, b):
    w00 = int(np.rint(f32(f32(f32(1) - a) * f32(f32(1) - b)) * f32(1 << 14)))
    w01 = int(np.rint(f32(a * f32(f32(1) - b)) * f32(1 << 14)))
    w10 = int(np.rint(f32(f32(f32(1) - a) * b) * f32(1 << 14)))
    return w00, w01, w10, (1 << 14) - w00 - w01 - w10
def facc(prod, paired=False):
    """float32 sum of the window's integer products in the order OpenCV's 128-bit vector loop leaves: in every window row
    the first 8 * floor(cols / 8) values (cols = win * channels) go to four float accumulators by x mod 4, the remaining
    values of the row to a scalar accumulator; at the end  scalar + ((acc0 + acc2) + (acc1 + acc3)).
    paired=False (the A11 / A12 / A22 sums): every product is converted and added on its own.
    paired=True  (the b1 / b2 sums of the Newton steps): within each group of 8 the products x and x + 4 are first added
    as integers (a 16-bit multiply-add instruction produces both at once), then converted and added -- one rounding
    instead of two.  For a gray 7-wide window both are the plain raster order.  Pinned by tests/golden/klt_cv2.npz for
    the reference's 7 x 7 window and live against cv2 for windows 3 ... 21, gray and colour, on hard-edged frames (where
    any other order differs in the last bit)."""
    P = prod.reshape(prod.shape[0], -1)
    rows, cols = P.shape
    n8 = (cols // 8) * 8
    acc = [f32(0)] * 4
    s = f32(0)
    for y in range(rows):
        if paired:
            for x0 in range(0, n8, 8):
                for k in range(4):
                    acc[k] = f32(acc[k] + f32(int(P[y, x0 + k]) + int(P[y, x0 + k + 4])))
        else:
            for x in range(n8):
                acc[x & 3] = f32(acc[x & 3] + f32(int(P[y, x])))
        for x in range(n8, cols):
            s = f32(s + f32(int(P[y, x])))
    return f32(s + f32(f32(acc[0] + acc[2]) + f32(acc[1] + acc[3])))


def lk_pyr(I0, J0, pts, win=7, max_level=3, max_iter=30, eps=0.01, min_eig_thr=1e-4, init=None, min_eig_err=False):
    if I0.ndim == 2: I0 = I0[..., None]; J0 = J0[..., None]
    cn = I0.shape[2]
    # calcOpticalFlowPyrLK's normalisation of the criteria (count to [0, 100], epsilon to [0, 10]) and
    # buildOpticalFlowPyramid's depth: the pyramid ends where the next level would not exceed the window
    max_iter = min(max(int(max_iter), 0), 100); eps = min(max(float(eps), 0.0), 10.0)
    Is, Js = [I0], [J0]
    for l in range(max_level):
        h, w = Is[-1].shape[:2]
        if (w + 1) // 2 <= win or (h + 1) // 2 <= win:
            max_level = l
            break
        Is.append(pyr_down(Is[-1])); Js.append(pyr_down(Js[-1]))
    n = len(pts)
    nxt = np.zeros((n, 2), np.float32) if init is None else np.array(init, np.float32).copy()
    status = np.ones(n, np.uint8); err = np.zeros(n, np.float32)
    half = f32((win - 1) * 0.5); P = win + 1; FLT_SCALE = f32(1.0) / f32(1 << 20)
    for level in range(max_level, -1, -1):
        I, J = Is[level], Js[level]; H, W = I.shape[:2]
        dxI, dyI = scharr_deriv(I)
        pad = ((P, P), (P, P), (0, 0))
        Ip = np.pad(I.astype(np.int32), pad, mode="reflect"); Jp = np.pad(J.astype(np.int32), pad, mode="reflect")
        dxp = np.pad(dxI, pad, mode="constant"); dyp = np.pad(dyI, pad, mode="constant")
        def interp(A, x0, y0, ws, sh):
            blk = A[y0 + P:y0 + P + win + 1, x0 + P:x0 + P + win + 1]
            return descale(blk[:-1, :-1] * ws[0] + blk[:-1, 1:] * ws[1] + blk[1:, :-1] * ws[2] + blk[1:, 1:] * ws[3], sh)
        scale = f32(1.0 / (1 << level))
        for i in range(n):
            px = f32(f32(pts[i, 0]) * scale); py = f32(f32(pts[i, 1]) * scale)
            if level == max_level:
                if init is not None: nx = f32(nxt[i, 0] * scale); ny = f32(nxt[i, 1] * scale)
                else: nx, ny = px, py
            else: nx = f32(nxt[i, 0] * f32(2)); ny = f32(nxt[i, 1] * f32(2))
            nxt[i] = (nx, ny)
            px = f32(px - half); py = f32(py - half)
            ix = int(np.floor(px)); iy = int(np.floor(py))
            if ix < -win or ix >= W or iy < -win or iy >= H:
                if level == 0: status[i] = 0; err[i] = 0
                continue
            ws = weights(f32(px - f32(ix)), f32(py - f32(iy)))
            Iw = interp(Ip, ix, iy, ws, 9); Ix = interp(dxp, ix, iy, ws, 14); Iy = interp(dyp, ix, iy, ws, 14)
            A11 = f32(facc(Ix * Ix) * FLT_SCALE); A12 = f32(facc(Ix * Iy) * FLT_SCALE); A22 = f32(facc(Iy * Iy) * FLT_SCALE)
            D = f32(f32(A11 * A22) - f32(A12 * A12))
            minEig = f32(f32(f32(A22 + A11) - f32(np.sqrt(f32(f32(f32(A11 - A22) * f32(A11 - A22)) + f32(f32(f32(4) * A12) * A12))))) / f32(2 * win * win))
            if min_eig_err: err[i] = minEig
            if float(minEig) < min_eig_thr or D < np.finfo(np.float32).eps:                       # threshold is a double
                if level == 0: status[i] = 0
                continue
            D = f32(f32(1) / D)
            nx = f32(nx - half); ny = f32(ny - half)
            pdx = pdy = f32(0)
            for j in range(max_iter):
                jx = int(np.floor(nx)); jy = int(np.floor(ny))
                if jx < -win or jx >= W or jy < -win or jy >= H:
                    if level == 0: status[i] = 0
                    break
                ws2 = weights(f32(nx - f32(jx)), f32(ny - f32(jy)))
                diff = interp(Jp, jx, jy, ws2, 9) - Iw
                b1 = f32(facc(diff * Ix, True) * FLT_SCALE); b2 = f32(facc(diff * Iy, True) * FLT_SCALE)
                dx = f32(f32(f32(A12 * b2) - f32(A22 * b1)) * D); dy = f32(f32(f32(A12 * b1) - f32(A11 * b2)) * D)
                nx = f32(nx + dx); ny = f32(ny + dy)
                nxt[i] = (f32(nx + half), f32(ny + half))
                if float(dx) * float(dx) + float(dy) * float(dy) <= eps * eps: break
                if j > 0 and abs(float(f32(dx + pdx))) < 0.01 and abs(float(f32(dy + pdy))) < 0.01:   # float sum, double compare
                    nxt[i] = (f32(nxt[i, 0] - f32(dx * f32(0.5))), f32(nxt[i, 1] - f32(dy * f32(0.5)))); break
                pdx, pdy = dx, dy
            if status[i] and level == 0 and not min_eig_err:
                qx = f32(nxt[i, 0] - half); qy = f32(nxt[i, 1] - half)
                jx = int(np.floor(qx)); jy = int(np.floor(qy))
                if jx < -win or jx >= W or jy < -win or jy >= H:
                    status[i] = 0; continue
                ws3 = weights(f32(qx - f32(jx)), f32(qy - f32(jy)))
                diff = interp(Jp, jx, jy, ws3, 9) - Iw
                e = f32(0)
                for v in diff.ravel(): e = f32(e + f32(abs(int(v))))
                err[i] = f32(f32(e * f32(1.0)) / f32(32 * win * cn * win))
    return nxt, status, err


def perform_tracking(err, status, pts, error_threshold, min_distance):
    """the part of MatcherOpenCV::performTracking after calcOpticalFlowPyrLK (matcherOpenCV.cpp:247-290):
    status cleared where err > error_threshold; of every pair of tracked positions closer than min_distance (ALL
    features, whatever their status) the one with the larger err is dropped (ties: the second); survivors in order.
    -> indices kept (DMatch.queryIdx), new index = position in the list"""
    err = np.asarray(err, np.float32); pts = np.asarray(pts, np.float32).reshape(-1, 2)
    st = np.asarray(status, np.uint8).copy()
    st[err > error_threshold] = 0
    n = len(pts)
    remove = np.zeros(n, bool)
    for i in range(n):
        d = pts[i + 1:] - pts[i]                                     # float32 differences
        close = np.sqrt(d[:, 0].astype(np.float64) ** 2 + d[:, 1].astype(np.float64) ** 2) < min_distance
        for j in np.nonzero(close)[0] + i + 1:
            if err[i] > err[j]:
                remove[i] = True
            else:
                remove[j] = True
    return np.nonzero((st != 0) & ~remove)[0]


# ---- the host-side list edits of Matcher::trackKLT ------------------------------------------------------------------
def remove_too_close(und, xyz, min_euclid, min_reproj):
    """Matcher::removeTooCloseFeatures (src/Matcher/matcher.cpp:886-974): feature j is removed when some i < j (removed or
    not) is closer than min_euclid in 3-D or than min_reproj in the undistorted image; float differences, double norms.
    -> sorted indices removed"""
    und = np.asarray(und, np.float32).reshape(-1, 2); xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    gone = np.zeros(len(und), bool)
    for i in range(len(und)):
        d3 = (xyz[i] - xyz[i + 1:]).astype(np.float64); d2 = (und[i] - und[i + 1:]).astype(np.float64)
        dist3 = np.sqrt(d3[:, 0] * d3[:, 0] + d3[:, 1] * d3[:, 1] + d3[:, 2] * d3[:, 2])
        dist2 = np.sqrt(d2[:, 0] * d2[:, 0] + d2[:, 1] * d2[:, 1])
        gone[i + 1:] |= (dist3 < min_euclid) | (dist2 < min_reproj)
    return np.nonzero(gone)[0]


def merge_tracked(und, sandbox_und, min_reproj):
    """Matcher::mergeTrackedFeatures (matcher.cpp:96-131): a newly detected feature is appended unless some feature already
    in the list (tracked, or appended before it) is closer than min_reproj in the undistorted image.
    -> indices of the sandbox features appended, in order"""
    cur = [np.asarray(p, np.float32) for p in np.asarray(und, np.float32).reshape(-1, 2)]
    added = []
    for i, c in enumerate(np.asarray(sandbox_und, np.float32).reshape(-1, 2)):
        if cur:
            d = (c - np.array(cur, np.float32)).astype(np.float64)
            if (np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < min_reproj).any():
                continue
        cur.append(c); added.append(i)
    return np.array(added, np.int64)


def predict_description_levels(xyz, octave, det_dist, scale_factor=1.2, n_levels=8):
    """matcher.cpp:281-322: level = clamp(ceil(log(scale^octave * detDist / |p|) / log(scale)), 0, n_levels - 1) with |p| a
    float norm ((x*x + y*y) + z*z, float square root), everything else double with the C library's pow / log -- and the
    stable regrouping by that level.  -> (levels per original index, order = original indices level by level)"""
    import math
    xyz = np.asarray(xyz, np.float32).reshape(-1, 3)
    lv = np.zeros(len(xyz), np.int32)
    for i, (p, o, d) in enumerate(zip(xyz, octave, det_dist)):
        cur = float(np.sqrt(f32(f32(f32(p[0] * p[0]) + f32(p[1] * p[1])) + f32(p[2] * p[2]))))
        with np.errstate(all="ignore"):
            s = np.float64(math.pow(scale_factor, int(o))) * np.float64(d) / np.float64(cur)
            v = np.ceil(np.log(s) / math.log(scale_factor))
        # (int) of a NaN / out-of-range double is undefined in C; x86-64 yields INT_MIN, which the clamps turn into 0
        c = int(v) if np.isfinite(v) and abs(v) < 2 ** 31 else -2 ** 31
        lv[i] = min(n_levels - 1, max(0, c))
    order = np.concatenate([np.nonzero(lv == l)[0] for l in range(n_levels)]) if len(lv) else np.zeros(0, np.int64)
    return lv, order
