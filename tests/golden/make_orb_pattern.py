"""Recovers the ORB descriptor's 256 x 4 sampling pattern (OpenCV's bit_pattern_31_) from the cv2 binary.

OpenCV is an un-vendored dependency of the reference (describeFeatures -> cv::ORB::compute,
src/Matcher/matcherOpenCV.cpp:181-195); its source is not available offline, cv2 4.13.0 is.  For a keypoint at angle 0
on level 0, bit i of the descriptor is  I(p0_i) < I(p1_i)  on the 7x7-blurred image.  An impulse image (one bright
pixel at q) sets bit i exactly where q lies in the blur footprint of p1_i (and closer to it than to p0_i); the
inverted image does the same for p0_i.  Step 1 reads the footprint centres, step 2 resolves the clipped footprints of
close point pairs by simulating the blur.  Output: oracle/orb_pattern_31.txt and putslam_b200/csrc/orb_pattern.inc
(written by the snippet at the end).  Run in the build container:  python tests/golden/make_orb_pattern.py
"""
import numpy as np, cv2, time
orb = cv2.ORB_create()
S = 129; C = 64
def desc_for(img):
    kp = [cv2.KeyPoint(float(C), float(C), 31.0, 0.0, 1.0, 0)]   # x, y, size, angle=0, response, octave 0
    k2, d = orb.compute(img, kp)
    assert len(k2) == 1
    return np.unpackbits(d[0], bitorder="little")          # bit i of byte j = test 8j+i ? check order below
R = 19
hits1 = np.zeros((256, 2 * R + 1, 2 * R + 1), np.uint8)   # bright blob: bit=1 where p1 near q
hits0 = np.zeros((256, 2 * R + 1, 2 * R + 1), np.uint8)   # dark blob on bright: bit=1 where p0 near q
t0 = time.time()
for dy in range(-R, R + 1):
    for dx in range(-R, R + 1):
        img = np.zeros((S, S), np.uint8); img[C + dy, C + dx] = 255
        hits1[:, dy + R, dx + R] = desc_for(img)
        img = np.full((S, S), 255, np.uint8); img[C + dy, C + dx] = 0
        hits0[:, dy + R, dx + R] = desc_for(img)
print("probing", time.time() - t0, "s")
np.savez("/tmp/orb_hits.npz", hits0=hits0, hits1=hits1)
pat = np.zeros((256, 4), np.int32)
ok = True
for i in range(256):
    for (h, col) in ((hits0[i], 0), (hits1[i], 2)):
        ys, xs = np.nonzero(h)
        if len(ys) == 0:
            print("bit", i, "no hits", col); ok = False; continue
        # the hit region is the 7x7 blur footprint around the point (possibly clipped by the other point's footprint)
        cy = (ys.min() + ys.max()) / 2 - R; cx = (xs.min() + xs.max()) / 2 - R
        pat[i, col] = int(round(cx)); pat[i, col + 1] = int(round(cy))
        if ys.max() - ys.min() != 6 or xs.max() - xs.min() != 6:
            print("bit", i, "col", col, "footprint", ys.min() - R, ys.max() - R, xs.min() - R, xs.max() - R, "n", len(ys))
print(pat[:8])
np.save("/tmp/orb_pattern_raw.npy", pat)

# ---- step 2: clipped footprints
import itertools
d = np.load("/tmp/orb_hits.npz"); hits0, hits1 = d["hits0"], d["hits1"]
pat = np.load("/tmp/orb_pattern_raw.npy")
R = 19
k = np.array([18, 34, 48, 56, 48, 34, 18], np.int64)
def val(dy, dx, bright):
    """blurred intensity at offset (dy,dx) from an impulse (bright: 255 on 0; dark: 0 on 255)"""
    if abs(dy) > 3 or abs(dx) > 3:
        w = 0
    else:
        w = k[dy + 3] * k[dx + 3]
    v = w * 255 if bright else (65536 - w) * 255
    return (v + 32768) >> 16
qs = [(dy, dx) for dy in range(-R, R + 1) for dx in range(-R, R + 1)]
def sim(p0, p1):
    h1 = np.zeros((2 * R + 1, 2 * R + 1), np.uint8); h0 = np.zeros_like(h1)
    for (qy, qx) in qs:
        a = val(p0[1] - qy, p0[0] - qx, True); b = val(p1[1] - qy, p1[0] - qx, True)
        h1[qy + R, qx + R] = a < b
        a = val(p0[1] - qy, p0[0] - qx, False); b = val(p1[1] - qy, p1[0] - qx, False)
        h0[qy + R, qx + R] = a < b
    return h0, h1
fixed = 0
for i in range(256):
    p0 = tuple(pat[i, :2]); p1 = tuple(pat[i, 2:])
    h0, h1 = sim(p0, p1)
    if np.array_equal(h0, hits0[i]) and np.array_equal(h1, hits1[i]):
        continue
    best = None
    for d0 in itertools.product((-1, 0, 1), repeat=2):
        for d1 in itertools.product((-1, 0, 1), repeat=2):
            c0 = (p0[0] + d0[0], p0[1] + d0[1]); c1 = (p1[0] + d1[0], p1[1] + d1[1])
            h0, h1 = sim(c0, c1)
            if np.array_equal(h0, hits0[i]) and np.array_equal(h1, hits1[i]):
                best = (c0, c1); break
        if best: break
    if best is None:
        print("bit", i, "unresolved", p0, p1)
    else:
        pat[i, :2] = best[0]; pat[i, 2:] = best[1]; fixed += 1
print("refined", fixed)

print(pat[:4].tolist(), pat[-2:].tolist(), np.abs(pat).max())

# ---- write the two copies of the table
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
np.savetxt(os.path.join(ROOT, "oracle", "orb_pattern_31.txt"), pat, fmt="%d",
           header="ORB rBRIEF pattern, patch 31: x0 y0 x1 y1 per test (recovered from cv2 4.13.0, tests/golden/make_orb_pattern.py)")
lines = ["// 256 binary tests (x0, y0, x1, y1) of the ORB descriptor, patch 31 -- the learned rBRIEF pattern of Rublee et al. 2011",
         "// as shipped in OpenCV (features2d, bit_pattern_31_).  OpenCV is a third-party dependency absent from the reference",
         "// tree; the table was recovered from the cv2 4.13.0 binary by probing cv::ORB::compute with impulse images",
         "// (tests/golden/make_orb_pattern.py) and is pinned by tests/golden/orb_cv2.npz."]
for i in range(0, 256, 4):
    lines.append("    " + "  ".join("%d,%d,%d,%d," % tuple(pat[j]) for j in range(i, i + 4)))
open(os.path.join(ROOT, "putslam_b200", "csrc", "orb_pattern.inc"), "w").write("\n".join(lines) + "\n")
