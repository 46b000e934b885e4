"""Seeded synthetic inputs for the five BASELINE.json configs.

Modelled on the reference's own synthetic tooling -- `Simulator` (src/Utilities/simulator.cpp:40-69,
119-171: random landmarks, noisy re-observation, ground-truth correspondences by landmark id) and
demoKabsch (demos/demoKabsch.cpp:118-126, 995-1019) -- but feeding keypoints + descriptors + depth
directly, because detection/description are outside the hot path.  numpy only; no GPU, no oracle.
"""
import math

import numpy as np

# resources/datasetConfig/freiburg1_desk.xml:5-17
FX, FY, CX, CY = 517.3, 516.5, 318.6, 255.3
DIST = (-0.0410, 0.3286, 0.0087, 0.0051, -0.5643)  # k1 k2 p1 p2 k3
DEPTH_SCALE = 5000.0
W, H = 640, 480
VAR_U, VAR_V = 1.1046, 0.6416                      # sigmaU/sigmaV attributes (used as Ruvd diagonal)
DIST_VAR_COEFS = (0.0, 0.002797, -0.004249, 0.007311)  # depth variance polynomial c3..c0 (fr1-like)


def rot_from_rotvec(rv):
    rv = np.asarray(rv, np.float64)
    th = float(np.linalg.norm(rv))
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


def random_descriptors(rng, n, nbytes=32):
    return rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)


def flip_bits(rng, desc, p):
    """Per-bit flip with probability p."""
    if p <= 0:
        return desc.copy()
    bits = np.unpackbits(desc, axis=1)
    flips = (rng.random(bits.shape) < p).astype(np.uint8)
    return np.packbits(bits ^ flips, axis=1)


def distort_points(uv_und, fx=FX, fy=FY, cx=CX, cy=CY, dist=DIST):
    """Forward Brown model: undistorted pixel -> distorted pixel (what a real lens would give)."""
    k1, k2, p1, p2, k3 = dist
    x = (uv_und[:, 0] - cx) / fx
    y = (uv_und[:, 1] - cy) / fy
    r2 = x * x + y * y
    rad = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([xd * fx + cx, yd * fy + cy], 1)


def _paint_depth(uv, z, rng=None, background=0):
    depth = np.full((H, W), background, np.uint16)
    u = np.clip(np.rint(uv[:, 0]).astype(int), 0, W - 1)
    v = np.clip(np.rint(uv[:, 1]).astype(int), 0, H - 1)
    depth[v, u] = np.clip(np.rint(z * DEPTH_SCALE), 0, 65535).astype(np.uint16)
    return depth


def frame_pair(n=500, seed=0, outlier_frac=0.25, flip_p=0.05, t_norm=0.05, rot=0.03, distorted=False):
    """C1: one synthetic 640x480 RGB-D frame pair with `n` planted features.

    Returns dict: desc1/desc2 (n x 32 u8), uv1/uv2 (float32 keypoints as a detector would report them,
    i.e. distorted pixels when `distorted`), depth1/depth2 (480x640 u16), T_gt (4x4, p1 ~= R p2 + t),
    is_outlier (n bool).  Frame-2 rows are shuffled so indices carry no information.
    """
    rng = np.random.default_rng(seed)
    uv1 = np.stack([rng.uniform(8, 631, n), rng.uniform(8, 471, n)], 1)
    z1 = rng.uniform(0.8, 5.0, n)
    p1 = np.stack([(uv1[:, 0] - CX) / FX * z1, (uv1[:, 1] - CY) / FY * z1, z1], 1)
    axis = rng.standard_normal(3); axis /= np.linalg.norm(axis)
    R = rot_from_rotvec(axis * rot)
    tdir = rng.standard_normal(3); tdir /= np.linalg.norm(tdir)
    t = tdir * t_norm
    # p1 = R p2 + t  ->  p2 = R^T (p1 - t)
    p2 = (p1 - t) @ R
    n_out = int(round(outlier_frac * n))
    is_out = np.zeros(n, bool)
    is_out[rng.choice(n, n_out, replace=False)] = True
    # outliers: fresh 3D point in frame 2
    uvo = np.stack([rng.uniform(8, 631, n), rng.uniform(8, 471, n)], 1)
    zo = rng.uniform(0.8, 5.0, n)
    po = np.stack([(uvo[:, 0] - CX) / FX * zo, (uvo[:, 1] - CY) / FY * zo, zo], 1)
    p2 = np.where(is_out[:, None], po, p2)
    uv2 = np.stack([p2[:, 0] / p2[:, 2] * FX + CX, p2[:, 1] / p2[:, 2] * FY + CY], 1)
    z2 = p2[:, 2]
    # points that leave the image or the depth range in frame 2 are re-planted as outliers inside it
    bad = (uv2[:, 0] < 8) | (uv2[:, 0] > 631) | (uv2[:, 1] < 8) | (uv2[:, 1] > 471) | (z2 < 0.8) | (z2 > 5.0)
    uv2 = np.where(bad[:, None], uvo, uv2); z2 = np.where(bad, zo, z2)
    is_out |= bad
    desc1 = random_descriptors(rng, n)
    desc2 = flip_bits(rng, desc1, flip_p)
    fresh = random_descriptors(rng, n)
    desc2 = np.where(is_out[:, None], fresh, desc2)
    depth1 = _paint_depth(uv1, z1)
    depth2 = _paint_depth(uv2, z2)
    perm = rng.permutation(n)
    uv2, desc2, is_out2 = uv2[perm], desc2[perm], is_out[perm]
    kp1, kp2 = uv1, uv2
    if distorted:
        kp1, kp2 = distort_points(uv1), distort_points(uv2)
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return dict(desc1=desc1, desc2=np.ascontiguousarray(desc2), uv1=kp1.astype(np.float32),
                uv2=kp2.astype(np.float32), depth1=depth1, depth2=depth2, T_gt=T, perm=perm,
                is_outlier=is_out2)


def helix_pose(i, n_frames=1000, radius=1.0, pitch=0.5):
    """C2 trajectory: one helix turn over n_frames (about 6.3 mm, 6.3 mrad per frame)."""
    a = 2 * math.pi * i / n_frames
    R = rot_from_rotvec([0.0, a, 0.0])
    t = np.array([radius * math.cos(a) - radius, pitch * i / n_frames, radius * math.sin(a)])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return T


class Sequence:
    """C2: TUM-shaped sequence, `n_kp` keypoints per frame through a room of random landmarks.

    frame(i) -> dict(desc, uv, depth) for the first n_kp landmarks (by id) visible from helix_pose(i).
    """

    def __init__(self, n_frames=1000, n_kp=1000, n_landmarks=20000, seed=42, flip_p=0.05):
        self.n_frames, self.n_kp, self.flip_p = n_frames, n_kp, flip_p
        rng = np.random.default_rng(seed)
        self.seed = seed
        # landmarks on a shell 2..5 m around the helix axis so that every pose sees plenty of them
        d = rng.standard_normal((n_landmarks, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        self.landmarks = d * rng.uniform(2.0, 5.0, (n_landmarks, 1)) + np.array([-1.0, 0.25, 0.0])
        self.desc = random_descriptors(rng, n_landmarks)

    def frame(self, i):
        rng = np.random.default_rng([self.seed, i])
        T = helix_pose(i, self.n_frames)
        R, t = T[:3, :3], T[:3, 3]
        pc = (self.landmarks - t) @ R            # world -> camera
        z = pc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            u = pc[:, 0] / z * FX + CX
            v = pc[:, 1] / z * FY + CY
        vis = (z > 0.8) & (z < 5.0) & (u > 8) & (u < 631) & (v > 8) & (v < 471)
        ids = np.nonzero(vis)[0][: self.n_kp]
        uv = np.stack([u[ids], v[ids]], 1)
        desc = flip_bits(rng, self.desc[ids], self.flip_p)
        return dict(desc=desc, uv=uv.astype(np.float32), depth=_paint_depth(uv, z[ids]), ids=ids, T_wc=T)


def map_frame(M=5000, N=1000, n_reobs=700, seed=0, sigma=0.01, flip_p=0.05, pose_err_t=0.03,
              pose_err_rot=0.02):
    """C3: frame-to-map inputs for Matcher::matchXYZ.

    Map features live in the (predicted) camera frame; the current frame re-observes n_reobs of them
    under a small unknown pose error, plus N-n_reobs clutter keypoints.  Returns map_xyz (float64, as
    MapFeature.position), map_desc, map_octave, map_detdist, cur_xyz (float32), cur_desc, cur_octave,
    cur_detdist (double), T_gt (map ~= R cur + t).
    """
    rng = np.random.default_rng(seed)

    def frustum(n):
        z = rng.uniform(0.8, 5.0, n)
        u = rng.uniform(8, 631, n); v = rng.uniform(8, 471, n)
        return np.stack([(u - CX) / FX * z, (v - CY) / FY * z, z], 1)

    map_xyz = frustum(M)
    map_desc = random_descriptors(rng, M)
    map_oct = rng.integers(0, 8, M).astype(np.int32)
    # distance at which the feature was described: around the current distance (+-25 %)
    map_detdist = np.linalg.norm(map_xyz, axis=1) * rng.uniform(0.75, 1.25, M)
    axis = rng.standard_normal(3); axis /= np.linalg.norm(axis)
    R = rot_from_rotvec(axis * pose_err_rot)
    tdir = rng.standard_normal(3); tdir /= np.linalg.norm(tdir)
    t = tdir * pose_err_t
    ids = rng.choice(M, n_reobs, replace=False)
    cur_re = (map_xyz[ids] - t) @ R + rng.normal(0, sigma, (n_reobs, 3))
    cur_cl = frustum(N - n_reobs)
    cur_xyz = np.concatenate([cur_re, cur_cl]).astype(np.float32)
    cur_desc = np.concatenate([flip_bits(rng, map_desc[ids], flip_p), random_descriptors(rng, N - n_reobs)])
    # the current frame detects at roughly the octave the map predicts, jittered by one level
    cur_oct = np.concatenate([np.clip(map_oct[ids] + rng.integers(-1, 2, n_reobs), 0, 7),
                              rng.integers(0, 8, N - n_reobs)]).astype(np.int32)
    perm = rng.permutation(N)
    cur_xyz, cur_desc, cur_oct = cur_xyz[perm], np.ascontiguousarray(cur_desc[perm]), cur_oct[perm]
    # detDist of the current keypoints = float32 norm widened (matcher.cpp:51-58)
    x, y, z = cur_xyz[:, 0], cur_xyz[:, 1], cur_xyz[:, 2]
    cur_detdist = np.sqrt((x * x + y * y + z * z).astype(np.float32)).astype(np.float64)
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return dict(map_xyz=map_xyz, map_desc=map_desc, map_octave=map_oct, map_detdist=map_detdist,
                cur_xyz=cur_xyz, cur_desc=cur_desc, cur_octave=cur_oct, cur_detdist=cur_detdist,
                T_gt=T, reobs_ids=ids, perm=perm)


def keyframe_db(n_kf=10000, per_kf=1000, n_query=1000, n_planted=20, shared=400, seed=7, flip_p=0.05,
                ragged=False):
    """C4: loop-closure database.  `n_planted` keyframes share `shared` landmarks with the query.

    Returns db (sum(per_kf) x 32 u8), kf_off (n_kf+1 int64), query (n_query x 32), planted (ids).
    Generated in chunks so that the 320 MB C4 database does not need a second copy.
    """
    rng = np.random.default_rng(seed)
    if ragged:
        counts = rng.integers(max(1, per_kf // 2), per_kf + 1, n_kf)
    else:
        counts = np.full(n_kf, per_kf)
    kf_off = np.zeros(n_kf + 1, np.int64); kf_off[1:] = np.cumsum(counts)
    db = np.empty((int(kf_off[-1]), 32), np.uint8)
    step = 1 << 20
    for s in range(0, db.shape[0], step):
        e = min(db.shape[0], s + step)
        db[s:e] = rng.integers(0, 256, size=(e - s, 32), dtype=np.uint8)
    query = random_descriptors(rng, n_query)
    planted = np.sort(rng.choice(n_kf, min(n_planted, n_kf), replace=False))
    for k in planted:
        c = int(counts[k])
        s = min(shared, c, n_query)
        qi = rng.choice(n_query, s, replace=False)
        ti = rng.choice(c, s, replace=False)
        db[kf_off[k] + ti] = flip_bits(rng, query[qi], flip_p)
    return dict(db=db, kf_off=kf_off, query=query, planted=planted)


def matched_clouds(m=1000, inlier_frac=0.6, seed=0, sigma=0.005):
    """C5 RANSAC sweep input: m correspondences between two clouds, a fraction of them consistent."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(0.8, 5.0, m)
    u = rng.uniform(8, 631, m); v = rng.uniform(8, 471, m)
    cur = np.stack([(u - CX) / FX * z, (v - CY) / FY * z, z], 1)
    axis = rng.standard_normal(3); axis /= np.linalg.norm(axis)
    R = rot_from_rotvec(axis * 0.05)
    t = np.array([0.03, -0.02, 0.04])
    prev = cur @ R.T + t + rng.normal(0, sigma, (m, 3))
    n_out = m - int(round(inlier_frac * m))
    out = rng.choice(m, n_out, replace=False)
    zo = rng.uniform(0.8, 5.0, n_out)
    prev[out] = np.stack([(rng.uniform(8, 631, n_out) - CX) / FX * zo, (rng.uniform(8, 471, n_out) - CY) / FY * zo, zo], 1)
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    mq = np.arange(m, dtype=np.int32); mt = rng.permutation(m).astype(np.int32)
    cur_p = np.empty_like(cur); cur_p[mt] = cur      # cur_p[mt[k]] pairs with prev[k]
    return dict(prev=prev.astype(np.float32), cur=cur_p.astype(np.float32), mq=mq, mt=mt, T_gt=T)
