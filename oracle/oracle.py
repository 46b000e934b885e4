"""ctypes front for oracle/liboracle.so -- the CPU restatement of PUTSLAM's hot path.

TEST INFRASTRUCTURE ONLY (see oracle.c header): imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never by putslam_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class RansacParams(C.Structure):
    _fields_ = [
        ("error_version", C.c_int),
        ("inlier_threshold_euclidean", C.c_double),
        ("inlier_threshold_reprojection", C.c_double),
        ("minimal_inlier_ratio_threshold", C.c_double),
        ("minimal_number_of_matches", C.c_int),
        ("used_pairs", C.c_int),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
    ]


def default_ransac_params(error_version=0):
    # resources/putslammatcherOpenCVParameters.xml:29-37, freiburg1 intrinsics
    return RansacParams(error_version, 0.04, 2.0, 0.2, 15, 3, 517.3, 516.5, 318.6, 255.3)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_point_inlier_ratio.restype = C.c_double
        _lib.orc_norm3f.restype = C.c_float
        _lib.orc_norm3f.argtypes = [C.c_float] * 3
        _lib.orc_pred_level.argtypes = [C.c_int, C.c_double, C.c_double]
        _lib.orc_ransac_iterations.argtypes = [C.c_double]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def hamming_xor(a, b):
    a, b = _u8(a), _u8(b)
    return lib().orc_hamming_xor(_p(a, C.c_uint8), _p(b, C.c_uint8), a.size)


def hamming_satsub(a, b):
    a, b = _u8(a), _u8(b)
    return lib().orc_hamming_satsub(_p(a, C.c_uint8), _p(b, C.c_uint8), a.size)


def bf_mutual(q, t, threads=0):
    """-> (queryIdx int32[n], trainIdx int32[n], distance float32[n])"""
    q, t = _u8(q), _u8(t)
    nq, nt = q.shape[0], t.shape[0]
    nb = q.shape[1] if q.ndim == 2 else (t.shape[1] if t.ndim == 2 else 32)
    cap = max(1, min(nq, nt))
    oq = np.empty(cap, np.int32); ot = np.empty(cap, np.int32); od = np.empty(cap, np.float32)
    if threads and threads > 1:
        n = lib().orc_bf_mutual_mt(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb,
                                   _p(oq, C.c_int), _p(ot, C.c_int), _p(od, C.c_float), threads)
    else:
        n = lib().orc_bf_mutual(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb,
                                _p(oq, C.c_int), _p(ot, C.c_int), _p(od, C.c_float))
    return oq[:n].copy(), ot[:n].copy(), od[:n].copy()


def knn2(q, t):
    """-> (idx int32[nq,2], dist int32[nq,2]); -1 where fewer than 2 train descriptors"""
    q, t = _u8(q), _u8(t)
    nq, nt = q.shape[0], t.shape[0]
    nb = q.shape[1]
    idx = np.empty((nq, 2), np.int32); dist = np.empty((nq, 2), np.int32)
    lib().orc_knn2(_p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb, _p(idx, C.c_int), _p(dist, C.c_int))
    return idx, dist


def pred_level(octave, det_dist, cur_dist):
    return lib().orc_pred_level(int(octave), float(det_dist), float(cur_dist))


def pred_levels(octaves, det_dists, cur_dists):
    return np.array([pred_level(o, d, c) for o, d, c in zip(octaves, det_dists, cur_dists)], np.int32)


def guided_match(map_xyz, map_desc, map_level, cur_xyz, cur_desc, cur_level, radius, ratio, mode=0, cap=None):
    """-> (queryIdx, trainIdx, distance, perfect_count)"""
    map_xyz = np.ascontiguousarray(map_xyz, np.float32); cur_xyz = np.ascontiguousarray(cur_xyz, np.float32)
    map_desc, cur_desc = _u8(map_desc), _u8(cur_desc)
    map_level = np.ascontiguousarray(map_level, np.int32); cur_level = np.ascontiguousarray(cur_level, np.int32)
    M, N = map_xyz.shape[0], cur_xyz.shape[0]
    nb = map_desc.shape[1] if map_desc.ndim == 2 else 32
    if cap is None:
        cap = max(1, M * N)
    oq = np.empty(cap, np.int32); ot = np.empty(cap, np.int32); od = np.empty(cap, np.float32)
    perfect = C.c_int(0)
    n = lib().orc_guided_match(_p(map_xyz, C.c_float), _p(map_desc, C.c_uint8), _p(map_level, C.c_int), M,
                               _p(cur_xyz, C.c_float), _p(cur_desc, C.c_uint8), _p(cur_level, C.c_int), N,
                               nb, C.c_double(radius), C.c_double(ratio), mode,
                               _p(oq, C.c_int), _p(ot, C.c_int), _p(od, C.c_float), cap, C.byref(perfect))
    n = min(n, cap)
    return oq[:n].copy(), ot[:n].copy(), od[:n].copy(), perfect.value


def lc_scores(query, db, kf_off, tau=64, threads=1):
    query, db = _u8(query), _u8(db)
    kf_off = np.ascontiguousarray(kf_off, np.int64)
    n_kf = kf_off.size - 1
    scores = np.zeros(max(1, n_kf), np.int32)
    lib().orc_lc_scores(_p(query, C.c_uint8), query.shape[0], _p(db, C.c_uint8), _p(kf_off, C.c_int64),
                        n_kf, query.shape[1], tau, _p(scores, C.c_int), threads)
    return scores[:n_kf]


def topk(scores, k):
    scores = np.ascontiguousarray(scores, np.int32)
    ids = np.empty(k, np.int32); sc = np.empty(k, np.int32)
    lib().orc_topk(_p(scores, C.c_int), scores.size, k, _p(ids, C.c_int), _p(sc, C.c_int))
    return ids, sc


def undistort(uv, fx, fy, cx, cy, dist5):
    uv = np.ascontiguousarray(uv, np.float32)
    d = np.ascontiguousarray(dist5, np.float32)
    out = np.empty_like(uv)
    lib().orc_undistort(_p(uv, C.c_float), uv.shape[0], C.c_float(fx), C.c_float(fy), C.c_float(cx),
                        C.c_float(cy), _p(d, C.c_float), _p(out, C.c_float))
    return out


def backproject(uv, depth, fx, fy, cx, cy, depth_scale):
    """-> (xyz float32[n,3], det_dist float64[n])"""
    uv = np.ascontiguousarray(uv, np.float32)
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    n = uv.shape[0]
    xyz = np.empty((n, 3), np.float32); dd = np.empty(n, np.float64)
    lib().orc_backproject(_p(uv, C.c_float), n, _p(depth, C.c_uint16), W, H, W, C.c_float(fx), C.c_float(fy),
                          C.c_float(cx), C.c_float(cy), C.c_double(depth_scale), _p(xyz, C.c_float),
                          _p(dd, C.c_double))
    return xyz, dd


def compute_cov(u, v, depth, fx, fy, cx, cy, varU, varV, coefs):
    c = np.ascontiguousarray(coefs, np.float64)
    cov = np.empty(9, np.float64)
    lib().orc_compute_cov(C.c_uint(int(u)), C.c_uint(int(v)), C.c_double(depth), C.c_double(fx), C.c_double(fy),
                          C.c_double(cx), C.c_double(cy), C.c_double(varU), C.c_double(varV),
                          _p(c, C.c_double), _p(cov, C.c_double))
    return cov.reshape(3, 3)


def information_matrix(u, v, z, fx, fy, cx, cy, varU, varV, coefs):
    """-> (cov[3,3], info[3,3]) of DepthSensorModel::informationMatrixFromImageCoordinates"""
    c = np.ascontiguousarray(coefs, np.float64)
    cov = np.empty(9, np.float64); info = np.empty(9, np.float64)
    lib().orc_information_matrix(C.c_double(u), C.c_double(v), C.c_double(z), C.c_double(fx), C.c_double(fy),
                                 C.c_double(cx), C.c_double(cy), C.c_double(varU), C.c_double(varV),
                                 _p(c, C.c_double), _p(cov, C.c_double), _p(info, C.c_double))
    return cov.reshape(3, 3), info.reshape(3, 3)


def compute_normal(depth, u, v, fx, fy, cx, cy, depth_scale):
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    n = np.empty(3, np.float64)
    lib().orc_compute_normal(_p(depth, C.c_uint16), W, H, W, int(u), int(v), C.c_float(fx), C.c_float(fy), C.c_float(cx),
                             C.c_float(cy), C.c_double(depth_scale), _p(n, C.c_double))
    return n


def uncertainty_from_normal(normal, scale):
    nrm = np.ascontiguousarray(normal, np.float64)
    cov = np.empty(9, np.float64)
    lib().orc_uncertainty_from_normal(_p(nrm, C.c_double), C.c_double(scale), _p(cov, C.c_double))
    return cov.reshape(3, 3)


def inverse3d(m):
    m = np.ascontiguousarray(m, np.float64).reshape(9)
    r = np.empty(9, np.float64)
    lib().orc_inverse3d(_p(m, C.c_double), _p(r, C.c_double))
    return r.reshape(3, 3)


def compute_rgb_gradient(rgb, depth, u, v, fx, fy, cx, cy, depth_scale):
    """RGBD::computeRGBGradient on an H x W x 3 uint8 image (read as uint16 words like the reference)"""
    rgb = np.ascontiguousarray(rgb, np.uint8)
    depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    assert rgb.shape == (H, W, 3)
    g = np.empty(3, np.float64)
    lib().orc_compute_rgb_gradient(_p(rgb, C.c_uint8), 3 * W, _p(depth, C.c_uint16), W, H, W, int(u), int(v),
                                   C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy), C.c_double(depth_scale),
                                   _p(g, C.c_double))
    return g


def gradient_diag_table():
    t = np.empty(16, np.int32)
    lib().orc_gradient_diag_table(_p(t, C.c_int))
    return t.reshape(4, 4)


def uncertainty_from_gradient(grad, scale):
    g = np.ascontiguousarray(grad, np.float64)
    cov = np.empty(9, np.float64)
    lib().orc_uncertainty_from_gradient(_p(g, C.c_double), C.c_double(scale), _p(cov, C.c_double))
    return cov.reshape(3, 3)


def map_prepare(xyz, view_axis, pose, fx, fy, cx, cy, img_w=640, img_h=480, max_angle=0.6, max_z=5.0):
    """-> (kept int32[n], xyz_local f64[n,3], uv f64[n,2], angles f64[n]); pose = 4x4 camera->global (numpy row-major)"""
    xyz = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    va = np.ascontiguousarray(view_axis, np.float32).reshape(-1, 3)
    pcm = np.ascontiguousarray(np.asarray(pose, np.float64).T)      # column-major
    M = xyz.shape[0]
    kept = np.empty(max(1, M), np.int32); xl = np.empty((max(1, M), 3), np.float64)
    uv = np.empty((max(1, M), 2), np.float64); ang = np.empty(max(1, M), np.float64)
    n = lib().orc_map_prepare(_p(xyz, C.c_double), _p(va, C.c_float), M, _p(pcm, C.c_double), C.c_double(fx), C.c_double(fy),
                              C.c_double(cx), C.c_double(cy), C.c_double(img_w), C.c_double(img_h), C.c_double(max_angle),
                              C.c_double(max_z), _p(kept, C.c_int), _p(xl, C.c_double), _p(uv, C.c_double), _p(ang, C.c_double))
    return kept[:n].copy(), xl[:n].copy(), uv[:n].copy(), ang[:n].copy()


def svd3f(A):
    A = np.ascontiguousarray(A, np.float32)
    U = np.empty((3, 3), np.float32); S = np.empty(3, np.float32); V = np.empty((3, 3), np.float32)
    lib().orc_svd3f(_p(A, C.c_float), _p(U, C.c_float), _p(S, C.c_float), _p(V, C.c_float))
    return U, S, V


def svd3d(A):
    A = np.ascontiguousarray(A, np.float64)
    U = np.empty((3, 3), np.float64); S = np.empty(3, np.float64); V = np.empty((3, 3), np.float64)
    lib().orc_svd3d(_p(A, C.c_double), _p(U, C.c_double), _p(S, C.c_double), _p(V, C.c_double))
    return U, S, V


def umeyama(src, dst):
    """-> (ok, T float32[4,4]) with dst ~= R src + t"""
    src = np.ascontiguousarray(src, np.float32); dst = np.ascontiguousarray(dst, np.float32)
    T = np.empty((4, 4), np.float32)
    ok = lib().orc_umeyama(_p(src, C.c_float), _p(dst, C.c_float), src.shape[0], _p(T, C.c_float))
    return ok, T


def kabsch(A, B):
    """-> T float64[3,4] with B ~= R A + t"""
    A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
    T = np.empty((3, 4), np.float64)
    lib().orc_kabsch(_p(A, C.c_double), _p(B, C.c_double), A.shape[0], _p(T, C.c_double))
    return T


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, np.uint32); k = np.ascontiguousarray(key, np.uint32)
    out = np.empty(4, np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def sample3(seed, h, m):
    out = np.empty(3, np.int32)
    lib().orc_sample3(C.c_uint64(seed), C.c_uint32(h), m, _p(out, C.c_int))
    return out


def ransac_iterations(w):
    return lib().orc_ransac_iterations(float(w))


def inverse4(m):
    m = np.ascontiguousarray(m, np.float32)
    r = np.empty((4, 4), np.float32)
    lib().orc_inverse4(_p(m, C.c_float), _p(r, C.c_float))
    return r


def ransac(prev, cur, mq, mt, params=None, seed=0, num_hyp=0, want_counts=False, counts_cap=None):
    """-> dict(T[4,4] f32, inliers int32[], best_ratio, hyp_used, counts)"""
    prev = np.ascontiguousarray(prev, np.float32).reshape(-1, 3)
    cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 3)
    mq = np.ascontiguousarray(mq, np.int32); mt = np.ascontiguousarray(mt, np.int32)
    m = mq.size
    if params is None:
        params = default_ransac_params()
    T = np.empty((4, 4), np.float32)
    inl = np.empty(max(1, m), np.int32)
    n_inl = C.c_int(0); best = C.c_double(0); used = C.c_int(0)
    cap = (counts_cap or max(num_hyp, 487)) if want_counts else 0
    counts = np.full(max(1, cap), -2, np.int32)
    lib().orc_ransac(_p(prev, C.c_float), prev.shape[0], _p(cur, C.c_float), cur.shape[0],
                     _p(mq, C.c_int), _p(mt, C.c_int), m, C.byref(params), C.c_uint64(seed), num_hyp,
                     _p(T, C.c_float), _p(inl, C.c_int), C.byref(n_inl), C.byref(best), C.byref(used),
                     _p(counts, C.c_int) if want_counts else None, cap)
    return dict(T=T, inliers=inl[:n_inl.value].copy(), best_ratio=best.value, hyp_used=used.value,
                counts=counts[:cap] if want_counts else None)


def ransac_fixed_mt(prev, cur, mq, mt, params=None, seed=0, num_hyp=4096, threads=0):
    """all-cores (OpenMP over hypotheses) fixed-H RANSAC; same answer as ransac(..., num_hyp=H) -> dict(T, inliers, best_ratio)"""
    prev = np.ascontiguousarray(prev, np.float32).reshape(-1, 3)
    cur = np.ascontiguousarray(cur, np.float32).reshape(-1, 3)
    mq = np.ascontiguousarray(mq, np.int32); mt = np.ascontiguousarray(mt, np.int32)
    m = mq.size
    if params is None:
        params = default_ransac_params()
    T = np.empty((4, 4), np.float32)
    inl = np.empty(max(1, m), np.int32)
    n_inl = C.c_int(0); best = C.c_double(0)
    lib().orc_ransac_fixed_mt(_p(prev, C.c_float), prev.shape[0], _p(cur, C.c_float), cur.shape[0], _p(mq, C.c_int), _p(mt, C.c_int),
                              m, C.byref(params), C.c_uint64(seed), int(num_hyp), int(threads or num_threads()), _p(T, C.c_float),
                              _p(inl, C.c_int), C.byref(n_inl), C.byref(best))
    return dict(T=T, inliers=inl[:n_inl.value].copy(), best_ratio=best.value)


def point_inlier_ratio(inl_t, all_t, n_train):
    inl_t = np.ascontiguousarray(inl_t, np.int32); all_t = np.ascontiguousarray(all_t, np.int32)
    return lib().orc_point_inlier_ratio(_p(inl_t, C.c_int), inl_t.size, _p(all_t, C.c_int), all_t.size, n_train)


def set_stopping(rule, conf=0.99):
    lib().orc_set_stopping(int(rule), C.c_double(conf))


def num_threads():
    return lib().orc_num_threads()
