"""ctypes front for oracle/_ref/libref_frontend*.so -- the REFERENCE's own front-end sources (RANSAC.cpp, RGBD.cpp,
kabschEst.cpp, depthSensorModel.cpp, matcher.cpp, dbscan.cpp) compiled where they lie under /root/reference against the
Eigen / OpenCV stand-ins of oracle/ref_shim (`make -C oracle ref`; entry points in oracle/ref_frontend_wrap.cpp).

TEST INFRASTRUCTURE ONLY: tests/ use it to pin the CPU oracle (and, on the GPU box, the CUDA path) to the reference's
compiled scalar code.  Never imported by putslam_b200/ or adapter/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VARIANTS = ("", "_seq", "_scalelhs", "_jac32", "_sweep", "_f64", "_allseq")


class RansacArgs(C.Structure):
    _fields_ = [("error_version", C.c_int), ("thr_e", C.c_double), ("thr_r", C.c_double), ("min_ratio", C.c_double),
                ("min_matches", C.c_int), ("used_pairs", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class SensorArgs(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("varU", C.c_double),
                ("varV", C.c_double), ("c", C.c_double * 4), ("img_w", C.c_int), ("img_h", C.c_int),
                ("scale_normal", C.c_double), ("scale_gradient", C.c_double)]


class MatcherArgs(C.Structure):
    _fields_ = [("ransac", RansacArgs), ("dist", C.c_float * 5), ("radius", C.c_double), ("accept_ratio", C.c_double),
                ("dbscan_eps", C.c_double), ("min_reproj_dist", C.c_double), ("min_euclid_dist", C.c_double),
                ("remove_too_close", C.c_int), ("minimal_tracked_features", C.c_int), ("ransac_verbose", C.c_int)]


def ransac_args(error_version=0, thr_e=0.04, thr_r=2.0, min_ratio=0.2, min_matches=15, fx=517.3, fy=516.5, cx=318.6, cy=255.3):
    return RansacArgs(error_version, thr_e, thr_r, min_ratio, min_matches, 3, fx, fy, cx, cy)


def from_oracle_params(p):
    return RansacArgs(p.error_version, p.inlier_threshold_euclidean, p.inlier_threshold_reprojection,
                      p.minimal_inlier_ratio_threshold, p.minimal_number_of_matches, p.used_pairs, p.fx, p.fy, p.cx, p.cy)


def matcher_args(ransac=None, dist=(0, 0, 0, 0, 0), radius=0.12, accept_ratio=0.55, dbscan_eps=0.0, min_reproj=3.0,
                 min_euclid=0.0, remove_too_close=0, minimal_tracked=0):
    return MatcherArgs(ransac or ransac_args(), (C.c_float * 5)(*dist), radius, accept_ratio, dbscan_eps, min_reproj, min_euclid,
                       remove_too_close, minimal_tracked, 0)


def sensor_args(fx=517.3, fy=516.5, cx=318.6, cy=255.3, varU=1.1046, varV=0.6416, coefs=(0.0, 0.002797, -0.004249, 0.007311),
                img_w=640, img_h=480, scale_normal=0.8, scale_gradient=0.8):
    return SensorArgs(fx, fy, cx, cy, varU, varV, (C.c_double * 4)(*coefs), img_w, img_h, scale_normal, scale_gradient)


def available(variant=""):
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_frontend{variant}.so"))


_libs = {}


def lib(variant=""):
    if variant not in _libs:
        L = C.CDLL(os.path.join(_HERE, "_ref", f"libref_frontend{variant}.so"))
        L.ref_point_inlier_ratio.restype = C.c_double
        L.ref_shim_model.restype = C.c_char_p
        L.ref_round_size.argtypes = [C.c_double, C.c_int]
        _libs[variant] = L
    return _libs[variant]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a, w):
    return np.ascontiguousarray(a, np.float32).reshape(-1, w)


def _f64(a, w):
    return np.ascontiguousarray(a, np.float64).reshape(-1, w)


def shim_model(variant=""):
    return lib(variant).ref_shim_model().decode()


def ransac(prev, cur, mq, mt, args=None, seed=0, variant=""):
    """RANSAC::estimateTransformation -> dict(T, inliers (indices into the match list), best_ratio_pct, hyp_used, filtered)"""
    prev, cur = _f32(prev, 3), _f32(cur, 3)
    mq = np.ascontiguousarray(mq, np.int32); mt = np.ascontiguousarray(mt, np.int32)
    m = mq.size
    args = args or ransac_args()
    T = np.empty((4, 4), np.float32); inl = np.empty(max(1, m), np.int32)
    n_inl = C.c_int(0); best = C.c_double(0); used = C.c_int(0)
    mf = lib(variant).ref_ransac(_p(prev, C.c_float), prev.shape[0], _p(cur, C.c_float), cur.shape[0], _p(mq, C.c_int), _p(mt, C.c_int),
                                 m, C.byref(args), C.c_uint64(seed), _p(T, C.c_float), _p(inl, C.c_int), C.byref(n_inl),
                                 C.byref(best), C.byref(used))
    return dict(T=T, inliers=inl[:n_inl.value].copy(), best_ratio_pct=best.value, hyp_used=used.value, filtered=mf)


def point_inlier_ratio(inl_t, all_t):
    a = np.ascontiguousarray(inl_t, np.int32); b = np.ascontiguousarray(all_t, np.int32)
    return lib().ref_point_inlier_ratio(_p(a, C.c_int), a.size, _p(b, C.c_int), b.size)


def backproject(uv, depth, k4=(517.3, 516.5, 318.6, 255.3), dist5=None, depth_scale=5000.0):
    """removeImageDistortion (optional) + keypoints2Dto3D + the detDist expression -> (uv_used, xyz f32, det_dist f64)"""
    uv = _f32(uv, 2); depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    n = uv.shape[0]
    k = np.asarray(k4, np.float32)
    d = None if dist5 is None else np.asarray(dist5, np.float32)
    used = np.empty((n, 2), np.float32); xyz = np.empty((n, 3), np.float32); dd = np.empty(n, np.float64)
    lib().ref_backproject(_p(uv, C.c_float), n, _p(depth, C.c_uint16), W, H, _p(k, C.c_float), None if d is None else _p(d, C.c_float),
                          C.c_double(depth_scale), _p(used, C.c_float), _p(xyz, C.c_float), _p(dd, C.c_double))
    return used, xyz, dd


def round_size(x, size):
    return lib().ref_round_size(float(x), int(size))


def point3Dto2D(xyz, k4=(517.3, 516.5, 318.6, 255.3)):
    xyz = _f32(xyz, 3); k = np.asarray(k4, np.float32)
    uv = np.empty((xyz.shape[0], 2), np.float32)
    lib().ref_point3Dto2D(_p(xyz, C.c_float), xyz.shape[0], _p(k, C.c_float), _p(uv, C.c_float))
    return uv


def compute_cov(u, v, depth, s=None):
    s = s or sensor_args(); out = np.empty(9, np.float64)
    lib().ref_compute_cov(C.c_uint(int(u)), C.c_uint(int(v)), C.c_double(depth), C.byref(s), _p(out, C.c_double))
    return out.reshape(3, 3)


def information_matrix_uvz(u, v, z, s=None):
    s = s or sensor_args(); out = np.empty(9, np.float64)
    lib().ref_information_matrix_uvz(C.c_double(u), C.c_double(v), C.c_double(z), C.byref(s), _p(out, C.c_double))
    return out.reshape(3, 3)


def information_matrix_xyz(x, y, z, s=None):
    s = s or sensor_args(); out = np.empty(9, np.float64)
    lib().ref_information_matrix_xyz(C.c_double(x), C.c_double(y), C.c_double(z), C.byref(s), _p(out, C.c_double))
    return out.reshape(3, 3)


def inverse_model(x, y, z, s=None):
    s = s or sensor_args(); out = np.empty(3, np.float64)
    lib().ref_inverse_model(C.c_double(x), C.c_double(y), C.c_double(z), C.byref(s), _p(out, C.c_double))
    return out


def uncertainty_from_normal(n, s=None):
    s = s or sensor_args(); n = np.ascontiguousarray(n, np.float64); out = np.empty(9, np.float64)
    lib().ref_uncertainty_from_normal(_p(n, C.c_double), C.byref(s), _p(out, C.c_double))
    return out.reshape(3, 3)


def uncertainty_from_gradient(g, s=None):
    s = s or sensor_args(); g = np.ascontiguousarray(g, np.float64); out = np.empty(9, np.float64)
    lib().ref_uncertainty_from_gradient(_p(g, C.c_double), C.byref(s), _p(out, C.c_double))
    return out.reshape(3, 3)


def compute_normal(depth, u, v, k4=(517.3, 516.5, 318.6, 255.3), depth_scale=5000.0):
    depth = np.ascontiguousarray(depth, np.uint16); H, W = depth.shape
    k = np.asarray(k4, np.float32); out = np.empty(3, np.float64)
    lib().ref_compute_normal(_p(depth, C.c_uint16), W, H, int(u), int(v), _p(k, C.c_float), C.c_double(depth_scale), _p(out, C.c_double))
    return out


def compute_rgb_gradient(rgb, depth, u, v, k4=(517.3, 516.5, 318.6, 255.3), depth_scale=5000.0):
    rgb = np.ascontiguousarray(rgb, np.uint8); depth = np.ascontiguousarray(depth, np.uint16); H, W = depth.shape
    k = np.asarray(k4, np.float32); out = np.empty(3, np.float64)
    lib().ref_compute_rgb_gradient(_p(rgb, C.c_uint8), 3 * W, _p(depth, C.c_uint16), W, H, int(u), int(v), _p(k, C.c_float),
                                   C.c_double(depth_scale), _p(out, C.c_double))
    return out


def kabsch(A, B, variant=""):
    A, B = _f64(A, 3), _f64(B, 3); T = np.empty((3, 4), np.float64)
    lib(variant).ref_kabsch(_p(A, C.c_double), _p(B, C.c_double), A.shape[0], _p(T, C.c_double))
    return T


def transform_uncertainty(A, B, CA, CB, T, mode=0):
    A, B = _f64(A, 3), _f64(B, 3); CA, CB = _f64(CA, 9), _f64(CB, 9)
    T12 = np.ascontiguousarray(np.asarray(T, np.float64)[:3, :4]); U = np.empty((6, 6), np.float64)
    lib().ref_transform_uncertainty(_p(A, C.c_double), _p(B, C.c_double), _p(CA, C.c_double), _p(CB, C.c_double), A.shape[0],
                                    _p(T12, C.c_double), int(mode), _p(U, C.c_double))
    return U


def match_xyz(map_xyz, map_desc, map_octave, map_detdist, cur_xyz, cur_desc, cur_octave, cur_detdist, args=None,
              computation_number=1, seed=0, use_frame_ids=False, variant=""):
    """Matcher::matchXYZ (guided matching + RANSAC) -> dict(T, pairs (map, cur), ratio, n_matches, n_perfect, hyp_used)"""
    mx = _f64(map_xyz, 3); md = np.ascontiguousarray(map_desc, np.uint8).reshape(-1, 32)
    mo = np.ascontiguousarray(map_octave, np.int32); mdd = np.ascontiguousarray(map_detdist, np.float64)
    cx = _f32(cur_xyz, 3); cd = np.ascontiguousarray(cur_desc, np.uint8).reshape(-1, 32)
    co = np.ascontiguousarray(cur_octave, np.int32); cdd = np.ascontiguousarray(cur_detdist, np.float64)
    M, N = mx.shape[0], cx.shape[0]
    args = args or matcher_args()
    cap = max(1, M * N)
    T = np.empty((4, 4), np.float32); pm = np.empty(cap, np.int32); pc = np.empty(cap, np.int32)
    ratio = C.c_double(0); nm = C.c_int(0); npf = C.c_int(0); used = C.c_int(0)
    n = lib(variant).ref_match_xyz(_p(mx, C.c_double), _p(md, C.c_uint8), _p(mo, C.c_int), _p(mdd, C.c_double), M,
                                   _p(cx, C.c_float), _p(cd, C.c_uint8), _p(co, C.c_int), _p(cdd, C.c_double), N, C.byref(args),
                                   int(computation_number), C.c_uint64(seed), int(use_frame_ids), _p(T, C.c_float), _p(pm, C.c_int),
                                   _p(pc, C.c_int), C.byref(ratio), C.byref(nm), C.byref(npf), C.byref(used))
    return dict(T=T, pairs=np.stack([pm[:n], pc[:n]], 1).copy(), ratio=ratio.value, n_matches=nm.value, n_perfect=npf.value,
                hyp_used=used.value)


def match_vo(prev_desc, prev_xyz, cur_kp_xy, cur_octave, cur_desc, depth, args=None, depth_scale=5000.0, seed=0, variant=""):
    """Matcher::match (one VO step with a scripted detector) -> dict(T, inliers (q, t), ratio, kept, xyz, uv, hyp_used)"""
    pd = np.ascontiguousarray(prev_desc, np.uint8).reshape(-1, 32); px = _f32(prev_xyz, 3)
    kp = _f32(cur_kp_xy, 2); co = np.ascontiguousarray(cur_octave, np.int32)
    cd = np.ascontiguousarray(cur_desc, np.uint8).reshape(-1, 32); depth = np.ascontiguousarray(depth, np.uint16)
    H, W = depth.shape
    n_prev, n_cur = pd.shape[0], kp.shape[0]
    args = args or matcher_args()
    cap = max(1, min(n_prev, n_cur))
    T = np.empty((4, 4), np.float32); iq = np.empty(cap, np.int32); it = np.empty(cap, np.int32)
    kept = np.empty(max(1, n_cur), np.int32); kxyz = np.empty((max(1, n_cur), 3), np.float32); kuv = np.empty((max(1, n_cur), 2), np.float32)
    ratio = C.c_double(0); nk = C.c_int(0); used = C.c_int(0)
    n = lib(variant).ref_match_vo(_p(pd, C.c_uint8), _p(px, C.c_float), n_prev, _p(kp, C.c_float), _p(co, C.c_int), _p(cd, C.c_uint8),
                                  n_cur, _p(depth, C.c_uint16), W, H, C.c_double(depth_scale), C.byref(args), C.c_uint64(seed),
                                  _p(T, C.c_float), _p(iq, C.c_int), _p(it, C.c_int), C.byref(ratio), _p(kept, C.c_int),
                                  _p(kxyz, C.c_float), _p(kuv, C.c_float), C.byref(nk), C.byref(used))
    k = nk.value
    return dict(T=T, inliers=np.stack([iq[:n], it[:n]], 1).copy(), ratio=ratio.value, kept=kept[:k].copy(), xyz=kxyz[:k].copy(),
                uv=kuv[:k].copy(), hyp_used=used.value)


def loop_closure(desc0, xyz0, desc1, xyz1, args=None, seed=0, variant=""):
    """Matcher::matchFeatureLoopClosure -> dict(T, pairs, ret, hyp_used)"""
    d0 = np.ascontiguousarray(desc0, np.uint8).reshape(-1, 32); d1 = np.ascontiguousarray(desc1, np.uint8).reshape(-1, 32)
    x0, x1 = _f64(xyz0, 3), _f64(xyz1, 3)
    args = args or matcher_args()
    cap = max(1, min(d0.shape[0], d1.shape[0]))
    T = np.empty((4, 4), np.float32); p0 = np.empty(cap, np.int32); p1 = np.empty(cap, np.int32)
    ret = C.c_double(0); used = C.c_int(0)
    n = lib(variant).ref_loop_closure(_p(d0, C.c_uint8), _p(x0, C.c_double), d0.shape[0], _p(d1, C.c_uint8), _p(x1, C.c_double),
                                      d1.shape[0], C.byref(args), C.c_uint64(seed), _p(T, C.c_float), _p(p0, C.c_int), _p(p1, C.c_int),
                                      C.byref(ret), C.byref(used))
    return dict(T=T, pairs=np.stack([p0[:n], p1[:n]], 1).copy(), ret=ret.value, hyp_used=used.value)


def remove_too_close(dist_xy, undist_xy, xyz, mq, mt, min_euclid, min_reproj):
    """Matcher::removeTooCloseFeatures -> (kept indices, mq, mt)"""
    d, u, p = _f32(dist_xy, 2), _f32(undist_xy, 2), _f32(xyz, 3)
    mq = np.array(mq, np.int32); mt = np.array(mt, np.int32)
    n = d.shape[0]; m = C.c_int(mq.size)
    kept = np.empty(max(1, n), np.int32)
    k = lib().ref_remove_too_close(_p(d, C.c_float), _p(u, C.c_float), _p(p, C.c_float), n, _p(mq, C.c_int), _p(mt, C.c_int), C.byref(m),
                                   C.c_double(min_euclid), C.c_double(min_reproj), _p(kept, C.c_int))
    return kept[:k].copy(), mq[:m.value].copy(), mt[:m.value].copy()


def merge_tracked(undist_xy, sandbox_undist_xy, min_reproj):
    """Matcher::mergeTrackedFeatures -> indices of the sandbox features that get appended, in order"""
    u, s = _f32(undist_xy, 2), _f32(sandbox_undist_xy, 2)
    added = np.empty(max(1, s.shape[0]), np.int32)
    k = lib().ref_merge_tracked(_p(u, C.c_float), u.shape[0], _p(s, C.c_float), s.shape[0], C.c_double(min_reproj), _p(added, C.c_int))
    return added[:k].copy()


# ---- libref_tree.so: the reference's Matcher driving adapter/putslam_tree's MatcherB200 (needs a GPU to run) ------------
class TreeArgs(C.Structure):
    _fields_ = [("vo_tracking", C.c_int), ("error_version", C.c_int), ("thr_e", C.c_double), ("thr_r", C.c_double),
                ("min_ratio", C.c_double), ("min_matches", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("dist", C.c_float * 5), ("grid_cols", C.c_int), ("grid_rows", C.c_int),
                ("maximal_tracked_features", C.c_int), ("minimal_tracked_features", C.c_int), ("dbscan_eps", C.c_double),
                ("win_size", C.c_int), ("max_levels", C.c_int), ("max_iter", C.c_int), ("eps", C.c_float),
                ("tracking_error_threshold", C.c_double), ("tracking_min_eig_threshold", C.c_double),
                ("min_reproj_dist", C.c_double), ("min_euclid_dist", C.c_double), ("remove_too_close", C.c_int)]


def tree_args(vo_tracking=0, error_version=0, dist=(0, 0, 0, 0, 0), dbscan_eps=10.0, maximal_tracked=500, minimal_tracked=0):
    """defaults = resources/putslammatcherOpenCVParameters.xml (RANSAC 0.04 / 2.0 / 0.2 / 15, grid 1 x 1, window 7, 3 levels,
    30 iterations / 0.01, tracking thresholds 25 / 0 / 3 px)"""
    return TreeArgs(vo_tracking, error_version, 0.04, 2.0, 0.2, 15, 517.3, 516.5, 318.6, 255.3, (C.c_float * 5)(*dist), 1, 1,
                    maximal_tracked, minimal_tracked, dbscan_eps, 7, 3, 30, 0.01, 25.0, 0.0, 3.0, 0.0, 0)


def tree_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_tree.so"))


_tree = None


def tree_run_vo(rgb0, depth0, rgb1, depth1, args=None, depth_scale=5000.0, seed=0, cap=4096):
    """Matcher::detectInitFeatures(frame 0) + Matcher::runVO(frame 1) of the reference on a MatcherB200
    -> dict(T, inliers (q, t), ratio, kp0 (x, y, octave), kp1, xyz1, hyp_used)"""
    global _tree
    if _tree is None:
        _tree = C.CDLL(os.path.join(_HERE, "_ref", "libref_tree.so"))
    rgb0 = np.ascontiguousarray(rgb0, np.uint8); rgb1 = np.ascontiguousarray(rgb1, np.uint8)
    depth0 = np.ascontiguousarray(depth0, np.uint16); depth1 = np.ascontiguousarray(depth1, np.uint16)
    H, W = depth0.shape
    ch = 1 if rgb0.ndim == 2 else rgb0.shape[2]
    args = args or tree_args()
    T = np.empty((4, 4), np.float32); iq = np.empty(cap, np.int32); it = np.empty(cap, np.int32)
    kp0 = np.empty((cap, 3), np.float32); kp1 = np.empty((cap, 3), np.float32); xyz1 = np.empty((cap, 3), np.float32)
    ratio = C.c_double(0); n0 = C.c_int(0); n1 = C.c_int(0); used = C.c_int(0)
    n = _tree.tree_run_vo(_p(rgb0, C.c_uint8), _p(depth0, C.c_uint16), _p(rgb1, C.c_uint8), _p(depth1, C.c_uint16), W, H, ch,
                          C.c_double(depth_scale), C.byref(args), C.c_uint64(seed), _p(T, C.c_float), _p(iq, C.c_int), _p(it, C.c_int),
                          C.byref(ratio), cap, _p(kp0, C.c_float), C.byref(n0), _p(kp1, C.c_float), _p(xyz1, C.c_float), C.byref(n1),
                          C.byref(used))
    return dict(T=T, inliers=np.stack([iq[:n], it[:n]], 1).copy(), ratio=ratio.value, kp0=kp0[:n0.value].copy(),
                kp1=kp1[:n1.value].copy(), xyz1=xyz1[:n1.value].copy(), hyp_used=used.value)
