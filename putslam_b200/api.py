"""ctypes binding of libpslam_b200.so (include/pslam_b200.h) with numpy-facing wrappers.

This is host-side plumbing for the tests and bench.py; the product is the C ABI + adapter/ C++ classes.
There is no CPU path here: if the shared library is missing or no sm_100 GPU is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpslam_b200.so")

PSLAM_OK = 0
ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED, ERR_NCCL, ERR_NO_DEVICE = -1, -2, -3, -4, -5, -6

# every symbol include/pslam_b200.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = [
    "pslam_ctx_create", "pslam_ctx_destroy", "pslam_last_error", "pslam_version", "pslam_ctx_stream",
    "pslam_ctx_sync", "pslam_kernel_launches", "pslam_sm_count", "pslam_backproject", "pslam_information_matrices", "pslam_normal_uncertainty", "pslam_gradient_uncertainty", "pslam_orb_describe", "pslam_orb_detect", "pslam_fast_detect", "pslam_klt_track", "pslam_klt_perform_tracking", "pslam_klt_frame", "pslam_transform_uncertainty_batch", "pslam_match_bf_mutual",
    "pslam_match_knn2", "pslam_match_guided_xyz", "pslam_ransac_estimate", "pslam_ransac_set_stopping", "pslam_ransac_last_counts",
    "pslam_ransac_sample", "pslam_point_inlier_ratio", "pslam_kabsch_batch", "pslam_frame_to_frame",
    "pslam_frame_to_map", "pslam_frame_to_map_features", "pslam_map_prepare", "pslam_map_reserve", "pslam_map_write", "pslam_map_truncate", "pslam_map_size", "pslam_frame_to_resident_map", "pslam_loop_closure_pair", "pslam_frame_to_map_resident", "pslam_frame_to_frame_resident", "pslam_lc_db_reserve",
    "pslam_lc_db_append", "pslam_lc_db_clear", "pslam_lc_db_size", "pslam_lc_set_id_base", "pslam_lc_set_work_unit", "pslam_lc_query",
    "pslam_lc_query_resident", "pslam_lc_last_sweep_ms", "pslam_lc_exchange_mode", "pslam_lc_tensor_status", "pslam_debug_host_stamps", "pslam_host_register", "pslam_host_unregister", "pslam_comm_unique_id", "pslam_comm_init", "pslam_comm_destroy",
    "pslam_lc_query_sharded", "pslam_lc_query_sharded_resident", "pslam_lc_query_sharded_resident_bcast", "pslam_lc_knn2", "pslam_lc_set_desc_base",
    "pslam_lc_knn2_sharded", "pslam_lc_knn2_resident",
]


class PslamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pslam error {code}: {msg}")
        self.code = code


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("dist", C.c_float * 5)]


class CovParams(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("var_u", C.c_double), ("var_v", C.c_double), ("dist_var_coefs", C.c_double * 4)]


class RansacParams(C.Structure):
    _fields_ = [("error_version", C.c_int), ("inlier_threshold_euclidean", C.c_double),
                ("inlier_threshold_reprojection", C.c_double), ("minimal_inlier_ratio_threshold", C.c_double),
                ("minimal_number_of_matches", C.c_int), ("used_pairs", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class MapPrepareParams(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("image_w", C.c_double), ("image_h", C.c_double), ("max_angle", C.c_double), ("max_z", C.c_double)]


class FrameResult(C.Structure):
    _fields_ = [("n_matches", C.c_int), ("n_inliers", C.c_int), ("hyp_used", C.c_int), ("n_filtered", C.c_int),
                ("best_ratio", C.c_double), ("inlier_ratio", C.c_double), ("T", C.c_float * 16)]


def default_ransac_params(error_version=0):
    """resources/putslammatcherOpenCVParameters.xml:29-37 defaults + freiburg1 intrinsics."""
    return RansacParams(error_version, 0.04, 2.0, 0.2, 15, 3, 517.3, 516.5, 318.6, 255.3)


def make_camera(fx=517.3, fy=516.5, cx=318.6, cy=255.3, dist=(-0.0410, 0.3286, 0.0087, 0.0051, -0.5643)):
    return Camera(fx, fy, cx, cy, (C.c_float * 5)(*dist))


_lib = None


def load_library():
    """Load libpslam_b200.so; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PslamError(ERR_NO_DEVICE, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    lib.pslam_last_error.restype = C.c_char_p
    lib.pslam_ctx_stream.restype = C.c_void_p
    lib.pslam_kernel_launches.restype = C.c_uint64
    lib.pslam_point_inlier_ratio.restype = C.c_double
    lib.pslam_ctx_destroy.restype = None
    lib.pslam_ransac_sample.restype = None
    for name in ("pslam_ctx_destroy", "pslam_last_error", "pslam_ctx_stream", "pslam_ctx_sync",
                 "pslam_kernel_launches", "pslam_sm_count"):
        getattr(lib, name).argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _arr(a, dt, shape_last=None):
    if a is None:
        return np.zeros((0,), dt)
    a = np.ascontiguousarray(a, dtype=dt)
    if shape_last is not None and a.size:
        a = a.reshape(-1, shape_last)
    return a


class Context:
    """One pslam_ctx (one CUDA stream, one staging arena) -- one per reference Matcher instance."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        r = self.lib.pslam_ctx_create(int(device), C.byref(h))
        if r != PSLAM_OK:
            raise PslamError(r, "pslam_ctx_create failed (no usable sm_100 GPU; there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.pslam_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r, allow=()):
        if r != PSLAM_OK and r not in allow:
            raise PslamError(r, (self.lib.pslam_last_error(self.h) or b"").decode())
        return r

    @property
    def stream(self):
        return self.lib.pslam_ctx_stream(self.h)

    @property
    def launches(self):
        return int(self.lib.pslam_kernel_launches(self.h))

    @property
    def sm_count(self):
        return int(self.lib.pslam_sm_count(self.h))

    def sync(self):
        self._ck(self.lib.pslam_ctx_sync(self.h))

    # ---- stage 1 ----
    def backproject(self, uv, depth, cam=None, undistort=False, depth_scale=5000.0, cov=None):
        uv = _arr(uv, np.float32, 2)
        depth = np.ascontiguousarray(depth, np.uint16)
        H, W = depth.shape
        n = uv.shape[0] if uv.size else 0
        cam = cam or make_camera()
        und = np.empty((n, 2), np.float32); xyz = np.empty((n, 3), np.float32); dd = np.empty(n, np.float64)
        covo = np.empty((n, 3, 3), np.float64) if cov is not None else None
        self._ck(self.lib.pslam_backproject(self.h, _p(uv, C.c_float), n, _p(depth, C.c_uint16), W, H, W, C.byref(cam),
                                            int(bool(undistort)), C.c_double(depth_scale), _p(und, C.c_float),
                                            _p(xyz, C.c_float), _p(dd, C.c_double), _p(covo, C.c_double),
                                            C.byref(cov) if cov is not None else None))
        return dict(xyz=xyz, uv_undist=und, det_dist=dd, cov=covo)

    def normal_uncertainty(self, px, depth, cam=None, depth_scale=5000.0, scale=0.8, with_info=False):
        px = _arr(px, np.int32, 2)
        depth = np.ascontiguousarray(depth, np.uint16)
        H, W = depth.shape
        n = px.shape[0] if px.size else 0
        cam = cam or make_camera()
        nrm = np.empty((n, 3), np.float64); cov = np.empty((n, 3, 3), np.float64)
        info = np.empty((n, 3, 3), np.float64) if with_info else None
        self._ck(self.lib.pslam_normal_uncertainty(self.h, _p(px, C.c_int), n, _p(depth, C.c_uint16), W, H, W, C.byref(cam),
                                                   C.c_double(depth_scale), C.c_double(scale), _p(nrm, C.c_double),
                                                   _p(cov, C.c_double), _p(info, C.c_double) if with_info else None))
        return (nrm, cov, info) if with_info else (nrm, cov)

    def gradient_uncertainty(self, px, rgb, depth, cam=None, depth_scale=5000.0, scale=0.8):
        """uncertainty model 2 -> (grad n x 3, cov n x 3 x 3, info n x 3 x 3); rgb = H x W x 3 uint8"""
        px = _arr(px, np.int32, 2)
        depth = np.ascontiguousarray(depth, np.uint16)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        H, W = depth.shape
        if rgb.shape != (H, W, 3):
            raise ValueError("rgb must be H x W x 3 uint8 with the depth image's size")
        n = px.shape[0] if px.size else 0
        cam = cam or make_camera()
        g = np.empty((n, 3), np.float64); cov = np.empty((n, 3, 3), np.float64); info = np.empty((n, 3, 3), np.float64)
        self._ck(self.lib.pslam_gradient_uncertainty(self.h, _p(px, C.c_int), n, _p(rgb, C.c_uint8), 3 * W,
                                                     _p(depth, C.c_uint16), W, H, W, C.byref(cam), C.c_double(depth_scale),
                                                     C.c_double(scale), _p(g, C.c_double), _p(cov, C.c_double),
                                                     _p(info, C.c_double)))
        return g, cov, info

    def information_matrices(self, uvz, cov):
        uvz = _arr(uvz, np.float64, 3)
        n = uvz.shape[0] if uvz.size else 0
        info = np.empty((n, 3, 3), np.float64); covo = np.empty((n, 3, 3), np.float64)
        self._ck(self.lib.pslam_information_matrices(self.h, _p(uvz, C.c_double), n, C.byref(cov), _p(info, C.c_double),
                                                     _p(covo, C.c_double)))
        return info, covo

    # ---- descriptor production ----
    def orb_describe(self, image, xy, octave, angle_deg, resident_shape=None):
        """cv::ORB::compute with provided keypoints -> (order int32[n_out] into the input keypoints, desc uint8[n_out, 32]).
        image: H x W uint8 (gray) or H x W x 3 uint8 (BGR); None + resident_shape=(H, W[, 3]) = the frame uploaded by the
        previous orb_detect / orb_describe call on this context."""
        if image is None:
            img = None
            ch = 3 if len(resident_shape) == 3 else 1
            H, W = resident_shape[:2]
        else:
            img = np.ascontiguousarray(image, np.uint8)
            ch = 3 if img.ndim == 3 else 1
            H, W = img.shape[:2]
        xy = _arr(xy, np.float32, 2); oc = _arr(octave, np.int32); an = _arr(angle_deg, np.float32)
        n = oc.size
        order = np.empty(max(1, n), np.int32); desc = np.empty((max(1, n), 32), np.uint8); n_out = C.c_int(0)
        self._ck(self.lib.pslam_orb_describe(self.h, _p(img, C.c_uint8) if img is not None else None, W, H, ch * W, ch,
                                             _p(xy, C.c_float), _p(oc, C.c_int),
                                             _p(an, C.c_float), n, _p(order, C.c_int), C.byref(n_out), _p(desc, C.c_uint8)))
        k = n_out.value
        return order[:k].copy(), desc[:k].copy()

    def orb_detect(self, image, nfeatures=500, colour_order=0, cap=None):
        """cv::ORB::create(nfeatures).detect(image) -> dict(xy [n,2], size, angle, response float32[n], octave int32[n]),
        in OpenCV's output order.  image: H x W (gray) or H x W x 3 uint8; colour_order 0 = BGR2GRAY, 1 = RGB2GRAY."""
        img = np.ascontiguousarray(image, np.uint8)
        ch = 3 if img.ndim == 3 else 1
        H, W = img.shape[:2]
        cap = int(cap if cap is not None else max(16, 4 * nfeatures + 64))
        xy = np.empty((cap, 2), np.float32); size = np.empty(cap, np.float32); ang = np.empty(cap, np.float32)
        resp = np.empty(cap, np.float32); octv = np.empty(cap, np.int32); n = C.c_int(0)
        self._ck(self.lib.pslam_orb_detect(self.h, _p(img, C.c_uint8), W, H, ch * W, ch, int(colour_order), int(nfeatures),
                                           _p(xy, C.c_float), _p(size, C.c_float), _p(ang, C.c_float), _p(resp, C.c_float),
                                           _p(octv, C.c_int), cap, C.byref(n)))
        k = n.value
        return dict(xy=xy[:k].copy(), size=size[:k].copy(), angle=ang[:k].copy(), response=resp[:k].copy(), octave=octv[:k].copy())

    def fast_detect(self, image, threshold=10, colour_order=0, cap=200000):
        """cv::FastFeatureDetector::create(threshold, True).detect(image) -> (xy float32[n, 2], response float32[n]), raster order"""
        img = np.ascontiguousarray(image, np.uint8)
        ch = 3 if img.ndim == 3 else 1
        H, W = img.shape[:2]
        xy = np.empty((max(1, cap), 2), np.float32); resp = np.empty(max(1, cap), np.float32); n = C.c_int(0)
        self._ck(self.lib.pslam_fast_detect(self.h, _p(img, C.c_uint8), W, H, ch * W, ch, int(colour_order), int(threshold),
                                            _p(xy, C.c_float), _p(resp, C.c_float), cap, C.byref(n)))
        return xy[:n.value].copy(), resp[:n.value].copy()

    def transform_uncertainty(self, A_list, B_list, covA_list, covB_list, T_list, mode="euler"):
        """TransformEst::computeUncertainty ("euler") / computeUncertaintyG2O ("quat") for a batch: lists of n_i x 3 points,
        n_i x 3 x 3 covariances and 4 x 4 (or 3 x 4) transforms with A ~ R B + t -> (U [batch, 6, 6], ok [batch])"""
        batch = len(A_list)
        off = np.zeros(batch + 1, np.int32)
        for i, a in enumerate(A_list):
            off[i + 1] = off[i] + len(a)
        cat = lambda L, w: (np.concatenate([np.asarray(x, np.float64).reshape(-1, w) for x in L]) if off[-1] else np.zeros((0, w)))
        A = np.ascontiguousarray(cat(A_list, 3)); B = np.ascontiguousarray(cat(B_list, 3))
        CA = np.ascontiguousarray(cat(covA_list, 9)); CB = np.ascontiguousarray(cat(covB_list, 9))
        T = np.ascontiguousarray(np.stack([np.asarray(t, np.float64)[:3, :4].T.ravel() for t in T_list])) if batch else np.zeros((0, 12))
        U = np.zeros((max(batch, 1), 36), np.float64); ok = np.zeros(max(batch, 1), np.int32)
        self._ck(self.lib.pslam_transform_uncertainty_batch(self.h, _p(A, C.c_double), _p(B, C.c_double), _p(CA, C.c_double),
                                                            _p(CB, C.c_double), _p(off, C.c_int), _p(T, C.c_double), batch,
                                                            0 if mode == "euler" else 1, _p(U, C.c_double), _p(ok, C.c_int)))
        return U[:batch].reshape(batch, 6, 6).transpose(0, 2, 1).copy(), ok[:batch].copy()

    # ---- KLT tracking (performTracking seam) ----
    KLT_USE_INITIAL_FLOW, KLT_GET_MIN_EIGENVALS = 4, 8

    def klt_track(self, prev_image, cur_image, prev_xy, init_xy=None, win=7, max_level=3, max_iter=30, eps=0.01,
                  criteria_type=3, min_eig_err=False, min_eig_threshold=1e-4, prune=None):
        """cv::calcOpticalFlowPyrLK(prev_image, cur_image, prev_xy, ...) -> dict(xy float32[n, 2], status uint8[n],
        err float32[n]); prev_image None = the current frame of the previous klt call on this context.
        prune=(error_threshold, min_distance): the whole of MatcherOpenCV::performTracking, adds kept int32[m]."""
        cur = np.ascontiguousarray(cur_image, np.uint8)
        prev = None if prev_image is None else np.ascontiguousarray(prev_image, np.uint8)
        if prev is not None and prev.shape != cur.shape:
            raise ValueError("klt_track: frames differ in shape")
        ch = 3 if cur.ndim == 3 else 1
        H, W = cur.shape[:2]
        p = _arr(prev_xy, np.float32, 2); n = len(p)
        flags = (self.KLT_GET_MIN_EIGENVALS if min_eig_err else 0) | (self.KLT_USE_INITIAL_FLOW if init_xy is not None else 0)
        xy = np.zeros((max(n, 1), 2), np.float32)
        if init_xy is not None:
            xy[:n] = _arr(init_xy, np.float32, 2)
        st = np.zeros(max(n, 1), np.uint8); err = np.zeros(max(n, 1), np.float32)
        common = (self.h, _p(prev, C.c_uint8), _p(cur, C.c_uint8), W, H, ch * W, ch, _p(p, C.c_float), _p(xy, C.c_float), n,
                  int(win), int(max_level), int(criteria_type), int(max_iter), C.c_double(eps), flags, C.c_double(min_eig_threshold))
        if prune is None:
            self._ck(self.lib.pslam_klt_track(*common, _p(st, C.c_uint8), _p(err, C.c_float)))
            return dict(xy=xy[:n].copy(), status=st[:n].copy(), err=err[:n].copy())
        kept = np.zeros(max(n, 1), np.int32); m = C.c_int(0)
        self._ck(self.lib.pslam_klt_perform_tracking(*common, C.c_double(prune[0]), C.c_double(prune[1]), _p(st, C.c_uint8),
                                                     _p(err, C.c_float), _p(kept, C.c_int), C.byref(m)))
        return dict(xy=xy[:n].copy(), status=st[:n].copy(), err=err[:n].copy(), kept=kept[:m.value].copy())

    def klt_frame(self, prev_image, cur_image, prev_xy, prev_xyz, depth, cam=None, undistort=False, depth_scale=5000.0,
                  params=None, seed=0, num_hyp=0, win=7, max_level=3, max_iter=30, eps=0.01, criteria_type=3,
                  min_eig_err=False, min_eig_threshold=0.0, error_threshold=25.0, min_distance=3.0, init_xy=None):
        """the data-parallel part of Matcher::trackKLT in one submission: track -> threshold / too-close rule ->
        undistort + back-project the survivors -> RANSAC against prev_xyz.  Defaults = the reference's shipped parameters."""
        cur = np.ascontiguousarray(cur_image, np.uint8)
        prev = None if prev_image is None else np.ascontiguousarray(prev_image, np.uint8)
        ch = 3 if cur.ndim == 3 else 1
        H, W = cur.shape[:2]
        depth = np.ascontiguousarray(depth, np.uint16)
        if depth.shape != (H, W):
            raise ValueError("klt_frame: depth image and frame differ in size")
        p = _arr(prev_xy, np.float32, 2); px = _arr(prev_xyz, np.float32, 3); n = len(p)
        cam = cam or make_camera(); params = params or default_ransac_params()
        flags = (self.KLT_GET_MIN_EIGENVALS if min_eig_err else 0) | (self.KLT_USE_INITIAL_FLOW if init_xy is not None else 0)
        m1 = max(n, 1)
        xy = np.zeros((m1, 2), np.float32)
        if init_xy is not None:
            xy[:n] = _arr(init_xy, np.float32, 2)
        st = np.zeros(m1, np.uint8); err = np.zeros(m1, np.float32); kept = np.zeros(m1, np.int32); m = C.c_int(0)
        und = np.zeros((m1, 2), np.float32); xyz = np.zeros((m1, 3), np.float32); dd = np.zeros(m1, np.float64)
        inl = np.zeros(m1, np.int32); res = FrameResult()
        self._ck(self.lib.pslam_klt_frame(self.h, _p(prev, C.c_uint8), _p(cur, C.c_uint8), W, H, ch * W, ch, _p(p, C.c_float),
                                          _p(px, C.c_float), _p(xy, C.c_float), n, int(win), int(max_level), int(criteria_type),
                                          int(max_iter), C.c_double(eps), flags, C.c_double(min_eig_threshold),
                                          C.c_double(error_threshold), C.c_double(min_distance), _p(depth, C.c_uint16), W,
                                          C.byref(cam), int(bool(undistort)), C.c_double(depth_scale), C.byref(params),
                                          C.c_uint64(seed), int(num_hyp), _p(st, C.c_uint8), _p(err, C.c_float),
                                          _p(kept, C.c_int), C.byref(m), _p(und, C.c_float), _p(xyz, C.c_float),
                                          _p(dd, C.c_double), _p(inl, C.c_int), C.byref(res)))
        k = m.value
        return dict(xy=xy[:n].copy(), status=st[:n].copy(), err=err[:n].copy(), kept=kept[:k].copy(), uv_undist=und[:k].copy(),
                    xyz=xyz[:k].copy(), det_dist=dd[:k].copy(), inliers=inl[:res.n_inliers].copy(),
                    T=np.array(res.T, np.float32).reshape(4, 4).T.copy(), best_ratio=res.best_ratio,
                    inlier_ratio=res.inlier_ratio, hyp_used=res.hyp_used, n_filtered=res.n_filtered, n_matches=res.n_matches)

    # ---- stage 2 ----
    def match_bf_mutual(self, query, train):
        q = _arr(query, np.uint8); t = _arr(train, np.uint8)
        nq = q.shape[0] if q.ndim == 2 else 0
        nt = t.shape[0] if t.ndim == 2 else 0
        nb = q.shape[1] if q.ndim == 2 else (t.shape[1] if t.ndim == 2 else 32)
        cap = max(1, min(nq, nt))
        oq = np.empty(cap, np.int32); ot = np.empty(cap, np.int32); od = np.empty(cap, np.float32)
        n = C.c_int(0)
        self._ck(self.lib.pslam_match_bf_mutual(self.h, _p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, nb, _p(oq, C.c_int),
                                                _p(ot, C.c_int), _p(od, C.c_float), C.byref(n)))
        return oq[:n.value].copy(), ot[:n.value].copy(), od[:n.value].copy()

    def match_knn2(self, query, train):
        q = _arr(query, np.uint8); t = _arr(train, np.uint8)
        nq = q.shape[0]
        nt = t.shape[0] if t.ndim == 2 else 0
        idx = np.empty((nq, 2), np.int32); dist = np.empty((nq, 2), np.float32)
        self._ck(self.lib.pslam_match_knn2(self.h, _p(q, C.c_uint8), nq, _p(t, C.c_uint8), nt, q.shape[1],
                                           _p(idx, C.c_int), _p(dist, C.c_float)))
        return idx, dist

    def match_guided_xyz(self, map_xyz, map_desc, map_level, cur_xyz, cur_desc, cur_level, radius, ratio, mode=0,
                         cap=65536):
        mx = _arr(map_xyz, np.float32, 3); cx = _arr(cur_xyz, np.float32, 3)
        md = _arr(map_desc, np.uint8); cd = _arr(cur_desc, np.uint8)
        ml = _arr(map_level, np.int32); cl = _arr(cur_level, np.int32)
        M, N = ml.size, cl.size
        oq = np.empty(cap, np.int32); ot = np.empty(cap, np.int32); od = np.empty(cap, np.float32)
        n = C.c_int(0); perfect = C.c_int(0)
        r = self.lib.pslam_match_guided_xyz(self.h, _p(mx, C.c_float), _p(md, C.c_uint8), _p(ml, C.c_int), M,
                                            _p(cx, C.c_float), _p(cd, C.c_uint8), _p(cl, C.c_int), N, 32,
                                            C.c_double(radius), C.c_double(ratio), mode, _p(oq, C.c_int),
                                            _p(ot, C.c_int), _p(od, C.c_float), cap, C.byref(n), C.byref(perfect))
        self._ck(r, allow=(ERR_CAPACITY,))
        k = min(n.value, cap)
        return dict(q=oq[:k].copy(), t=ot[:k].copy(), d=od[:k].copy(), total=n.value, perfect=perfect.value,
                    truncated=(r == ERR_CAPACITY))

    # ---- stage 3 ----
    def ransac_estimate(self, prev, cur, mq, mt, params=None, seed=0, num_hyp=0, want_counts=False, counts_cap=None):
        prev = _arr(prev, np.float32, 3); cur = _arr(cur, np.float32, 3)
        mq = _arr(mq, np.int32); mt = _arr(mt, np.int32)
        m = mq.size
        params = params or default_ransac_params()
        T = np.empty(16, np.float32); inl = np.empty(max(1, m), np.int32)
        n_inl = C.c_int(0); best = C.c_double(0); used = C.c_int(0)
        self._ck(self.lib.pslam_ransac_estimate(self.h, _p(prev, C.c_float), prev.shape[0] if prev.size else 0,
                                                _p(cur, C.c_float), cur.shape[0] if cur.size else 0, _p(mq, C.c_int),
                                                _p(mt, C.c_int), m, C.byref(params), C.c_uint64(seed), num_hyp,
                                                _p(T, C.c_float), _p(inl, C.c_int), C.byref(n_inl), C.byref(best),
                                                C.byref(used)))
        counts = None
        if want_counts:
            cap = counts_cap or max(num_hyp, 487)
            counts = np.empty(cap, np.int32); n = C.c_int(0)
            self._ck(self.lib.pslam_ransac_last_counts(self.h, _p(counts, C.c_int), cap, C.byref(n)))
            counts = counts[:n.value]
        return dict(T=T.reshape(4, 4).T.copy(), inliers=inl[:n_inl.value].copy(), best_ratio=best.value,
                    hyp_used=used.value, counts=counts)

    def ransac_set_stopping(self, rule, confidence=0.99):
        self._ck(self.lib.pslam_ransac_set_stopping(self.h, int(rule), C.c_double(confidence)))

    def kabsch_batch(self, A_list, B_list):
        off = np.zeros(len(A_list) + 1, np.int32)
        for i, a in enumerate(A_list):
            off[i + 1] = off[i] + len(a)
        A = np.concatenate([np.asarray(a, np.float64).reshape(-1, 3) for a in A_list]) if off[-1] else np.zeros((0, 3))
        B = np.concatenate([np.asarray(b, np.float64).reshape(-1, 3) for b in B_list]) if off[-1] else np.zeros((0, 3))
        A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
        T = np.empty((len(A_list), 12), np.float64)
        self._ck(self.lib.pslam_kabsch_batch(self.h, _p(A, C.c_double), _p(B, C.c_double), _p(off, C.c_int),
                                             len(A_list), _p(T, C.c_double)))
        return [t.reshape(4, 3).T.copy() for t in T]   # column-major 3x4 -> numpy [3,4]

    # ---- fused pipelines ----
    def frame_to_frame(self, prev_desc, prev_xyz, cur_desc, cur_uv, depth, cam=None, undistort=False,
                       depth_scale=5000.0, params=None, seed=0, num_hyp=0):
        pd = _arr(prev_desc, np.uint8); px = _arr(prev_xyz, np.float32, 3)
        cd = _arr(cur_desc, np.uint8); uv = _arr(cur_uv, np.float32, 2)
        depth = np.ascontiguousarray(depth, np.uint16)
        H, W = depth.shape
        n_prev = pd.shape[0] if pd.ndim == 2 else 0
        n_cur = cd.shape[0]
        cam = cam or make_camera(); params = params or default_ransac_params()
        cap = max(1, min(n_prev, n_cur))
        xyz = np.empty((n_cur, 3), np.float32); und = np.empty((n_cur, 2), np.float32); dd = np.empty(n_cur, np.float64)
        mq = np.empty(cap, np.int32); mt = np.empty(cap, np.int32); md = np.empty(cap, np.float32)
        inl = np.empty(cap, np.int32)
        res = FrameResult()
        self._ck(self.lib.pslam_frame_to_frame(self.h, _p(pd, C.c_uint8), _p(px, C.c_float), n_prev, _p(cd, C.c_uint8),
                                               _p(uv, C.c_float), n_cur, _p(depth, C.c_uint16), W, H, W, C.byref(cam),
                                               int(bool(undistort)), C.c_double(depth_scale), C.byref(params),
                                               C.c_uint64(seed), num_hyp, _p(xyz, C.c_float), _p(und, C.c_float),
                                               _p(dd, C.c_double), _p(mq, C.c_int), _p(mt, C.c_int), _p(md, C.c_float),
                                               _p(inl, C.c_int), C.byref(res)))
        n = res.n_matches
        return dict(xyz=xyz, uv_undist=und, det_dist=dd, mq=mq[:n].copy(), mt=mt[:n].copy(), md=md[:n].copy(),
                    inliers=inl[:res.n_inliers].copy(), T=np.array(res.T, np.float32).reshape(4, 4).T.copy(),
                    best_ratio=res.best_ratio, inlier_ratio=res.inlier_ratio, hyp_used=res.hyp_used,
                    n_filtered=res.n_filtered)

    def frame_to_map(self, map_xyz, map_desc, map_level, cur_xyz, cur_desc, cur_level, radius=0.12, ratio=0.55,
                     mode=0, params=None, seed=0, num_hyp=0, match_cap=16384):
        mx = _arr(map_xyz, np.float32, 3); cx = _arr(cur_xyz, np.float32, 3)
        md_ = _arr(map_desc, np.uint8); cd = _arr(cur_desc, np.uint8)
        ml = _arr(map_level, np.int32); cl = _arr(cur_level, np.int32)
        params = params or default_ransac_params()
        mq = np.empty(match_cap, np.int32); mt = np.empty(match_cap, np.int32); md = np.empty(match_cap, np.float32)
        inl = np.empty(match_cap, np.int32)
        res = FrameResult()
        self._ck(self.lib.pslam_frame_to_map(self.h, _p(mx, C.c_float), _p(md_, C.c_uint8), _p(ml, C.c_int), ml.size,
                                             _p(cx, C.c_float), _p(cd, C.c_uint8), _p(cl, C.c_int), cl.size,
                                             C.c_double(radius), C.c_double(ratio), mode, C.byref(params),
                                             C.c_uint64(seed), num_hyp, match_cap, _p(mq, C.c_int), _p(mt, C.c_int),
                                             _p(md, C.c_float), _p(inl, C.c_int), C.byref(res)))
        n = min(res.n_matches, match_cap)
        return dict(mq=mq[:n].copy(), mt=mt[:n].copy(), md=md[:n].copy(), inliers=inl[:res.n_inliers].copy(),
                    T=np.array(res.T, np.float32).reshape(4, 4).T.copy(), best_ratio=res.best_ratio,
                    inlier_ratio=res.inlier_ratio, hyp_used=res.hyp_used, n_filtered=res.n_filtered)

    def loop_closure_pair(self, desc0, xyz0, desc1, xyz1, params=None, seed=0, num_hyp=0):
        d0 = _arr(desc0, np.uint8); x0 = _arr(xyz0, np.float32, 3); d1 = _arr(desc1, np.uint8); x1 = _arr(xyz1, np.float32, 3)
        n0 = d0.shape[0] if d0.ndim == 2 else 0
        n1 = d1.shape[0] if d1.ndim == 2 else 0
        params = params or default_ransac_params()
        cap = max(1, min(n0, n1))
        mq = np.empty(cap, np.int32); mt = np.empty(cap, np.int32); md = np.empty(cap, np.float32); inl = np.empty(cap, np.int32)
        res = FrameResult()
        self._ck(self.lib.pslam_loop_closure_pair(self.h, _p(d0, C.c_uint8), _p(x0, C.c_float), n0, _p(d1, C.c_uint8),
                                                  _p(x1, C.c_float), n1, C.byref(params), C.c_uint64(seed), num_hyp,
                                                  _p(mq, C.c_int), _p(mt, C.c_int), _p(md, C.c_float), _p(inl, C.c_int),
                                                  C.byref(res)))
        n = res.n_matches
        return dict(mq=mq[:n].copy(), mt=mt[:n].copy(), md=md[:n].copy(), inliers=inl[:res.n_inliers].copy(),
                    T=np.array(res.T, np.float32).reshape(4, 4).T.copy(), best_ratio=res.best_ratio,
                    inlier_ratio=res.inlier_ratio, hyp_used=res.hyp_used)

    def map_prepare(self, map_xyz, view_axis, pose, params):
        """pose: numpy 4x4 (camera -> global)."""
        x = _arr(map_xyz, np.float64, 3); a = _arr(view_axis, np.float32, 3)
        M = x.shape[0] if x.size else 0
        pcm = np.ascontiguousarray(np.asarray(pose, np.float64).T)
        kept = np.empty(max(1, M), np.int32); xl = np.empty((max(1, M), 3), np.float64)
        uv = np.empty((max(1, M), 2), np.float64); ang = np.empty(max(1, M), np.float64)
        n = C.c_int(0)
        self._ck(self.lib.pslam_map_prepare(self.h, _p(x, C.c_double), _p(a, C.c_float), M, _p(pcm, C.c_double), C.byref(params),
                                            _p(kept, C.c_int), _p(xl, C.c_double), _p(uv, C.c_double), _p(ang, C.c_double),
                                            C.byref(n)))
        k = n.value
        return kept[:k].copy(), xl[:k].copy(), uv[:k].copy(), ang[:k].copy()

    def frame_to_map_features(self, map_xyz, map_desc, map_octave, map_detdist, cur_xyz, cur_desc, cur_octave, cur_detdist,
                              radius=0.12, ratio=0.55, mode=0, params=None, seed=0, num_hyp=0, match_cap=16384):
        mx = _arr(map_xyz, np.float64, 3); cx = _arr(cur_xyz, np.float32, 3)
        md_ = _arr(map_desc, np.uint8); cd = _arr(cur_desc, np.uint8)
        mo = _arr(map_octave, np.int32); co = _arr(cur_octave, np.int32)
        mdd = _arr(map_detdist, np.float64); cdd = _arr(cur_detdist, np.float64)
        params = params or default_ransac_params()
        mq = np.empty(match_cap, np.int32); mt = np.empty(match_cap, np.int32); md = np.empty(match_cap, np.float32)
        inl = np.empty(match_cap, np.int32)
        res = FrameResult()
        self._ck(self.lib.pslam_frame_to_map_features(self.h, _p(mx, C.c_double), _p(md_, C.c_uint8), _p(mo, C.c_int),
                                                      _p(mdd, C.c_double), mo.size, _p(cx, C.c_float), _p(cd, C.c_uint8),
                                                      _p(co, C.c_int), _p(cdd, C.c_double), co.size, C.c_double(radius),
                                                      C.c_double(ratio), mode, C.byref(params), C.c_uint64(seed), num_hyp,
                                                      match_cap, _p(mq, C.c_int), _p(mt, C.c_int), _p(md, C.c_float),
                                                      _p(inl, C.c_int), C.byref(res)))
        n = min(res.n_matches, match_cap)
        return dict(mq=mq[:n].copy(), mt=mt[:n].copy(), md=md[:n].copy(), inliers=inl[:res.n_inliers].copy(),
                    T=np.array(res.T, np.float32).reshape(4, 4).T.copy(), best_ratio=res.best_ratio,
                    inlier_ratio=res.inlier_ratio, hyp_used=res.hyp_used, n_filtered=res.n_filtered)

    # ---- resident feature map ----
    def map_reserve(self, n):
        self._ck(self.lib.pslam_map_reserve(self.h, int(n)))

    def map_write(self, first, xyz=None, desc=None, octave=None, detdist=None, view_axis=None, count=None):
        x = None if xyz is None else _arr(xyz, np.float64, 3)
        d = None if desc is None else _arr(desc, np.uint8)
        o = None if octave is None else _arr(octave, np.int32)
        t = None if detdist is None else _arr(detdist, np.float64)
        a = None if view_axis is None else _arr(view_axis, np.float32, 3)
        if count is None:
            count = next(v.shape[0] if v.ndim > 1 else v.size for v in (x, o, t, a, d) if v is not None)
        self._ck(self.lib.pslam_map_write(self.h, int(first), int(count), None if x is None else _p(x, C.c_double),
                                          None if d is None else _p(d, C.c_uint8), None if o is None else _p(o, C.c_int),
                                          None if t is None else _p(t, C.c_double), None if a is None else _p(a, C.c_float)))

    def map_truncate(self, n):
        self._ck(self.lib.pslam_map_truncate(self.h, int(n)))

    def map_size(self):
        n = C.c_int(0)
        self._ck(self.lib.pslam_map_size(self.h, C.byref(n)))
        return n.value

    def frame_to_resident_map(self, pose, prep, cur_xyz, cur_desc, cur_octave, cur_detdist, radius=0.12, ratio=0.55, mode=0,
                              params=None, seed=0, num_hyp=0, match_cap=16384, want_local=False):
        """pose: numpy 4x4 (camera -> global); prep: MapPrepareParams.  mq indexes `kept`."""
        cx = _arr(cur_xyz, np.float32, 3); cd = _arr(cur_desc, np.uint8)
        co = _arr(cur_octave, np.int32); cdd = _arr(cur_detdist, np.float64)
        pcm = np.ascontiguousarray(np.asarray(pose, np.float64).T)
        params = params or default_ransac_params()
        M = max(1, self.map_size())
        kept = np.empty(M, np.int32); nk = C.c_int(0)
        xl = np.empty((M, 3), np.float64) if want_local else None
        uv = np.empty((M, 2), np.float64) if want_local else None
        mq = np.empty(match_cap, np.int32); mt = np.empty(match_cap, np.int32); md = np.empty(match_cap, np.float32)
        inl = np.empty(match_cap, np.int32)
        res = FrameResult()
        self._ck(self.lib.pslam_frame_to_resident_map(
            self.h, _p(pcm, C.c_double), C.byref(prep), _p(cx, C.c_float), _p(cd, C.c_uint8), _p(co, C.c_int),
            _p(cdd, C.c_double), co.size, C.c_double(radius), C.c_double(ratio), mode, C.byref(params), C.c_uint64(seed),
            num_hyp, match_cap, _p(kept, C.c_int), C.byref(nk), _p(xl, C.c_double) if want_local else None,
            _p(uv, C.c_double) if want_local else None, _p(mq, C.c_int), _p(mt, C.c_int), _p(md, C.c_float),
            _p(inl, C.c_int), C.byref(res)))
        n = min(res.n_matches, match_cap); k = nk.value
        out = dict(kept=kept[:k].copy(), mq=mq[:n].copy(), mt=mt[:n].copy(), md=md[:n].copy(),
                   inliers=inl[:res.n_inliers].copy(), T=np.array(res.T, np.float32).reshape(4, 4).T.copy(),
                   best_ratio=res.best_ratio, inlier_ratio=res.inlier_ratio, hyp_used=res.hyp_used,
                   n_filtered=res.n_filtered)
        if want_local:
            out["xyz_local"] = xl[:k].copy(); out["uv"] = uv[:k].copy()
        return out

    def frame_to_map_resident(self):
        self._ck(self.lib.pslam_frame_to_map_resident(self.h))

    def frame_to_frame_resident(self):
        self._ck(self.lib.pslam_frame_to_frame_resident(self.h))

    # ---- loop-closure database ----
    def lc_reserve(self, max_desc, max_kf):
        self._ck(self.lib.pslam_lc_db_reserve(self.h, C.c_int64(max_desc), int(max_kf)))

    def lc_append(self, desc, kf_off):
        d = _arr(desc, np.uint8); off = _arr(kf_off, np.int64)
        self._ck(self.lib.pslam_lc_db_append(self.h, _p(d, C.c_uint8), _p(off, C.c_int64), off.size - 1))

    def lc_clear(self):
        self._ck(self.lib.pslam_lc_db_clear(self.h))

    def lc_size(self):
        nk = C.c_int(0); nd = C.c_int64(0)
        self._ck(self.lib.pslam_lc_db_size(self.h, C.byref(nk), C.byref(nd)))
        return nk.value, nd.value

    def lc_set_work_unit(self, mode):
        self._ck(self.lib.pslam_lc_set_work_unit(self.h, int(mode)))

    def lc_tensor_status(self):
        """(used_tensor_cores, timed_out) of the last V1 sweep"""
        used, to = C.c_int(0), C.c_int(0)
        self._ck(self.lib.pslam_lc_tensor_status(self.h, C.byref(used), C.byref(to)))
        return bool(used.value), int(to.value)

    def lc_set_id_base(self, base):
        self._ck(self.lib.pslam_lc_set_id_base(self.h, int(base)))

    def lc_query(self, query, tau=64, k=16, want_scores=False):
        q = _arr(query, np.uint8)
        ids = np.empty(k, np.int32); sc = np.empty(k, np.int32)
        scores = np.empty(max(1, self.lc_size()[0]), np.int32) if want_scores else None
        self._ck(self.lib.pslam_lc_query(self.h, _p(q, C.c_uint8), q.shape[0], tau, k, _p(ids, C.c_int), _p(sc, C.c_int),
                                         _p(scores, C.c_int)))
        if want_scores:
            return ids, sc, scores[:self.lc_size()[0]]
        return ids, sc

    def lc_query_resident(self, tau=64, k=16):
        self._ck(self.lib.pslam_lc_query_resident(self.h, tau, k))

    def lc_last_sweep_ms(self):
        ms = C.c_float(0)
        self._ck(self.lib.pslam_lc_last_sweep_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def comm_init(self, uid_bytes, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(uid_bytes))
        self._ck(self.lib.pslam_comm_init(self.h, buf, rank, world))

    def host_stamps(self):
        out = (C.c_double * 8)()
        self._ck(self.lib.pslam_debug_host_stamps(self.h, out))
        return list(out)

    def host_register(self, arr):
        """page-lock a numpy array that will be passed to frame_to_map* repeatedly (skips the staging copy)"""
        self._ck(self.lib.pslam_host_register(self.h, C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes)))

    def host_unregister(self, arr):
        self._ck(self.lib.pslam_host_unregister(self.h, C.c_void_p(arr.ctypes.data)))

    def lc_exchange_mode(self):
        """0 single rank, 1 NCCL collectives, 2 peer memory over NVLink inside the sweep kernel"""
        return int(self.lib.pslam_lc_exchange_mode(self.h))

    def comm_destroy(self):
        self._ck(self.lib.pslam_comm_destroy(self.h))

    def lc_query_sharded(self, query, root=-1, tau=64, k=16, nq=None):
        q = _arr(query, np.uint8) if query is not None else None
        nq = q.shape[0] if q is not None else int(nq)
        ids = np.empty(k, np.int32); sc = np.empty(k, np.int32)
        self._ck(self.lib.pslam_lc_query_sharded(self.h, _p(q, C.c_uint8), nq, root, tau, k, _p(ids, C.c_int),
                                                 _p(sc, C.c_int)))
        return ids, sc

    def lc_set_desc_base(self, base):
        self._ck(self.lib.pslam_lc_set_desc_base(self.h, C.c_int64(int(base))))

    def lc_knn2(self, query, sharded=False, root=-1, nq=None):
        q = _arr(query, np.uint8) if query is not None else None
        nq = q.shape[0] if q is not None else int(nq)
        idx = np.empty((nq, 2), np.int64); dist = np.empty((nq, 2), np.float32)
        if sharded:
            self._ck(self.lib.pslam_lc_knn2_sharded(self.h, _p(q, C.c_uint8), nq, root, _p(idx, C.c_int64), _p(dist, C.c_float)))
        else:
            self._ck(self.lib.pslam_lc_knn2(self.h, _p(q, C.c_uint8), nq, _p(idx, C.c_int64), _p(dist, C.c_float)))
        return idx, dist

    def lc_knn2_resident(self, sharded=False):
        self._ck(self.lib.pslam_lc_knn2_resident(self.h, int(bool(sharded))))

    def lc_query_sharded_resident(self, tau=64, k=16, root=None):
        """replay on the resident query; root >= 0: broadcast it from that rank first (the whole sharded exchange on the device)"""
        if root is None:
            self._ck(self.lib.pslam_lc_query_sharded_resident(self.h, tau, k))
        else:
            self._ck(self.lib.pslam_lc_query_sharded_resident_bcast(self.h, int(root), tau, k))


def comm_unique_id():
    lib = load_library()
    buf = (C.c_uint8 * 128)()
    r = lib.pslam_comm_unique_id(buf)
    if r != PSLAM_OK:
        raise PslamError(r, "pslam_comm_unique_id failed (libnccl.so.2 not loadable)")
    return bytes(buf)


def ransac_sample(seed, hyp, m):
    lib = load_library()
    out = (C.c_int * 3)()
    lib.pslam_ransac_sample(C.c_uint64(seed), C.c_uint32(hyp), int(m), out)
    return np.array(list(out), np.int32)


def point_inlier_ratio(inl_t, all_t):
    lib = load_library()
    a = _arr(inl_t, np.int32); b = _arr(all_t, np.int32)
    return lib.pslam_point_inlier_ratio(_p(a, C.c_int), a.size, _p(b, C.c_int), b.size)
