/*
 * oracle.c -- CPU restatement of PUTSLAM's front-end hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (putslam_b200/, adapter/) never includes, links or calls anything in oracle/.
 *
 * Every function cites the reference file:line (relative to /root/reference) it follows.
 * Compile with:  gcc -O2 -ffp-contract=off -fno-fast-math  (the reference is built
 * -std=c++11 for baseline x86-64: scalar IEEE-754, no FMA; CMakeLists.txt:23,147).
 *
 * PARITY STATUS
 *   pinned to the real OpenCV (cv2 4.13, golden vectors in tests/golden/): brute-force cross-check
 *              matching, knn-2, the saturating-subtract "Hamming" of matchXYZ, undistortPoints.
 *   pinned to the REFERENCE'S OWN COMPILED CODE (oracle/_ref/libref_frontend.so = RANSAC.cpp, RGBD.cpp,
 *              kabschEst.cpp, depthSensorModel.cpp, matcher.cpp, dbscan.cpp compiled from /root/reference
 *              against the Eigen / OpenCV stand-ins of oracle/ref_shim; tests/test_ref_build_cpu.py):
 *              the RANSAC driver (filter, adaptive bound, strict >, refit + Euclidean recount, rejection),
 *              all four live error versions, the guided gate and level prediction of matchXYZ, Matcher::match,
 *              matchFeatureLoopClosure, back-projection / projection, the sensor-model covariances, normals and
 *              RGB gradients, Kabsch, the transform covariance, the list edits of trackKLT -- inlier index sets,
 *              hypotheses drawn and poses equal bit for bit over 240 seeded RANSAC cases.
 *   NOT pinned (Eigen is neither vendored by the reference nor present in this image): the arithmetic ORDER
 *              inside Eigen's own operations (fixed-size sums, umeyama's scaling, JacobiSVD's threshold and sweep
 *              order).  The stand-in and this file model Eigen 3.3 (the reference needs >= 3.3: Eigen::Index);
 *              profiles/umeyama_sensitivity.json bounds the rest: over 630 C1-C3 frames no Eigen-version
 *              variant changes a final inlier set or the number of hypotheses, poses move < 4e-6 m / 1e-6 rad.
 */
#define _DEFAULT_SOURCE   /* M_PI under -std=c11 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Stage 2 -- Hamming matching
 * ---------------------------------------------------------------------------------------- */

/* cv::norm(a, b, NORM_HAMMING) over `bytes` bytes == popcount(a XOR b)
 * (third-party OpenCV; call site src/Matcher/matcherOpenCV.cpp:105,203). */
static inline int hamming_xor(const uint8_t *a, const uint8_t *b, int bytes) {
    int d = 0, i = 0;
    for (; i + 8 <= bytes; i += 8) {
        uint64_t x, y;
        memcpy(&x, a + i, 8);
        memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; i < bytes; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

/* The quirk distance of Matcher::matchXYZ (src/Matcher/matcher.cpp:719-721,737-739):
 *   cv::Mat x = descA - descB;   // CV_8U => per-byte SATURATING subtraction
 *   norm(x, NORM_HAMMING)        // popcount of the saturated difference bytes */
static inline int hamming_satsub(const uint8_t *a, const uint8_t *b, int bytes) {
    int d = 0;
    for (int i = 0; i < bytes; ++i) {
        int v = (int)a[i] - (int)b[i];
        if (v < 0) v = 0;
        d += __builtin_popcount((unsigned)v);
    }
    return d;
}

ORC_API int orc_hamming_xor(const uint8_t *a, const uint8_t *b, int bytes) { return hamming_xor(a, b, bytes); }
ORC_API int orc_hamming_satsub(const uint8_t *a, const uint8_t *b, int bytes) { return hamming_satsub(a, b, bytes); }

/* cv::BFMatcher(NORM_HAMMING, crossCheck=true).match(query, train)
 * (src/Matcher/matcherOpenCV.cpp:97-106 builds it, :198-206 calls it; semantics pinned
 * against cv2 in SURVEY Appendix A.2): mutual nearest neighbour, first-argmin tie-break in
 * both directions, output ascending in queryIdx, distance = integer Hamming as float.
 * Returns the number of matches written. */
ORC_API int orc_bf_mutual(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                          int *out_q, int *out_t, float *out_d) {
    if (nq <= 0 || nt <= 0) return 0;
    int *fwd = (int *)malloc(sizeof(int) * (size_t)nq);
    int *fwd_d = (int *)malloc(sizeof(int) * (size_t)nq);
    int *bwd = (int *)malloc(sizeof(int) * (size_t)nt);
    int *bwd_d = (int *)malloc(sizeof(int) * (size_t)nt);
    for (int j = 0; j < nt; ++j) { bwd[j] = -1; bwd_d[j] = INT_MAX; }
    for (int i = 0; i < nq; ++i) {
        int best = INT_MAX, bi = -1;
        for (int j = 0; j < nt; ++j) {
            int d = hamming_xor(q + (size_t)i * bytes, t + (size_t)j * bytes, bytes);
            if (d < best) { best = d; bi = j; }              /* first argmin over train */
            if (d < bwd_d[j]) { bwd_d[j] = d; bwd[j] = i; }  /* first argmin over query */
        }
        fwd[i] = bi; fwd_d[i] = best;
    }
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        int j = fwd[i];
        if (j >= 0 && bwd[j] == i) {
            out_q[n] = i; out_t[n] = j; out_d[n] = (float)fwd_d[i]; ++n;
        }
    }
    free(fwd); free(fwd_d); free(bwd); free(bwd_d);
    return n;
}

/* Multi-threaded variant of the same computation (rows split over OpenMP threads, column
 * minima merged in thread order so ties still resolve to the lowest query index).  Used only
 * as the "all host cores" CPU baseline in bench.py. */
ORC_API int orc_bf_mutual_mt(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                             int *out_q, int *out_t, float *out_d, int threads) {
#ifdef _OPENMP
    if (nq <= 0 || nt <= 0) return 0;
    if (threads < 1) threads = 1;
    int *fwd = (int *)malloc(sizeof(int) * (size_t)nq);
    int *fwd_d = (int *)malloc(sizeof(int) * (size_t)nq);
    int *bwd_all = (int *)malloc(sizeof(int) * (size_t)nt * threads);
    int *bwd_d_all = (int *)malloc(sizeof(int) * (size_t)nt * threads);
#pragma omp parallel num_threads(threads)
    {
        int tid = omp_get_thread_num(), nth = omp_get_num_threads();
        int lo = (int)((long long)nq * tid / nth), hi = (int)((long long)nq * (tid + 1) / nth);
        int *bwd = bwd_all + (size_t)tid * nt, *bwd_d = bwd_d_all + (size_t)tid * nt;
        for (int j = 0; j < nt; ++j) { bwd[j] = -1; bwd_d[j] = INT_MAX; }
        for (int i = lo; i < hi; ++i) {
            int best = INT_MAX, bi = -1;
            for (int j = 0; j < nt; ++j) {
                int d = hamming_xor(q + (size_t)i * bytes, t + (size_t)j * bytes, bytes);
                if (d < best) { best = d; bi = j; }
                if (d < bwd_d[j]) { bwd_d[j] = d; bwd[j] = i; }
            }
            fwd[i] = bi; fwd_d[i] = best;
        }
#pragma omp barrier
#pragma omp single
        {
            for (int k = 1; k < nth; ++k)
                for (int j = 0; j < nt; ++j)
                    if (bwd_d_all[(size_t)k * nt + j] < bwd_d_all[j]) {
                        bwd_d_all[j] = bwd_d_all[(size_t)k * nt + j];
                        bwd_all[j] = bwd_all[(size_t)k * nt + j];
                    }
        }
    }
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        int j = fwd[i];
        if (j >= 0 && bwd_all[j] == i) { out_q[n] = i; out_t[n] = j; out_d[n] = (float)fwd_d[i]; ++n; }
    }
    free(fwd); free(fwd_d); free(bwd_all); free(bwd_d_all);
    return n;
#else
    (void)threads;
    return orc_bf_mutual(q, nq, t, nt, bytes, out_q, out_t, out_d);
#endif
}

/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, k=2): the two smallest distances per
 * query, ascending, lowest train index first on ties (stable).  Extension named by
 * north_star (Lowe ratio test); oracle = cv2.  idx/dist are nq x 2; slots
 * beyond nt are -1. */
ORC_API void orc_knn2(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes,
                      int *idx, int *dist) {
    for (int i = 0; i < nq; ++i) {
        int d1 = INT_MAX, d2 = INT_MAX, i1 = -1, i2 = -1;
        for (int j = 0; j < nt; ++j) {
            int d = hamming_xor(q + (size_t)i * bytes, t + (size_t)j * bytes, bytes);
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
            else if (d < d2) { d2 = d; i2 = j; }
        }
        idx[2 * i] = i1; idx[2 * i + 1] = i2;
        dist[2 * i] = i1 < 0 ? -1 : d1; dist[2 * i + 1] = i2 < 0 ? -1 : d2;
    }
}

/* Predicted ORB pyramid level (src/Matcher/matcher.cpp:641-651 for current keypoints,
 * :682-692 for map features; constants include/putslam/Matcher/matcher.h:26-28):
 *   level = clamp((int)ceil(log(pow(1.2, octave) * detDist / curDist) / log(1.2)), 0, 7)
 * all in double.  curDist is supplied by the caller (float norm promoted for current
 * keypoints, double norm of the double position for map features). */
ORC_API int orc_pred_level(int octave, double det_dist, double cur_dist) {
    const double scaleFactor = 1.2;
    const double logScaleFactor = log(scaleFactor);
    double detLevelScaleFactor = pow(scaleFactor, octave);
    double curLevelScaleFactor = detLevelScaleFactor * det_dist / cur_dist;
    int curLevel = (int)ceil(log(curLevelScaleFactor) / logScaleFactor);
    if (curLevel < 0) curLevel = 0;
    if (curLevel > 7) curLevel = 7;
    return curLevel;
}

/* Eigen Vector3f::norm(): squaredNorm via the unrolled binary-split redux x^2 + (y^2 + z^2),
 * then sqrtf. */
static inline float norm3f(float x, float y, float z) {
    float yy = y * y, zz = z * z;
    float s = x * x + (yy + zz);
    return sqrtf(s);
}
ORC_API float orc_norm3f(float x, float y, float z) { return norm3f(x, y, z); }

/* Guided map matching -- the inner loops of Matcher::matchXYZ (src/Matcher/matcher.cpp:670-748).
 *   map_xyz   : M x 3 float (positions already cast double->float, :665 / :701-702)
 *   map_level : M ints (predicted level of each map feature, :682-692, host-computed)
 *   cur_level : N ints (:639-651)
 *   radius, ratio : already adjusted for the retry number by the caller (:619-622)
 *   mode 0 = saturating-subtract quirk (reference behaviour), 1 = XOR Hamming
 * Emits DMatch(queryIdx=j, trainIdx=i, distance=value) in (j, i) order; returns the count
 * (may exceed cap; only the first cap are written). */
ORC_API int orc_guided_match(const float *map_xyz, const uint8_t *map_desc, const int *map_level, int M,
                             const float *cur_xyz, const uint8_t *cur_desc, const int *cur_level, int N,
                             int bytes, double radius, double ratio, int mode,
                             int *out_q, int *out_t, float *out_d, int cap, int *perfect_out) {
    int n = 0, perfect = 0;
    int *cand = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    for (int j = 0; j < M; ++j) {
        int nc = 0;
        const float px = map_xyz[3 * j], py = map_xyz[3 * j + 1], pz = map_xyz[3 * j + 2];
        const int curLevel = map_level[j];
        for (int i = 0; i < N; ++i) {                                   /* :699-711 */
            float dx = px - cur_xyz[3 * i], dy = py - cur_xyz[3 * i + 1], dz = pz - cur_xyz[3 * i + 2];
            float nrm = norm3f(dx, dy, dz);
            int scaleCheck = (cur_level[i] - 1 <= curLevel) && (curLevel <= cur_level[i] + 1);
            int posCheck = (double)nrm < radius;
            if (posCheck && scaleCheck) cand[nc++] = i;
        }
        int bestId = -1;                                                /* :714-726 */
        float bestVal = 99999.f;
        for (int c = 0; c < nc; ++c) {
            int id = cand[c];
            const uint8_t *a = map_desc + (size_t)j * bytes, *b = cur_desc + (size_t)id * bytes;
            float value = (float)(mode == 0 ? hamming_satsub(a, b, bytes) : hamming_xor(a, b, bytes));
            if (value < bestVal || bestId == -1) { bestVal = value; bestId = id; }
        }
        if (bestVal < 0.1) perfect++;                                   /* :729-731 */
        for (int c = 0; c < nc; ++c) {                                  /* :734-747 */
            int id = cand[c];
            const uint8_t *a = map_desc + (size_t)j * bytes, *b = cur_desc + (size_t)id * bytes;
            float value = (float)(mode == 0 ? hamming_satsub(a, b, bytes) : hamming_xor(a, b, bytes));
            if (ratio * (double)value <= (double)bestVal) {
                if (n < cap) { out_q[n] = j; out_t[n] = id; out_d[n] = value; }
                ++n;
            }
        }
    }
    free(cand);
    if (perfect_out) *perfect_out = perfect;
    return n;
}

/* Loop-closure sweep, variant V1: score of keyframe k = number of mutual
 * nearest-neighbour matches between the query frame and keyframe k with distance <= tau.
 * Per keyframe this is exactly Matcher::matchFeatureLoopClosure's performMatching call
 * (src/Matcher/matcher.cpp:835 -> matcherOpenCV.cpp:198-206).  kf_off has n_kf+1 entries
 * (descriptor offsets into db). */
ORC_API void orc_lc_scores(const uint8_t *query, int nq, const uint8_t *db, const int64_t *kf_off,
                           int n_kf, int bytes, int tau, int *scores, int threads) {
    int maxn = 0;
    for (int k = 0; k < n_kf; ++k) { int c = (int)(kf_off[k + 1] - kf_off[k]); if (c > maxn) maxn = c; }
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        int cap = nq < maxn ? nq : maxn;
        if (cap < 1) cap = 1;
        int *oq = (int *)malloc(sizeof(int) * (size_t)cap);
        int *ot = (int *)malloc(sizeof(int) * (size_t)cap);
        float *od = (float *)malloc(sizeof(float) * (size_t)cap);
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < n_kf; ++k) {
            int nt = (int)(kf_off[k + 1] - kf_off[k]);
            int n = orc_bf_mutual(query, nq, db + (size_t)kf_off[k] * bytes, nt, bytes, oq, ot, od);
            int s = 0;
            for (int a = 0; a < n; ++a) if (od[a] <= (float)tau) ++s;
            scores[k] = s;
        }
        free(oq); free(ot); free(od);
    }
}

/* Top-k keyframes by score, ties -> lower keyframe id. */
ORC_API void orc_topk(const int *scores, int n, int k, int *out_id, int *out_score) {
    for (int a = 0; a < k; ++a) { out_id[a] = -1; out_score[a] = -1; }
    for (int i = 0; i < n; ++i) {
        int s = scores[i];
        int pos = k;
        while (pos > 0 && (out_id[pos - 1] < 0 || s > out_score[pos - 1])) --pos;
        if (pos < k) {
            for (int a = k - 1; a > pos; --a) { out_id[a] = out_id[a - 1]; out_score[a] = out_score[a - 1]; }
            out_id[pos] = i; out_score[pos] = s;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Stage 1 -- back-projection with per-point uncertainty
 * ---------------------------------------------------------------------------------------- */

/* cv::undistortPoints(pts, K, dist) followed by the re-projection of
 * RGBD::removeImageDistortion (src/RGBD/RGBD.cpp:268-281 / :298-311).  OpenCV model pinned
 * against cv2: doubles, exactly 5 fixed-point iterations, result cast to float;
 * K and dist are CV_32F values widened to double.  Then u = x*fx + cx in float32. */
ORC_API void orc_undistort(const float *uv, int n, float fx, float fy, float cx, float cy,
                           const float *dist5, float *uv_out) {
    double k1 = dist5[0], k2 = dist5[1], p1 = dist5[2], p2 = dist5[3], k3 = dist5[4];
    double ifx = 1. / (double)fx, ify = 1. / (double)fy;
    for (int i = 0; i < n; ++i) {
        double x = ((double)uv[2 * i] - (double)cx) * ifx;
        double y = ((double)uv[2 * i + 1] - (double)cy) * ify;
        double x0 = x, y0 = y;
        for (int it = 0; it < 5; ++it) {
            double r2 = x * x + y * y;
            double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
            double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        float xf = (float)x, yf = (float)y;
        uv_out[2 * i] = xf * fx + cx;          /* RGBD.cpp:274-279, float32 */
        uv_out[2 * i + 1] = yf * fy + cy;
    }
}

/* RGBD::roundSize (src/RGBD/RGBD.cpp:10-16).  The reference returns `size` for x > size-1
 * (an out-of-bounds pixel); the restatement AND the product clamp to size-1 instead and the
 * parity fixtures keep keypoints off the border (deliberate divergence). */
static inline int round_size(double x, int size) {
    if (x < 0) x = 0;
    else if (x > size - 1) x = size - 1;
    return (int)round(x);
}

/* RGBD::point2Dto3D / keypoints2Dto3D (src/RGBD/RGBD.cpp:30-65) and the detDist loop
 * (src/Matcher/matcher.cpp:51-58). */
ORC_API void orc_backproject(const float *uv, int n, const uint16_t *depth, int W, int H, int stride,
                             float fx, float fy, float cx, float cy, double depth_scale,
                             float *xyz, double *det_dist) {
    for (int i = 0; i < n; ++i) {
        float u2 = uv[2 * i], v2 = uv[2 * i + 1];
        int uR = round_size(u2, W), vR = round_size(v2, H);
        float Z = (float)(((double)depth[(size_t)vR * stride + uR]) / depth_scale);
        float u = (u2 - cx) / fx;
        float v = (v2 - cy) / fy;
        float X = u * Z, Y = v * Z;
        xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
        if (det_dist) {
            /* matcher.cpp:53-55: the operands are float (Vector3f coefficients), so the products,
             * the two adds AND std::sqrt (its float overload) are float32; the float result is
             * then widened to the `double dist`. */
            float s = X * X + Y * Y + Z * Z;
            det_dist[i] = (double)sqrtf(s);
        }
    }
}

/* DepthSensorModel::computeCov(u, v, depth, cov) (src/Grabber/depthSensorModel.cpp:28-36);
 * u, v arrive truncated to unsigned integers (:44,:51,:57).  c[4] = distVarCoefs, varU/varV =
 * Ruvd diagonal (:8-10).  cov is row-major 3x3 double. */
ORC_API void orc_compute_cov(unsigned u, unsigned v, double depth, double fx, double fy, double cx,
                             double cy, double varU, double varV, const double *c, double *cov) {
    double J[3][3] = {{depth / fx, 0.0, (((double)u / fx) - (cx / fx))},
                      {0.0, depth / fy, (((double)v / fy) - (cy / fy))},
                      {0.0, 0.0, 1.0}};
    double R[3] = {varU, varV, c[0] * pow(depth, 3.0) + c[1] * pow(depth, 2.0) + c[2] * depth + c[3]};
    /* cov = (J * Ruvd) * J^T, Ruvd diagonal: Eigen evaluates J*Ruvd first (3x3 products are
     * coefficient-based, sequential inner sum), zeros of the diagonal matrix included. */
    double JR[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = J[i][0] * (j == 0 ? R[0] : 0.0);
            s = s + J[i][1] * (j == 1 ? R[1] : 0.0);
            s = s + J[i][2] * (j == 2 ? R[2] : 0.0);
            JR[i][j] = s;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = JR[i][0] * J[j][0];
            s = s + JR[i][1] * J[j][1];
            s = s + JR[i][2] * J[j][2];
            cov[3 * i + j] = s;
        }
}

/* Eigen Matrix3d::inverse() (compute_inverse_size3: cofactors of column 0 -> determinant -> adjugate * 1/det),
 * as used by DepthSensorModel::informationMatrixFromImageCoordinates (src/Grabber/depthSensorModel.cpp:55-59)
 * and informationMatrix (:48-53).  Third-party arithmetic (Eigen); row-major in and out. */
static void inverse3d(const double *m, double *r) {
#define CF(i, j) (m[3 * (((i) + 1) % 3) + (((j) + 1) % 3)] * m[3 * (((i) + 2) % 3) + (((j) + 2) % 3)] - \
                  m[3 * (((i) + 1) % 3) + (((j) + 2) % 3)] * m[3 * (((i) + 2) % 3) + (((j) + 1) % 3)])
    const double c00 = CF(0, 0), c10 = CF(1, 0), c20 = CF(2, 0);
    const double det = c00 * m[0] + (c10 * m[3] + c20 * m[6]);
    const double invdet = 1.0 / det;
    r[0] = c00 * invdet; r[1] = c10 * invdet; r[2] = c20 * invdet;
    r[3] = CF(0, 1) * invdet; r[4] = CF(1, 1) * invdet; r[5] = CF(2, 1) * invdet;
    r[6] = CF(0, 2) * invdet; r[7] = CF(1, 2) * invdet; r[8] = CF(2, 2) * invdet;
#undef CF
}

/* DepthSensorModel::informationMatrixFromImageCoordinates(u, v, z) (depthSensorModel.cpp:55-59): u, v are
 * truncated to unsigned integers, cov = computeCov(u, v, z), info = cov.inverse().  This is what
 * FeaturesMap::addFeatures / addMeasurements attach to every map measurement when useUncertainty is on
 * (src/Map/featuresMap.cpp:110-114, 263-267). */
ORC_API void orc_information_matrix(double u, double v, double z, double fx, double fy, double cx, double cy,
                                    double varU, double varV, const double *c, double *cov, double *info) {
    double tmp[9];
    orc_compute_cov((unsigned)(unsigned long)u, (unsigned)(unsigned long)v, z, fx, fy, cx, cy, varU, varV, c, cov ? cov : tmp);
    inverse3d(cov ? cov : tmp, info);
}

/* RGBD::point2Dto3D for one pixel (src/RGBD/RGBD.cpp:47-65), shared by the normal estimator below. */
static void point2Dto3D(float fu, float fv, const uint16_t *depth, int W, int H, int stride, float fx, float fy,
                        float cx, float cy, double depth_scale, float *out) {
    int uR = round_size(fu, W), vR = round_size(fv, H);
    float Z = (float)(((double)depth[(size_t)vR * stride + uR]) / depth_scale);
    float u = (fu - cx) / fx, v = (fv - cy) / fy;
    out[0] = u * Z; out[1] = v * Z; out[2] = Z;
}

/* RGBD::computeNormal (src/RGBD/RGBD.cpp:101-144): vectors from the centre pixel's 3-D point to its 8
 * neighbours (those with depth > 0, fixed order), cross products of consecutive vectors (cyclic), average,
 * normalise.  Float differences widened to double, everything after in double.  normal = NaN when fewer than
 * two neighbours have depth (0/0), exactly like the reference. */
ORC_API void orc_compute_normal(const uint16_t *depth, int W, int H, int stride, int u, int v, float fx, float fy,
                                float cx, float cy, double depth_scale, double *normal) {
    static const int uidx[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    static const int vidx[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
    float c[3];
    point2Dto3D((float)u, (float)v, depth, W, H, stride, fx, fy, cx, cy, depth_scale, c);
    double vecs[8][3];
    int nv = 0;
    for (int i = 0; i < 8; ++i) {
        float e[3];
        point2Dto3D((float)(u + uidx[i]), (float)(v + vidx[i]), depth, W, H, stride, fx, fy, cx, cy, depth_scale, e);
        if (e[2] > 0) {
            vecs[nv][0] = (double)(e[0] - c[0]); vecs[nv][1] = (double)(e[1] - c[1]); vecs[nv][2] = (double)(e[2] - c[2]);
            ++nv;
        }
    }
    double sx = 0, sy = 0, sz = 0;
    for (int i = 0; i < nv; ++i) {
        const double *a = vecs[i], *b = vecs[(i == nv - 1) ? 0 : i + 1];
        sx += a[1] * b[2] - a[2] * b[1];
        sy += a[2] * b[0] - a[0] * b[2];
        sz += a[0] * b[1] - a[1] * b[0];
    }
    double nx = sx / (double)nv, ny = sy / (double)nv, nz = sz / (double)nv;
    double yy = ny * ny, zz = nz * nz;
    double norm = sqrt(nx * nx + (yy + zz));
    normal[0] = nx / norm; normal[1] = ny / norm; normal[2] = nz / norm;
}

/* DepthSensorModel::uncertinatyFromNormal (src/Grabber/depthSensorModel.cpp:62-76): orthonormal-ish frame
 * {x, y, n} from the normal (y = n x (1,0,0) is NOT normalised, as in the reference), cov = ((R*S)*S)*R^-1 with
 * S = diag(1, 1, scaleUncertaintyNormal).  Row-major 3x3 out. */
static void normalize3d(double *v) {
    double yy = v[1] * v[1], zz = v[2] * v[2];
    double n = sqrt(v[0] * v[0] + (yy + zz));
    v[0] /= n; v[1] /= n; v[2] /= n;
}
static void matmul3d(const double *A, const double *B, double *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            /* fixed inner size 3: c0 + (c1 + c2) (Eigen 3.3 unrolled redux, see transform_pt) */
            C[3 * i + j] = A[3 * i] * B[j] + (A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j]);
        }
}
ORC_API void orc_uncertainty_from_normal(const double *normal_in, double scale, double *cov) {
    double n[3] = {normal_in[0], normal_in[1], normal_in[2]};
    double x[3] = {1, 0, 0}, y[3];
    y[0] = n[1] * x[2] - n[2] * x[1]; y[1] = n[2] * x[0] - n[0] * x[2]; y[2] = n[0] * x[1] - n[1] * x[0];
    normalize3d(n);
    x[0] = y[1] * n[2] - y[2] * n[1]; x[1] = y[2] * n[0] - y[0] * n[2]; x[2] = y[0] * n[1] - y[1] * n[0];
    normalize3d(x);
    double R[9] = {x[0], y[0], n[0], x[1], y[1], n[1], x[2], y[2], n[2]};
    double S[9] = {1, 0, 0, 0, 1, 0, 0, 0, scale};
    double RS[9], RSS[9], Rinv[9];
    matmul3d(R, S, RS);
    matmul3d(RS, S, RSS);
    inverse3d(R, Rinv);
    matmul3d(RSS, Rinv, cov);
}

/* cov.inverse() of models 1 / 2 (src/Map/featuresMap.cpp:115-120, 268-273): Eigen's 3x3 cofactor inverse. */
ORC_API void orc_inverse3d(const double *m, double *r) { inverse3d(m, r); }

/* RGBD::computeRGBGradient (src/RGBD/RGBD.cpp:147-187).  rgb: H rows of rgb_row_bytes bytes, 3 bytes per pixel.
 * The 3x3 patch is a view at pixel (u-1, v-1) of the colour image, but it is read with at<uint16_t>(r, c): the
 * little-endian 16-bit word at byte offset 2c of patch row r (:156-159) -- reproduced literally.  The Scharr sums are
 * int arithmetic (uint16 promotes to int), stored in doubles.  The cvtColor result (:151-152) is never used.
 * Returns the un-normalised (1,1,1) when the border test (:154) fails. */
ORC_API void orc_compute_rgb_gradient(const uint8_t *rgb, int rgb_row_bytes, const uint16_t *depth, int W, int H,
                                      int stride, int u, int v, float fx, float fy, float cx, float cy,
                                      double depth_scale, double *grad) {
    if (!((u - 1 > 0) && (v - 1 > 0) && (u + 1 < W) && (v + 1 < H))) {
        grad[0] = 1; grad[1] = 1; grad[2] = 1;
        return;
    }
    int p[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            const uint8_t *b = rgb + (size_t)(v - 1 + r) * rgb_row_bytes + 3 * (size_t)(u - 1) + 2 * c;
            p[r][c] = (int)b[0] | ((int)b[1] << 8);
        }
    double gradx = -3 * p[0][0] - 10 * p[0][1] - 3 * p[0][2] + +3 * p[2][0] + 10 * p[2][1] + 3 * p[2][2];
    double grady = -3 * p[0][0] - 10 * p[1][0] - 3 * p[2][0] + +3 * p[0][2] + 10 * p[1][2] + 3 * p[2][2];
    double angle = atan2(grady, gradx) + (M_PI / 2.0);
    int coord1[2] = {(int)(sqrt(2) * sin(angle)), (int)(sqrt(2) * cos(angle))};
    int coord2[2] = {(int)(sqrt(2) * sin(angle + M_PI)), (int)(sqrt(2) * cos(angle + M_PI))};
    float pc[3], pe[3], pb[3];
    point2Dto3D((float)u, (float)v, depth, W, H, stride, fx, fy, cx, cy, depth_scale, pc);
    point2Dto3D((float)(u + coord1[0]), (float)(v + coord1[1]), depth, W, H, stride, fx, fy, cx, cy, depth_scale, pe);
    point2Dto3D((float)(u + coord2[0]), (float)(v + coord2[1]), depth, W, H, stride, fx, fy, cx, cy, depth_scale, pb);
    double g[3];
    if (pe[2] > 0 && pb[2]) {
        g[0] = (double)(pe[0] - pb[0]); g[1] = (double)(pe[1] - pb[1]); g[2] = (double)(pe[2] - pb[2]);
    } else if (pc[2] > 0 && pb[2]) {
        g[0] = (double)(pc[0] - pb[0]); g[1] = (double)(pc[1] - pb[1]); g[2] = (double)(pc[2] - pb[2]);
    } else if (pc[2] > 0 && pe[2]) {
        g[0] = (double)(pe[0] - pc[0]); g[1] = (double)(pe[1] - pc[1]); g[2] = (double)(pe[2] - pc[2]);
    } else {
        g[0] = coord1[0]; g[1] = coord1[1]; g[2] = 0;
    }
    normalize3d(g);
    grad[0] = g[0]; grad[1] = g[1]; grad[2] = g[2];
}

/* The four libm-sensitive direction pairs of computeRGBGradient: |gradx| == |grady| != 0, where sqrt(2)*sin(angle)
 * is 1 -+ 1 ulp and the int() truncation depends on the last bit.  out[q] = {c1x, c1y, c2x, c2y} for
 * q = (gx < 0) + 2 (gy < 0), evaluated with this host's libm (the library builds the same table at context creation). */
ORC_API void orc_gradient_diag_table(int *out) {
    for (int q = 0; q < 4; ++q) {
        volatile double gx = (q & 1) ? -1.0 : 1.0, gy = (q & 2) ? -1.0 : 1.0;   /* volatile: no compile-time folding */
        double angle = atan2(gy, gx) + (M_PI / 2.0);
        out[4 * q + 0] = (int)(sqrt(2) * sin(angle));
        out[4 * q + 1] = (int)(sqrt(2) * cos(angle));
        out[4 * q + 2] = (int)(sqrt(2) * sin(angle + M_PI));
        out[4 * q + 3] = (int)(sqrt(2) * cos(angle + M_PI));
    }
}

/* DepthSensorModel::uncertinatyFromRGBGradient (src/Grabber/depthSensorModel.cpp:79-95): y = (0,0,1) x grad,
 * normalised; z = y x grad, normalised; R = [grad y z]; S = diag(1, scaleUncertaintyGradient, 1) (the reference
 * assigns S(1,1) twice); cov = ((R*S)*S)*R^-1.  Row-major 3x3 out. */
ORC_API void orc_uncertainty_from_gradient(const double *grad, double scale, double *cov) {
    const double g[3] = {grad[0], grad[1], grad[2]};
    double z[3] = {0, 0, 1}, y[3];
    y[0] = z[1] * g[2] - z[2] * g[1]; y[1] = z[2] * g[0] - z[0] * g[2]; y[2] = z[0] * g[1] - z[1] * g[0];
    normalize3d(y);
    z[0] = y[1] * g[2] - y[2] * g[1]; z[1] = y[2] * g[0] - y[0] * g[2]; z[2] = y[0] * g[1] - y[1] * g[0];
    normalize3d(z);
    double R[9] = {g[0], y[0], z[0], g[1], y[1], z[1], g[2], y[2], z[2]};
    double S[9] = {1, 0, 0, 0, scale, 0, 0, 0, 1};
    double RS[9], RSS[9], Rinv[9];
    matmul3d(R, S, RS);
    matmul3d(RS, S, RSS);
    inverse3d(R, Rinv);
    matmul3d(RSS, Rinv, cov);
}

/* Map-side preparation of the frame-to-map matcher's input, i.e. the body of
 * PUTSLAM::getAndFilterFeaturesFromMap (src/PUTSLAM/PUTSLAM.cpp:624-674) after getCovisibleFeatures:
 *   FeaturesMap::findNearestFrame (src/Map/featuresMap.cpp:528-563)      angle between the current optical axis and
 *                                                                        the optical axis of the view that holds the
 *                                                                        feature's descriptor (one view per feature,
 *                                                                        SURVEY B#13); float dot / norms, acos
 *   removeMapFeaturesWithoutGoodObservationAngle (PUTSLAM.cpp:932-950)    drop angle > maxAngle (or NaN)
 *   moveMapFeaturesToLocalCordinateSystem (PUTSLAM.cpp:28-51)             p_local = cameraPose^-1 * p (double, Eigen
 *                                                                        affine inverse), (u, v) = inverseModel
 *   DepthSensorModel::inverseModel (depthSensorModel.cpp:18-25)           projection with validity box else (-1,-1,-1)
 *   RGBD::removeFarMapFeatures (src/RGBD/RGBD.cpp:232-252)                drop z_local > maxZ
 * pose: column-major 4x4 (camera -> global).  view_axis: M x 3 float, third column of the rotation of the feature's
 * view.  Outputs are compacted in input order; returns the number kept. */
ORC_API int orc_map_prepare(const double *xyz, const float *view_axis, int M, const double *pose, double fx, double fy,
                            double cx, double cy, double img_w, double img_h, double max_angle, double max_z,
                            int *kept, double *xyz_local, double *uv, double *angles) {
    double L[9], Li[9], ti[3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) L[3 * r + c] = pose[4 * c + r];
    inverse3d(L, Li);
    for (int r = 0; r < 3; ++r) {   /* translation of the inverse: -(Linv * t) */
        double s = Li[3 * r] * pose[12];
        s = s + Li[3 * r + 1] * pose[13];
        s = s + Li[3 * r + 2] * pose[14];
        ti[r] = -s;
    }
    const float zc[3] = {(float)pose[8], (float)pose[9], (float)pose[10]};
    int n = 0;
    for (int i = 0; i < M; ++i) {
        const float *zp = view_axis + 3 * i;
        float dot = zp[0] * zc[0] + (zp[1] * zc[1] + zp[2] * zc[2]);
        float na = norm3f(zp[0], zp[1], zp[2]), nb = norm3f(zc[0], zc[1], zc[2]);
        double angle = acos((double)(dot / (na * nb)));
        if (!(fabs(angle) < 10)) continue;            /* NaN: no view selected -> imageId -1 -> removed */
        if (fabs(angle) > max_angle) continue;
        double p[3];
        for (int r = 0; r < 3; ++r) {
            double s = Li[3 * r] * xyz[3 * i];
            s = s + Li[3 * r + 1] * xyz[3 * i + 1];
            s = s + Li[3 * r + 2] * xyz[3 * i + 2];
            p[r] = s + ti[r];
        }
        if (p[2] > max_z) continue;
        double u = ((fx * p[0]) / p[2]) + cx, v = ((fy * p[1]) / p[2]) + cy;
        if (u < 0 || u > img_w || v < 0 || v > img_h || p[2] < 0.8 || p[2] > 6.0) { u = -1; v = -1; }
        kept[n] = i;
        xyz_local[3 * n] = p[0]; xyz_local[3 * n + 1] = p[1]; xyz_local[3 * n + 2] = p[2];
        uv[2 * n] = u; uv[2 * n + 1] = v;
        angles[n] = fabs(angle);
        ++n;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Stage 3 -- SVD, Umeyama, Kabsch, RANSAC
 * ---------------------------------------------------------------------------------------- */

/* Two-sided Jacobi SVD of a 3x3 matrix, following Eigen 3.3's JacobiSVD<Matrix> for square
 * real input (QR preconditioner is a no-op): scale by max|a_ij|, sweep pairs (1,0),(2,0),(2,1)
 * until no off-diagonal exceeds max(min_normal, 2*eps*maxDiag), 2x2 kernel =
 * real_2x2_jacobi_svd + JacobiRotation::makeJacobi, make singular values positive, sort
 * descending with column swaps.  Row-major storage.  Third-party arithmetic (Eigen,
 * un-vendored; call sites RANSAC.cpp:225, kabschEst.cpp:47) -- see PARITY STATUS above. */
#define DEFINE_SVD3(NAME, REAL, SQRT, FABS, EPSV, MINV)                                            \
    static void NAME##_rot_rows(REAL *M, int p, int q, REAL c, REAL s) {                           \
        if (c == (REAL)1 && s == (REAL)0) return;                                                  \
        for (int i = 0; i < 3; ++i) {                                                              \
            REAL xi = M[3 * p + i], yi = M[3 * q + i];                                             \
            M[3 * p + i] = c * xi + s * yi;                                                        \
            M[3 * q + i] = -s * xi + c * yi;                                                       \
        }                                                                                          \
    }                                                                                              \
    static void NAME##_rot_cols(REAL *M, int p, int q, REAL c, REAL s) {                           \
        /* applyOnTheRight(p,q,j): in-plane rotation of columns with j.transpose() = (c,-s) */     \
        REAL st = -s;                                                                              \
        if (c == (REAL)1 && st == (REAL)0) return;                                                 \
        for (int i = 0; i < 3; ++i) {                                                              \
            REAL xi = M[3 * i + p], yi = M[3 * i + q];                                             \
            M[3 * i + p] = c * xi + st * yi;                                                       \
            M[3 * i + q] = -st * xi + c * yi;                                                      \
        }                                                                                          \
    }                                                                                              \
    static void NAME(const REAL *A, REAL *U, REAL *S, REAL *V) {                                   \
        REAL W[9];                                                                                 \
        REAL scale = 0;                                                                            \
        for (int i = 0; i < 9; ++i) { REAL a = FABS(A[i]); if (a > scale) scale = a; }             \
        if (scale == (REAL)0) scale = (REAL)1;                                                     \
        for (int i = 0; i < 9; ++i) { W[i] = A[i] / scale; U[i] = V[i] = (i % 4 == 0) ? (REAL)1 : (REAL)0; } \
        const REAL precision = (REAL)2 * EPSV;                                                     \
        const REAL considerAsZero = MINV;                                                          \
        REAL maxDiag = FABS(W[0]);                                                                 \
        if (FABS(W[4]) > maxDiag) maxDiag = FABS(W[4]);                                            \
        if (FABS(W[8]) > maxDiag) maxDiag = FABS(W[8]);                                            \
        int finished = 0, guard = 0;                                                               \
        while (!finished && guard++ < 1000) {                                                      \
            finished = 1;                                                                          \
            for (int p = 1; p < 3; ++p)                                                            \
                for (int q = 0; q < p; ++q) {                                                      \
                    REAL thr = precision * maxDiag;                                                \
                    if (considerAsZero > thr) thr = considerAsZero;                                \
                    if (FABS(W[3 * p + q]) > thr || FABS(W[3 * q + p]) > thr) {                    \
                        finished = 0;                                                              \
                        /* real_2x2_jacobi_svd */                                                  \
                        REAL m00 = W[3 * p + p], m01 = W[3 * p + q], m10 = W[3 * q + p], m11 = W[3 * q + q]; \
                        REAL r1c, r1s;                                                             \
                        REAL t = m00 + m11, d = m10 - m01;                                         \
                        if (FABS(d) < MINV) { r1s = 0; r1c = 1; }                                  \
                        else { REAL u = t / d; REAL tmp = SQRT((REAL)1 + u * u); r1s = (REAL)1 / tmp; r1c = u / tmp; } \
                        if (!(r1c == (REAL)1 && r1s == (REAL)0)) {                                 \
                            REAL x0 = m00, y0 = m10, x1 = m01, y1 = m11;                           \
                            m00 = r1c * x0 + r1s * y0; m10 = -r1s * x0 + r1c * y0;                 \
                            m01 = r1c * x1 + r1s * y1; m11 = -r1s * x1 + r1c * y1;                 \
                        }                                                                          \
                        /* j_right.makeJacobi(m00, m01, m11) */                                    \
                        REAL jrc, jrs;                                                             \
                        REAL deno = (REAL)2 * FABS(m01);                                           \
                        if (deno < MINV) { jrc = 1; jrs = 0; }                                     \
                        else {                                                                     \
                            REAL tau = (m00 - m11) / deno;                                         \
                            REAL w = SQRT(tau * tau + (REAL)1);                                    \
                            REAL tt;                                                               \
                            if (tau > (REAL)0) tt = (REAL)1 / (tau + w); else tt = (REAL)1 / (tau - w); \
                            REAL sign_t = tt > (REAL)0 ? (REAL)1 : (REAL)-1;                       \
                            REAL nn = (REAL)1 / SQRT(tt * tt + (REAL)1);                           \
                            jrs = -sign_t * (m01 / FABS(m01)) * FABS(tt) * nn;                     \
                            jrc = nn;                                                              \
                        }                                                                          \
                        /* j_left = rot1 * j_right.transpose() */                                  \
                        REAL jtc = jrc, jts = -jrs;                                                \
                        REAL jlc = r1c * jtc - r1s * jts;                                          \
                        REAL jls = r1c * jts + r1s * jtc;                                          \
                        NAME##_rot_rows(W, p, q, jlc, jls);      /* W.applyOnTheLeft(p,q,j_left) */ \
                        NAME##_rot_cols(U, p, q, jlc, -jls);     /* U.applyOnTheRight(p,q,j_left^T) */ \
                        NAME##_rot_cols(W, p, q, jrc, jrs);      /* W.applyOnTheRight(p,q,j_right) */ \
                        NAME##_rot_cols(V, p, q, jrc, jrs);      /* V.applyOnTheRight(p,q,j_right) */ \
                        REAL a = FABS(W[3 * p + p]), b = FABS(W[3 * q + q]);                       \
                        if (b > a) a = b;                                                          \
                        if (a > maxDiag) maxDiag = a;                                              \
                    }                                                                              \
                }                                                                                  \
        }                                                                                          \
        for (int i = 0; i < 3; ++i) {                                                              \
            REAL a = W[4 * i];                                                                     \
            S[i] = FABS(a);                                                                        \
            if (a < (REAL)0) for (int r = 0; r < 3; ++r) U[3 * r + i] = -U[3 * r + i];             \
        }                                                                                          \
        for (int i = 0; i < 3; ++i) S[i] = S[i] * scale;                                           \
        for (int i = 0; i < 3; ++i) {                                                              \
            int pos = i; REAL mx = S[i];                                                           \
            for (int k = i + 1; k < 3; ++k) if (S[k] > mx) { mx = S[k]; pos = k; }                 \
            if (mx == (REAL)0) break;                                                              \
            if (pos != i) {                                                                        \
                REAL tmp = S[i]; S[i] = S[pos]; S[pos] = tmp;                                      \
                for (int r = 0; r < 3; ++r) {                                                      \
                    tmp = U[3 * r + i]; U[3 * r + i] = U[3 * r + pos]; U[3 * r + pos] = tmp;       \
                    tmp = V[3 * r + i]; V[3 * r + i] = V[3 * r + pos]; V[3 * r + pos] = tmp;       \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
    static __attribute__((unused)) REAL NAME##_det3(const REAL *m) {                                                       \
        /* Eigen bruteforce_det3_helper order */                                                   \
        REAL a = m[0] * (m[4] * m[8] - m[5] * m[7]);                                               \
        REAL b = m[1] * (m[3] * m[8] - m[5] * m[6]);                                               \
        REAL c = m[2] * (m[3] * m[7] - m[4] * m[6]);                                               \
        return a - b + c;                                                                          \
    }

DEFINE_SVD3(svd3f, float, sqrtf, fabsf, FLT_EPSILON, FLT_MIN)
DEFINE_SVD3(svd3d, double, sqrt, fabs, DBL_EPSILON, DBL_MIN)

ORC_API void orc_svd3f(const float *A, float *U, float *S, float *V) { svd3f(A, U, S, V); }
ORC_API void orc_svd3d(const double *A, double *U, double *S, double *V) { svd3d(A, U, S, V); }

/* Eigen::umeyama(src, dst, with_scaling=false) in float32 for 3-D points
 * (call site src/TransformEst/RANSAC.cpp:225-226: src = current features, dst = previous).
 * Umeyama 1991 eq. 38-43 in the order Eigen 3.3 evaluates them:
 *   one_over_n = 1.f/n;  means = (sequential row sums) * one_over_n;
 *   sigma = one_over_n * sum_k dst_demean(:,k) * src_demean(:,k)^T   (sequential over k);
 *   S(2) = -1 iff det(U)*det(V) < 0;  R = (U*S)*V^T (sequential inner sum);
 *   t = dst_mean - R*src_mean.
 * src/dst are given through index lists so the caller need not gather.  T is row-major 4x4.
 * Returns 0 and sets T = identity if T(0,0) is NaN (RANSAC.cpp:238-242), else 1. */
static int umeyama_idx(const float *dst_pts, const int *dst_idx, const float *src_pts, const int *src_idx,
                       int n, float *T) {
    const float one_over_n = 1.f / (float)n;
    float sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
    for (int k = 0; k < n; ++k)
        for (int c = 0; c < 3; ++c) {
            sm[c] = sm[c] + src_pts[3 * src_idx[k] + c];
            dm[c] = dm[c] + dst_pts[3 * dst_idx[k] + c];
        }
    for (int c = 0; c < 3; ++c) { sm[c] = sm[c] * one_over_n; dm[c] = dm[c] * one_over_n; }
    float sig[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < n; ++k) {
        float sd[3], dd[3];
        for (int c = 0; c < 3; ++c) {
            sd[c] = src_pts[3 * src_idx[k] + c] - sm[c];
            dd[c] = dst_pts[3 * dst_idx[k] + c] - dm[c];
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) sig[3 * i + j] = sig[3 * i + j] + dd[i] * sd[j];
    }
    for (int i = 0; i < 9; ++i) sig[i] = one_over_n * sig[i];
    float U[9], S[3], V[9];
    svd3f(sig, U, S, V);
    float sgn[3] = {1.f, 1.f, 1.f};
    if (svd3f_det3(U) * svd3f_det3(V) < 0.f) sgn[2] = -1.f;
    float R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = (U[3 * i + 0] * sgn[0]) * V[3 * j + 0];
            s = s + (U[3 * i + 1] * sgn[1]) * V[3 * j + 1];
            s = s + (U[3 * i + 2] * sgn[2]) * V[3 * j + 2];
            R[3 * i + j] = s;
        }
    float t[3];
    for (int i = 0; i < 3; ++i) {
        float s = R[3 * i + 0] * sm[0];
        s = s + R[3 * i + 1] * sm[1];
        s = s + R[3 * i + 2] * sm[2];
        t[i] = dm[i] - s;
    }
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
    if (n <= 0 || isnan(R[0])) return 0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = t[i];
    }
    return 1;
}

/* MatrixXd::determinant() of a DYNAMIC-size matrix = partialPivLu().determinant() (Eigen LU/Determinant.h
 * determinant_impl<Derived, Dynamic>; LU/PartialPivLU.h unblocked_lu): first-max partial pivoting, column divided by
 * the pivot, rank-1 update, sign * ((u00 * u11) * u22).  This is what kabschEst.cpp:53 evaluates (A is a MatrixXd);
 * it only matters for rank-deficient H, where the sign of det H is rounding noise.  Row-major 3x3. */
static double det3d_lu(const double *m) {
    double a[9];
    memcpy(a, m, sizeof(a));
    double sign = 1.0;
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double best = fabs(a[3 * k + k]);
        for (int i = k + 1; i < 3; ++i) if (fabs(a[3 * i + k]) > best) { best = fabs(a[3 * i + k]); p = i; }
        if (p != k) {
            for (int j = 0; j < 3; ++j) { double t = a[3 * k + j]; a[3 * k + j] = a[3 * p + j]; a[3 * p + j] = t; }
            sign = -sign;
        }
        if (a[3 * k + k] == 0.0) continue;
        for (int i = k + 1; i < 3; ++i) {
            a[3 * i + k] = a[3 * i + k] / a[3 * k + k];
            for (int j = k + 1; j < 3; ++j) a[3 * i + j] = a[3 * i + j] - a[3 * i + k] * a[3 * k + j];
        }
    }
    return ((sign * a[0]) * a[4]) * a[8];
}

/* Plain-array entry: src, dst are n x 3 float; T row-major 4x4 with dst ~= R*src + t. */
ORC_API int orc_umeyama(const float *src, const float *dst, int n, float *T) {
    int *idx = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) idx[i] = i;
    int ok = umeyama_idx(dst, idx, src, idx, n, T);
    free(idx);
    return ok;
}

/* KabschEst::computeTransformation(setA, setB) (src/TransformEst/kabschEst.cpp:24-68), double.
 * A, B are n x 3 row-major here (the adapter converts Eigen's column-major MatrixXd).
 * T is row-major 3x4 with B ~= R*A + t.  Empty input -> identity (:28). */
ORC_API void orc_kabsch(const double *A, const double *B, int n, double *T) {
    for (int i = 0; i < 12; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (n == 0) return;
    double cA[3], cB[3];
    for (int c = 0; c < 3; ++c) {                 /* :31-34 col(i).mean() = sum / n */
        double sa = 0, sb = 0;
        for (int k = 0; k < n; ++k) { sa = sa + A[3 * k + c]; sb = sb + B[3 * k + c]; }
        cA[c] = sa / (double)n; cB[c] = sb / (double)n;
    }
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};     /* :37-44  H = Anew^T * Bnew */
    for (int k = 0; k < n; ++k) {
        double a[3], b[3];
        for (int c = 0; c < 3; ++c) { a[c] = A[3 * k + c] - cA[c]; b[c] = B[3 * k + c] - cB[c]; }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) H[3 * i + j] = H[3 * i + j] + a[i] * b[j];
    }
    double Us[9], S[3], Vs[9];
    svd3d(H, Us, S, Vs);                           /* :47-49  V = matrixU, W = matrixV */
    double det = det3d_lu(H);                      /* :52-53 (dynamic-size determinant) */
    double d = (det != 0) ? det : 1;
    double sg = (double)((d > 0) - (d < 0));
    double D[3] = {1.0, 1.0, sg};
    double R[9];                                   /* :56  U = W * U * V^T */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            /* (W * U) * V^T: both products have a fixed inner size of 3 -> c0 + (c1 + c2); W * diag(1,1,d) is exact */
            double a = (Vs[3 * i + 0] * D[0]) * Us[3 * j + 0];
            double b = (Vs[3 * i + 1] * D[1]) * Us[3 * j + 1];
            double c = (Vs[3 * i + 2] * D[2]) * Us[3 * j + 2];
            R[3 * i + j] = a + (b + c);
        }
    for (int i = 0; i < 3; ++i) {                  /* :59-62  T = U*(-cA) + cB */
        double a = R[3 * i + 0] * (-cA[0]);
        double b = R[3 * i + 1] * (-cA[1]);
        double c = R[3 * i + 2] * (-cA[2]);
        double s = (a + (b + c)) + cB[i];
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[3 * i + j];
        T[4 * i + 3] = s;
    }
}

/* Philox4x32-10 (Salmon et al., SC'11; Random123) -- the counter-based RNG north_star asks
 * for in place of the reference's srand(time(0)) / rand() (RANSAC.cpp:13,191). */
ORC_API void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
    uint32_t k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* RANSAC::getRandomMatches (src/TransformEst/RANSAC.cpp:180-205): draw `usedPairs`=3 distinct
 * indices by `r % m` with rejection.  The draw stream for hypothesis h is
 * Philox(ctr = {block, h, 0, 0}, key = {seed_lo, seed_hi}) words 0..3 of block 0, 1, ... */
ORC_API void orc_sample3(uint64_t seed, uint32_t h, int m, int *out3) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int got = 0;
    for (uint32_t blk = 0; got < 3; ++blk) {
        uint32_t ctr[4] = {blk, h, 0u, 0u}, r[4];
        orc_philox4x32_10(ctr, key, r);
        for (int w = 0; w < 4 && got < 3; ++w) {
            int idx = (int)(r[w] % (uint32_t)m);
            int dup = 0;
            for (int a = 0; a < got; ++a) if (out3[a] == idx) dup = 1;
            if (!dup) out3[got++] = idx;
        }
    }
}

/* RANSAC::computeRANSACIteration (RANSAC.cpp:457-461), p = 0.98, k = 3 (RANSAC.h:195-196).
 * For tiny inlier ratios (w^3 < ~1.8e-9, e.g. 1 inlier out of > 820 matches) the quotient exceeds
 * INT_MAX and the reference's double -> int conversion is UNDEFINED BEHAVIOUR (x86-64 cvttsd2si happens
 * to give INT_MIN, which ends the loop at once; aarch64 fcvtzs saturates).  Undefined behaviour cannot
 * be a parity target, so the restatement and the product both define it as the SATURATING conversion
 * (the mathematically meant value: "a huge number of iterations", which min() then caps at 487).
 * Deliberate divergence, listed in DESIGN.md. */
static int ransac_iterations(double inlierRatio) {
    double v = log(1 - 0.98) / log(1 - pow(inlierRatio, 3));
    if (v != v) return INT_MIN;
    if (v >= 2147483648.0) return INT_MAX;
    if (v <= -2147483649.0) return INT_MIN;
    return (int)v;
}
ORC_API int orc_ransac_iterations(double w) { return ransac_iterations(w); }

/* USAC<T>::updateStandardStopping (include/putslam/USAC/USAC.h:944-971) -- the standard RANSAC stopping criterion
 * of the (not compiled) USAC framework, offered as an alternative termination rule:
 * prob = prod_{i<k} (I - i)/(N - i);  ceil(log(1 - conf) / log(1 - prob)), max_hyp if prob < eps, 1 if 1-prob < eps.
 * `numInliers - i` is unsigned arithmetic in the reference. */
static int g_stop_rule = 0;        /* 0 = RANSAC::computeRANSACIteration (reference default), 1 = USAC standard */
static double g_usac_conf = 0.99;  /* USAC_wrapper.cpp:66 */
ORC_API void orc_set_stopping(int rule, double conf) { g_stop_rule = rule; g_usac_conf = conf; }
static unsigned usac_standard_stopping(unsigned numInliers, unsigned totPoints, unsigned sampleSize, unsigned max_hyp) {
    double n_inliers = 1.0, n_pts = 1.0;
    for (unsigned i = 0; i < sampleSize; ++i) {
        n_inliers *= numInliers - i;
        n_pts *= totPoints - i;
    }
    double prob_good_model = n_inliers / n_pts;
    if (prob_good_model < DBL_EPSILON) return max_hyp;
    else if (1 - prob_good_model < DBL_EPSILON) return 1;
    double nusample_s = log(1 - g_usac_conf) / log(1 - prob_good_model);
    return (unsigned)ceil(nusample_s);
}
ORC_API unsigned orc_usac_stopping(unsigned inl, unsigned tot, unsigned max_hyp) { return usac_standard_stopping(inl, tot, 3, max_hyp); }

typedef struct {
    int error_version;            /* RANSAC::ERROR_VERSION: 0 Euclidean, 1 reproj, 2 both, 4 adaptive */
    double inlier_threshold_euclidean;
    double inlier_threshold_reprojection;
    double minimal_inlier_ratio_threshold;
    int minimal_number_of_matches;
    int used_pairs;               /* must be 3 */
    float fx, fy, cx, cy;         /* cameraMatrix (CV_32F) for the reprojection metrics */
} orc_ransac_params;

/* RGBD::point3Dto2D (src/RGBD/RGBD.cpp:92-98), float32 in that operation order. */
static inline void point3Dto2D(const float *p, const orc_ransac_params *P, float *u, float *v) {
    *u = p[0] * P->fx / p[2] + P->cx;
    *v = p[1] * P->fy / p[2] + P->cy;
}

/* Eigen Matrix4f::inverse() for the generic (non-vectorised) path: cofactors divided by the
 * determinant expanded along column 0 (Eigen/src/LU/InverseImpl.h compute_inverse_size4).
 * Third-party arithmetic; used twice per call by the reprojection metrics (RANSAC.cpp:337-338). */
static float det3_helper(const float *m, int i1, int i2, int i3, int j1, int j2, int j3) {
    return m[4 * i1 + j1] * (m[4 * i2 + j2] * m[4 * i3 + j3] - m[4 * i2 + j3] * m[4 * i3 + j2]);
}
static float cofactor4(const float *m, int i, int j) {
    int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4;
    int j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
    return det3_helper(m, i1, i2, i3, j1, j2, j3) + det3_helper(m, i2, i3, i1, j1, j2, j3) +
           det3_helper(m, i3, i1, i2, j1, j2, j3);
}
static void inverse4(const float *m, float *r) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float c = cofactor4(m, i, j);
            r[4 * j + i] = ((i + j) & 1) ? -c : c;
        }
    /* (matrix.col(0).cwiseProduct(result.row(0).transpose())).sum(): size-4 unrolled redux
     * splits in halves -> (a0 + a1) + (a2 + a3) */
    float det = (m[0] * r[0] + m[4] * r[1]) + (m[8] * r[2] + m[12] * r[3]);
    for (int i = 0; i < 16; ++i) r[i] = r[i] / det;
}
ORC_API void orc_inverse4(const float *m, float *r) { inverse4(m, r); }

/* R*x + t with Eigen 3.3's 3x3 coefficient-based product: coeff = (row . col).sum(), and a fixed-size sum of three
 * terms is the unrolled binary split c0 + (c1 + c2) (Core/ProductEvaluators.h, Core/Redux.h redux_novec_unroller;
 * the reference needs Eigen >= 3.3: it uses Eigen::Index, transformEst.h:45).  Pinned to the reference's compiled
 * RANSAC.cpp through oracle/_ref (tests/test_ref_build_cpu.py); the product order itself is the stand-in's model. */
static inline void transform_pt(const float *T, const float *x, float *o) {
    for (int i = 0; i < 3; ++i) {
        float a = T[4 * i + 0] * x[0];
        float b = T[4 * i + 1] * x[1];
        float c = T[4 * i + 2] * x[2];
        float s = a + (b + c);
        o[i] = s + T[4 * i + 3];
    }
}

/* One inlier test for match (prev point p, current point c) under model T:
 * computeMatchInlierRatioEuclidean (RANSAC.cpp:251-281), computeInlierRatioReprojection
 * (:325-375), computeInlierRatioEuclideanAndReprojection (:377-436).  Tinv is only read for
 * error versions 1 and 2.  `force_euclid` reproduces the refit recount, which always calls the
 * Euclidean routine (:155) -- that routine still applies the ADAPTIVE scaling (:270-271). */
static int is_inlier(const float *T, const float *Tinv, const float *p, const float *c,
                     const orc_ransac_params *P, int force_euclid) {
    int ev = P->error_version;
    float est_old[3];
    transform_pt(T, c, est_old);
    if (force_euclid || ev == 0 || ev == 4) {
        double threshold = P->inlier_threshold_euclidean;
        if (ev == 4) threshold *= p[2];
        float nrm = norm3f(est_old[0] - p[0], est_old[1] - p[1], est_old[2] - p[2]);
        return (double)nrm < threshold;
    }
    float est_new[3];
    transform_pt(Tinv, p, est_new);
    float pnu, pnv, rnu, rnv, pou, pov, rou, rov;
    point3Dto2D(est_new, P, &pnu, &pnv);
    point3Dto2D(c, P, &rnu, &rnv);
    point3Dto2D(est_old, P, &pou, &pov);
    point3Dto2D(p, P, &rou, &rov);
    /* cv::norm(Point2f) = sqrt((double)x*x + (double)y*y) */
    float ax = pnu - rnu, ay = pnv - rnv, bx = pou - rou, by = pov - rov;
    double e0 = sqrt((double)ax * ax + (double)ay * ay);
    double e1 = sqrt((double)bx * bx + (double)by * by);
    int ok2d = e0 < P->inlier_threshold_reprojection && e1 < P->inlier_threshold_reprojection;
    if (ev == 1) return ok2d;
    double error3D = norm3f(est_old[0] - p[0], est_old[1] - p[1], est_old[2] - p[2]);
    return error3D < P->inlier_threshold_euclidean && ok2d;
}

/* RANSAC::estimateTransformation (src/TransformEst/RANSAC.cpp:50-174) with the sample stream
 * of orc_sample3.  num_hyp == 0 : the reference's adaptive loop (bound starts at 487 and
 * shrinks through saveBetterModel, :438-455).  num_hyp > 0 : the bound is pinned to num_hyp
 * ("fixed-H" mode).
 *   prev, cur   : n1 x 3, n2 x 3 float;  mq/mt : m matches (queryIdx -> prev, trainIdx -> cur)
 *   T_out       : row-major 4x4 (prev ~= R*cur + t)
 *   inl_idx     : indices into the ORIGINAL match list of the final inliers (ascending)
 *   counts_out  : optional, per evaluated hypothesis inlier count (-1 = model invalid / not run)
 * Returns 0. */
ORC_API int orc_ransac(const float *prev, int n1, const float *cur, int n2, const int *mq, const int *mt,
                       int m, const orc_ransac_params *P, uint64_t seed, int num_hyp, float *T_out,
                       int *inl_idx, int *n_inl_out, double *best_ratio_out, int *hyp_used_out,
                       int *counts_out, int counts_cap) {
    (void)n1; (void)n2;
    for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.f : 0.f;
    *n_inl_out = 0;
    if (best_ratio_out) *best_ratio_out = 0.0;
    if (hyp_used_out) *hyp_used_out = 0;
    /* :65-74 drop matches with NaN or z outside [0.1, 6] */
    int *keep = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    int mf = 0;
    for (int k = 0; k < m; ++k) {
        const float *p = prev + 3 * mq[k], *c = cur + 3 * mt[k];
        int bad = isnan(p[0]) || isnan(p[1]) || isnan(p[2]) || isnan(c[0]) || isnan(c[1]) || isnan(c[2]) ||
                  p[2] < 0.1 || p[2] > 6 || c[2] < 0.1 || c[2] > 6;
        if (!bad) keep[mf++] = k;
    }
    if (mf < P->minimal_number_of_matches) { free(keep); return 0; }   /* :77-80 */
    int *fq = (int *)malloc(sizeof(int) * (size_t)mf), *ft = (int *)malloc(sizeof(int) * (size_t)mf);
    for (int k = 0; k < mf; ++k) { fq[k] = mq[keep[k]]; ft[k] = mt[keep[k]]; }

    int iterationCount = num_hyp > 0 ? num_hyp : ransac_iterations(0.20);   /* :30 */
    double bestInlierRatio = 0.0;
    float bestT[16];
    for (int i = 0; i < 16; ++i) bestT[i] = (i % 5 == 0) ? 1.f : 0.f;
    int *best_inl = (int *)malloc(sizeof(int) * (size_t)mf), n_best = 0;
    int *cur_inl = (int *)malloc(sizeof(int) * (size_t)mf);
    int i;
    for (i = 0; i < iterationCount; ++i) {                               /* :87 */
        int s[3];
        orc_sample3(seed, (uint32_t)i, mf, s);                           /* :95 */
        int sq[3] = {fq[s[0]], fq[s[1]], fq[s[2]]}, st[3] = {ft[s[0]], ft[s[1]], ft[s[2]]};
        float T[16], Tinv[16];
        int ok = umeyama_idx(prev, sq, cur, st, 3, T);                   /* :102 */
        if (counts_out && i < counts_cap) counts_out[i] = -1;
        if (!ok) continue;
        if (P->error_version == 1 || P->error_version == 2) inverse4(T, Tinv);
        int cnt = 0;
        for (int k = 0; k < mf; ++k)
            if (is_inlier(T, Tinv, prev + 3 * fq[k], cur + 3 * ft[k], P, 0)) cur_inl[cnt++] = k;
        if (counts_out && i < counts_cap) counts_out[i] = cnt;
        float inlierRatio = (float)cnt / (float)mf;                      /* :280 */
        if ((double)inlierRatio > bestInlierRatio) {                     /* :443 */
            memcpy(bestT, T, sizeof(bestT));
            bestInlierRatio = inlierRatio;
            n_best = cnt;
            memcpy(best_inl, cur_inl, sizeof(int) * (size_t)cnt);
            if (g_stop_rule == 1) {   /* USAC standard stopping, capped by the hypothesis budget (USAC.h:326,447) */
                unsigned cap = (unsigned)(num_hyp > 0 ? num_hyp : 487);
                unsigned b = usac_standard_stopping((unsigned)cnt, (unsigned)mf, 3, cap);
                iterationCount = (int)(b < cap ? b : cap);
            } else if (num_hyp <= 0) {                                   /* :450-453 */
                int a = ransac_iterations(P->minimal_inlier_ratio_threshold);
                int b = ransac_iterations(bestInlierRatio);
                iterationCount = a < b ? a : b;
            }
        }
    }
    if (hyp_used_out) *hyp_used_out = i;
    /* :152-158 refit on the best inliers, recount (Euclidean) among them only */
    int *rq = (int *)malloc(sizeof(int) * (size_t)(n_best > 0 ? n_best : 1));
    int *rt = (int *)malloc(sizeof(int) * (size_t)(n_best > 0 ? n_best : 1));
    for (int k = 0; k < n_best; ++k) { rq[k] = fq[best_inl[k]]; rt[k] = ft[best_inl[k]]; }
    umeyama_idx(prev, rq, cur, rt, n_best, bestT);
    int n_final = 0;
    for (int k = 0; k < n_best; ++k)
        if (is_inlier(bestT, NULL, prev + 3 * rq[k], cur + 3 * rt[k], P, 1)) inl_idx[n_final++] = keep[best_inl[k]];
    if (bestInlierRatio < P->minimal_inlier_ratio_threshold) {           /* :161-164 */
        for (int a = 0; a < 16; ++a) bestT[a] = (a % 5 == 0) ? 1.f : 0.f;
        n_final = 0;
    }
    memcpy(T_out, bestT, sizeof(bestT));
    *n_inl_out = n_final;
    if (best_ratio_out) *best_ratio_out = bestInlierRatio;
    free(keep); free(fq); free(ft); free(best_inl); free(cur_inl); free(rq); free(rt);
    return 0;
}

/* All-cores variant of the fixed-H mode of orc_ransac (SURVEY 8d "OpenMP-over-hypotheses"): the H hypotheses are
 * independent, so models and inlier COUNTS are evaluated in parallel; the reference's selection (first strict maximum),
 * the refit and the recount then run once.  Same answer as orc_ransac(..., num_hyp = H) -- used only as the all-cores
 * CPU baseline of bench.py.  Returns 0. */
ORC_API int orc_ransac_fixed_mt(const float *prev, int n1, const float *cur, int n2, const int *mq, const int *mt,
                                int m, const orc_ransac_params *P, uint64_t seed, int num_hyp, int threads, float *T_out,
                                int *inl_idx, int *n_inl_out, double *best_ratio_out) {
    (void)n1; (void)n2;
    for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.f : 0.f;
    *n_inl_out = 0;
    if (best_ratio_out) *best_ratio_out = 0.0;
    int *keep = (int *)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    int mf = 0;
    for (int k = 0; k < m; ++k) {
        const float *p = prev + 3 * mq[k], *c = cur + 3 * mt[k];
        int bad = isnan(p[0]) || isnan(p[1]) || isnan(p[2]) || isnan(c[0]) || isnan(c[1]) || isnan(c[2]) ||
                  p[2] < 0.1 || p[2] > 6 || c[2] < 0.1 || c[2] > 6;
        if (!bad) keep[mf++] = k;
    }
    if (mf < P->minimal_number_of_matches || num_hyp <= 0) { free(keep); return 0; }
    int *fq = (int *)malloc(sizeof(int) * (size_t)mf), *ft = (int *)malloc(sizeof(int) * (size_t)mf);
    for (int k = 0; k < mf; ++k) { fq[k] = mq[keep[k]]; ft[k] = mt[keep[k]]; }
    int *counts = (int *)malloc(sizeof(int) * (size_t)num_hyp);
    float *models = (float *)malloc(sizeof(float) * 16 * (size_t)num_hyp);
    if (threads < 1) threads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
#endif
    for (int i = 0; i < num_hyp; ++i) {
        int s[3];
        orc_sample3(seed, (uint32_t)i, mf, s);
        int sq[3] = {fq[s[0]], fq[s[1]], fq[s[2]]}, st[3] = {ft[s[0]], ft[s[1]], ft[s[2]]};
        float *T = models + 16 * (size_t)i, Tinv[16];
        counts[i] = -1;
        if (!umeyama_idx(prev, sq, cur, st, 3, T)) continue;
        if (P->error_version == 1 || P->error_version == 2) inverse4(T, Tinv);
        int cnt = 0;
        for (int k = 0; k < mf; ++k) cnt += is_inlier(T, Tinv, prev + 3 * fq[k], cur + 3 * ft[k], P, 0);
        counts[i] = cnt;
    }
    int win = -1;
    double bestInlierRatio = 0.0;
    for (int i = 0; i < num_hyp; ++i) {
        if (counts[i] < 0) continue;
        float r = (float)counts[i] / (float)mf;
        if ((double)r > bestInlierRatio) { bestInlierRatio = r; win = i; }
    }
    float bestT[16];
    for (int i = 0; i < 16; ++i) bestT[i] = (i % 5 == 0) ? 1.f : 0.f;
    int *best_inl = (int *)malloc(sizeof(int) * (size_t)mf), n_best = 0;
    if (win >= 0) {
        float *T = models + 16 * (size_t)win, Tinv[16];
        if (P->error_version == 1 || P->error_version == 2) inverse4(T, Tinv);
        for (int k = 0; k < mf; ++k)
            if (is_inlier(T, Tinv, prev + 3 * fq[k], cur + 3 * ft[k], P, 0)) best_inl[n_best++] = k;
    }
    int *rq = (int *)malloc(sizeof(int) * (size_t)(n_best > 0 ? n_best : 1));
    int *rt = (int *)malloc(sizeof(int) * (size_t)(n_best > 0 ? n_best : 1));
    for (int k = 0; k < n_best; ++k) { rq[k] = fq[best_inl[k]]; rt[k] = ft[best_inl[k]]; }
    umeyama_idx(prev, rq, cur, rt, n_best, bestT);
    int n_final = 0;
    for (int k = 0; k < n_best; ++k)
        if (is_inlier(bestT, NULL, prev + 3 * rq[k], cur + 3 * rt[k], P, 1)) inl_idx[n_final++] = keep[best_inl[k]];
    if (bestInlierRatio < P->minimal_inlier_ratio_threshold) {
        for (int a = 0; a < 16; ++a) bestT[a] = (a % 5 == 0) ? 1.f : 0.f;
        n_final = 0;
    }
    memcpy(T_out, bestT, sizeof(bestT));
    *n_inl_out = n_final;
    if (best_ratio_out) *best_ratio_out = bestInlierRatio;
    free(keep); free(fq); free(ft); free(counts); free(models); free(best_inl); free(rq); free(rt);
    return 0;
}

/* RANSAC::pointInlierRatio (include/putslam/TransformEst/RANSAC.h:56-66):
 * |unique trainIdx of inliers| / |unique trainIdx of all (unfiltered) matches|. */
ORC_API double orc_point_inlier_ratio(const int *inl_t, int n_inl, const int *all_t, int n_all, int n_train) {
    uint8_t *seen = (uint8_t *)calloc((size_t)(n_train > 0 ? n_train : 1), 1);
    int a = 0, b = 0;
    for (int i = 0; i < n_all; ++i) if (!seen[all_t[i]]) { seen[all_t[i]] = 1; ++b; }
    memset(seen, 0, (size_t)(n_train > 0 ? n_train : 1));
    for (int i = 0; i < n_inl; ++i) if (!seen[inl_t[i]]) { seen[inl_t[i]] = 1; ++a; }
    free(seen);
    return (double)a / (double)b;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
