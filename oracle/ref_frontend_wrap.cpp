// C entry points around the REFERENCE's own front-end code, compiled from /root/reference where it lies by
// `make -C oracle ref` into oracle/_ref/libref_frontend*.so:
//     src/TransformEst/RANSAC.cpp   src/RGBD/RGBD.cpp   src/TransformEst/kabschEst.cpp   src/Grabber/depthSensorModel.cpp
//     src/Matcher/matcher.cpp   src/Matcher/MatchingOnPatches.cpp   src/Matcher/dbscan.cpp   3rdParty/tinyXML/tinyxml2.cpp
// against the Eigen / OpenCV stand-ins of oracle/ref_shim (what those stand-ins decide is listed in their headers).
// TEST INFRASTRUCTURE ONLY: tests/ load this library to check the CPU oracle (oracle.c) -- and through it the CUDA path --
// against the reference's compiled scalar code.  Nothing of the reference is copied here: this file only marshals plain
// arrays into the reference's types and calls its functions.
//
// rand() / srand() are defined in this library and the library is linked -Bsymbolic, so RANSAC::getRandomMatches
// (RANSAC.cpp:180-205) draws from the replayed Philox stream below instead of libc's generator: hypothesis h takes the
// 32-bit words of Philox4x32-10(ctr = {block, h, 0, 0}, key = seed) in order, each reduced modulo the number of matches
// -- the sample stream of the product (include/pslam_b200.h, pslam_ransac_sample) and of oracle.c (orc_sample3).
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <opencv2/shim_cv.h>
#include "../../3rdParty/tinyXML/tinyxml2.h"

// the seams we call are private / protected members of Matcher (matcher.h:380-447); access control does not change layout
#define private public
#define protected public
#include "Matcher/matcher.h"
#undef private
#undef protected
#include "Matcher/dbscan.h"
#include "TransformEst/kabschEst.h"
#include "Grabber/depthSensorModel.h"

#define REF_API extern "C" __attribute__((visibility("default")))

// ---- the replayed sample stream -------------------------------------------------------------------------------------
namespace {
struct Replay {
    bool active = false;
    uint64_t seed = 0;
    uint32_t m = 1, hyp = 0, blk = 0;
    int w = 4, got = 0, used_pairs = 3;
    uint32_t r[4];
    int chosen[8];
    uint32_t lcg = 12345u;
} g_replay;

void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {   // Salmon et al., SC'11
    uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3], k0 = key_in[0], k1 = key_in[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
}  // namespace

extern "C" __attribute__((visibility("default"))) void srand(unsigned) {}
extern "C" __attribute__((visibility("default"))) int rand(void) {
    Replay& g = g_replay;
    if (!g.active) { g.lcg = g.lcg * 1103515245u + 12345u; return (int)((g.lcg >> 1) & 0x7fffffff); }
    if (g.w == 4) {
        const uint32_t ctr[4] = {g.blk, g.hyp, 0u, 0u}, key[2] = {(uint32_t)g.seed, (uint32_t)(g.seed >> 32)};
        philox4x32_10(ctr, key, g.r);
        ++g.blk; g.w = 0;
    }
    const int idx = (int)(g.r[g.w++] % g.m);
    // follow getRandomMatches' rejection so that the next hypothesis starts a fresh counter block
    bool dup = false;
    for (int a = 0; a < g.got; ++a) if (g.chosen[a] == idx) dup = true;
    if (!dup) g.chosen[g.got++] = idx;
    if (g.got == g.used_pairs) { ++g.hyp; g.blk = 0; g.w = 4; g.got = 0; }
    return idx;                                   // idx % m == idx
}

namespace putslam { TransformEst* createG2OEstimator(void) { std::abort(); } }    // RANSAC.cpp:227-232, never selected

namespace {

struct CoutCapture {                               // the reference reports some intermediate values only on std::cout
    std::streambuf* old; std::ostringstream os;
    CoutCapture() : old(std::cout.rdbuf(os.rdbuf())) {}
    ~CoutCapture() { std::cout.rdbuf(old); }
    double after(const std::string& key, double dflt) const {
        const std::string s = os.str();
        const std::size_t p = s.rfind(key);
        if (p == std::string::npos) return dflt;
        return std::atof(s.c_str() + p + key.size());
    }
};

struct RansacArgs {
    int error_version; double thr_e, thr_r, min_ratio; int min_matches, used_pairs; float fx, fy, cx, cy;
};

RANSAC::parameters to_params(const RansacArgs& a, int verbose) {
    RANSAC::parameters p;
    std::memset(&p, 0, sizeof(p));
    p.verbose = verbose;
    p.errorVersion = p.errorVersionVO = p.errorVersionMap = a.error_version;
    p.inlierThresholdEuclidean = a.thr_e; p.inlierThresholdReprojection = a.thr_r; p.inlierThresholdMahalanobis = 0;
    p.minimalInlierRatioThreshold = a.min_ratio; p.minimalNumberOfMatches = a.min_matches; p.usedPairs = a.used_pairs;
    return p;
}
cv::Mat camera(const RansacArgs& a) {
    cv::Mat K = cv::Mat::zeros(3, 3, CV_32FC1);
    K.at<float>(0, 0) = a.fx; K.at<float>(1, 1) = a.fy; K.at<float>(0, 2) = a.cx; K.at<float>(1, 2) = a.cy; K.at<float>(2, 2) = 1.f;
    return K;
}
cv::Mat camera4(const float* k) { RansacArgs a{}; a.fx = k[0]; a.fy = k[1]; a.cx = k[2]; a.cy = k[3]; return camera(a); }
void out_T(const Eigen::Matrix4f& T, float* o) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o[4 * i + j] = T(i, j); }
std::vector<Eigen::Vector3f> points(const float* p, int n) {
    std::vector<Eigen::Vector3f> v((size_t)n);
    for (int i = 0; i < n; ++i) v[(size_t)i] = Eigen::Vector3f(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}

// Number of matches left by the filter of estimateTransformation (RANSAC.cpp:65-74), as the reference itself reports it
// with verbose = 1 ("RANSAC: matches.size() = N", :83-84); -1 when it returned early (:77-80).  A first, throw-away run.
int probe_filtered_size(const RansacArgs& a, const cv::Mat& K, const std::vector<Eigen::Vector3f>& prev,
                        const std::vector<Eigen::Vector3f>& cur, const std::vector<cv::DMatch>& matches) {
    CoutCapture cap;
    g_replay.active = false;
    RANSAC r(to_params(a, 1), K);
    std::vector<cv::DMatch> inl;
    r.estimateTransformation(prev, cur, matches, inl);
    return (int)cap.after("RANSAC: matches.size() = ", -1.0);
}

void start_replay(uint64_t seed, int m, int used_pairs) {
    g_replay.active = true; g_replay.seed = seed; g_replay.m = (uint32_t)(m > 0 ? m : 1); g_replay.hyp = 0; g_replay.blk = 0;
    g_replay.w = 4; g_replay.got = 0; g_replay.used_pairs = used_pairs;
}

}  // namespace

// ---- stage 3: RANSAC::estimateTransformation --------------------------------------------------------------------------
// inl_idx: indices into the caller's match list (carried through DMatch::imgIdx).  best_ratio_pct: the value the
// reference prints (bestInlierRatio * 100, 6 significant digits; -1 if not printed).  Returns the filtered match count.
REF_API int ref_ransac(const float* prev, int n1, const float* cur, int n2, const int* mq, const int* mt, int m,
                       const RansacArgs* a, uint64_t seed, float* T_out, int* inl_idx, int* n_inl, double* best_ratio_pct,
                       int* hyp_used) {
    const std::vector<Eigen::Vector3f> P = points(prev, n1), C = points(cur, n2);
    std::vector<cv::DMatch> matches((size_t)m);
    for (int k = 0; k < m; ++k) matches[(size_t)k] = cv::DMatch(mq[k], mt[k], k, 0.f);
    const cv::Mat K = camera(*a);
    const int mf = probe_filtered_size(*a, K, P, C, matches);
    CoutCapture cap;
    start_replay(seed, mf, a->used_pairs);
    RANSAC r(to_params(*a, 1), K);
    std::vector<cv::DMatch> inl;
    const Eigen::Matrix4f T = r.estimateTransformation(P, C, matches, inl);
    g_replay.active = false;
    out_T(T, T_out);
    *n_inl = (int)inl.size();
    for (size_t k = 0; k < inl.size(); ++k) inl_idx[k] = inl[k].imgIdx;
    if (best_ratio_pct) *best_ratio_pct = cap.after("RANSAC best model : inlierRatio = ", -1.0);
    if (hyp_used) *hyp_used = mf < 0 ? 0 : (int)g_replay.hyp;
    return mf;
}

REF_API double ref_point_inlier_ratio(const int* inl_t, int n_inl, const int* all_t, int n_all) {
    std::vector<cv::DMatch> a((size_t)n_inl), b((size_t)n_all);
    for (int i = 0; i < n_inl; ++i) a[(size_t)i].trainIdx = inl_t[i];
    for (int i = 0; i < n_all; ++i) b[(size_t)i].trainIdx = all_t[i];
    return RANSAC::pointInlierRatio(a, b);
}

// ---- stage 1: RGBD back-projection, projection, sensor model ------------------------------------------------------------
REF_API void ref_backproject(const float* uv, int n, const uint16_t* depth, int W, int H, const float* k4, const float* dist5,
                             double scale, float* uv_used, float* xyz, double* det_dist) {
    std::vector<cv::Point2f> pts((size_t)n);
    for (int i = 0; i < n; ++i) pts[(size_t)i] = cv::Point2f(uv[2 * i], uv[2 * i + 1]);
    const cv::Mat K = camera4(k4);
    if (dist5) {
        cv::Mat D = cv::Mat::zeros(1, 5, CV_32FC1);
        for (int i = 0; i < 5; ++i) D.at<float>(i) = dist5[i];
        pts = RGBD::removeImageDistortion(pts, K, D);                                   // RGBD.cpp:286-314
    }
    const cv::Mat depthImage(H, W, CV_16UC1, (void*)depth);
    const std::vector<Eigen::Vector3f> p = RGBD::keypoints2Dto3D(pts, depthImage, K, scale);   // RGBD.cpp:30-65
    for (int i = 0; i < n; ++i) {
        if (uv_used) { uv_used[2 * i] = pts[(size_t)i].x; uv_used[2 * i + 1] = pts[(size_t)i].y; }
        for (int c = 0; c < 3; ++c) xyz[3 * i + c] = p[(size_t)i][c];
        if (det_dist) {   // the detDist loop is inline in Matcher (matcher.cpp:51-58); same expression on the same floats
            det_dist[i] = std::sqrt(p[(size_t)i][0] * p[(size_t)i][0] + p[(size_t)i][1] * p[(size_t)i][1] + p[(size_t)i][2] * p[(size_t)i][2]);
        }
    }
}
REF_API int ref_round_size(double x, int size) { return RGBD::roundSize(x, size); }
REF_API void ref_point3Dto2D(const float* xyz, int n, const float* k4, float* uv) {
    const cv::Mat K = camera4(k4);
    for (int i = 0; i < n; ++i) {
        const cv::Point2f p = RGBD::point3Dto2D(Eigen::Vector3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), K);   // RGBD.cpp:92-98
        uv[2 * i] = p.x; uv[2 * i + 1] = p.y;
    }
}
namespace {
struct SensorArgs { double fx, fy, cx, cy, varU, varV, c[4]; int img_w, img_h; double scale_normal, scale_gradient; };
DepthSensorModel sensor(const SensorArgs& s) {
    DepthSensorModel m;                     // default Config, then every field the functions read (depthSensorModel.h:59-91)
    m.config.focalLength[0] = s.fx; m.config.focalLength[1] = s.fy; m.config.focalAxis[0] = s.cx; m.config.focalAxis[1] = s.cy;
    m.config.varU = s.varU; m.config.varV = s.varV;
    for (int i = 0; i < 4; ++i) m.config.distVarCoefs[i] = s.c[i];
    m.config.imageSize[0] = s.img_w; m.config.imageSize[1] = s.img_h;
    m.config.scaleUncertaintyNormal = s.scale_normal; m.config.scaleUncertaintyGradient = s.scale_gradient;
    m.Ruvd.setZero(); m.Ruvd(0, 0) = s.varU; m.Ruvd(1, 1) = s.varV;      // what the file constructor sets (depthSensorModel.cpp:8-10)
    return m;
}
void out33(const putslam::Mat33& M, double* o) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o[3 * i + j] = M(i, j); }
}  // namespace
REF_API void ref_compute_cov(unsigned u, unsigned v, double depth, const SensorArgs* s, double* cov) {
    DepthSensorModel m = sensor(*s);
    putslam::Mat33 c; m.computeCov((uint_fast16_t)u, (uint_fast16_t)v, depth, c);          // depthSensorModel.cpp:28-36
    out33(c, cov);
}
REF_API void ref_information_matrix_uvz(double u, double v, double z, const SensorArgs* s, double* info) {
    DepthSensorModel m = sensor(*s);
    out33(m.informationMatrixFromImageCoordinates(u, v, z), info);                         // :55-59
}
REF_API void ref_information_matrix_xyz(double x, double y, double z, const SensorArgs* s, double* info) {
    DepthSensorModel m = sensor(*s);
    out33(m.informationMatrix(x, y, z), info);                                             // :48-53
}
REF_API void ref_inverse_model(double x, double y, double z, const SensorArgs* s, double* uvz) {
    DepthSensorModel m = sensor(*s);
    const Eigen::Vector3d p = m.inverseModel(x, y, z);                                     // :18-25
    uvz[0] = p(0); uvz[1] = p(1); uvz[2] = p(2);
}
REF_API void ref_uncertainty_from_normal(const double* n, const SensorArgs* s, double* cov) {
    DepthSensorModel m = sensor(*s);
    out33(m.uncertinatyFromNormal(putslam::Vec3(n[0], n[1], n[2])), cov);                  // :62-76
}
REF_API void ref_uncertainty_from_gradient(const double* g, const SensorArgs* s, double* cov) {
    DepthSensorModel m = sensor(*s);
    out33(m.uncertinatyFromRGBGradient(putslam::Vec3(g[0], g[1], g[2])), cov);             // :79-95
}
REF_API void ref_compute_normal(const uint16_t* depth, int W, int H, int u, int v, const float* k4, double scale, double* out) {
    const cv::Mat depthImage(H, W, CV_16UC1, (void*)depth);
    const putslam::Vec3 n = RGBD::computeNormal(depthImage, u, v, camera4(k4), scale);     // RGBD.cpp:101-144
    out[0] = n.x(); out[1] = n.y(); out[2] = n.z();
}
REF_API void ref_compute_rgb_gradient(const uint8_t* rgb, int rgb_step, const uint16_t* depth, int W, int H, int u, int v,
                                      const float* k4, double scale, double* out) {
    const cv::Mat rgbImage(H, W, CV_8UC3, (void*)rgb, (size_t)rgb_step);
    const cv::Mat depthImage(H, W, CV_16UC1, (void*)depth);
    const putslam::Vec3 g = RGBD::computeRGBGradient(rgbImage, depthImage, u, v, camera4(k4), scale);   // RGBD.cpp:147-187
    out[0] = g.x(); out[1] = g.y(); out[2] = g.z();
}

// ---- TransformEst: Kabsch and the transform covariance ----------------------------------------------------------------
namespace {
Eigen::MatrixXd rows3(const double* p, int n) {
    Eigen::MatrixXd m(n, 3);
    for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) m(i, c) = p[3 * i + c];
    return m;
}
}  // namespace
REF_API void ref_kabsch(const double* A, const double* B, int n, double* T12) {
    putslam::KabschEst est;
    const putslam::Mat34& T = est.computeTransformation(rows3(A, n), rows3(B, n));        // kabschEst.cpp:24-68
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) T12[4 * i + j] = T.matrix()(i, j);
}
// mode 0: TransformEst::computeUncertainty (transformEst.h:29-144), 1: computeUncertaintyG2O (:147-272); T12 row-major 3x4
REF_API void ref_transform_uncertainty(const double* A, const double* B, const double* CA, const double* CB, int n,
                                       const double* T12, int mode, double* U36) {
    putslam::KabschEst est;
    std::vector<putslam::Mat33> ca((size_t)n), cb((size_t)n);
    for (int k = 0; k < n; ++k) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        ca[(size_t)k](i, j) = CA[9 * k + 3 * i + j]; cb[(size_t)k](i, j) = CB[9 * k + 3 * i + j];
    }
    putslam::Mat34 T; T.setIdentity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) T.matrix()(i, j) = T12[4 * i + j];
    const Eigen::MatrixXd a = rows3(A, n), b = rows3(B, n);
    const putslam::Mat66& U = mode == 0 ? est.computeUncertainty(a, ca, b, cb, T) : est.computeUncertaintyG2O(a, ca, b, cb, T);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) U36[6 * i + j] = U(i, j);
}

// ---- stage 2: the Matcher facade with scripted virtuals -----------------------------------------------------------------
namespace {
class ScriptedMatcher : public putslam::Matcher {
public:
    ScriptedMatcher() : Matcher("scripted") {}
    const std::string& getName() const { return name; }
    std::vector<cv::KeyPoint> scriptKeyPoints;      // what detectFeatures "finds" (class_id = row of scriptDescriptors)
    cv::Mat scriptDescriptors;
    std::vector<uint8_t> scriptDescribable;         // describeFeatures drops key points whose flag is 0 (ORB's border rule)
    // performTracking script
    std::vector<cv::Point2f> trkFeatures; std::vector<cv::KeyPoint> trkKeyPoints; std::vector<double> trkDetDists;
    std::vector<cv::DMatch> trkMatches;

    std::vector<cv::KeyPoint> detectFeatures(cv::Mat) { return scriptKeyPoints; }
    cv::Mat describeFeatures(cv::Mat, std::vector<cv::KeyPoint>& features) {
        cv::Mat d;
        std::vector<cv::KeyPoint> kept;
        for (const cv::KeyPoint& k : features) {
            if (!scriptDescribable.empty() && !scriptDescribable[(size_t)k.class_id]) continue;
            kept.push_back(k);
            d.push_back(scriptDescriptors.row(k.class_id));
        }
        features.swap(kept);
        return d;
    }
    // cv::BFMatcher(NORM_HAMMING, crossCheck = true).match(query = prev, train = cur)  (matcherOpenCV.cpp:97-106,198-206);
    // semantics pinned against cv2 (tests/golden/bf_cv2.npz): mutual first-argmin, ascending queryIdx, imgIdx 0
    std::vector<cv::DMatch> performMatching(cv::Mat q, cv::Mat t) {
        std::vector<cv::DMatch> out;
        if (q.rows == 0 || t.rows == 0) return out;
        const int nb = q.cols;
        std::vector<int> fwd((size_t)q.rows), fd((size_t)q.rows), bwd((size_t)t.rows, -1), bd((size_t)t.rows, 1 << 30);
        for (int i = 0; i < q.rows; ++i) {
            int best = 1 << 30, bi = -1;
            for (int j = 0; j < t.rows; ++j) {
                int d = 0;
                for (int b = 0; b < nb; ++b) d += __builtin_popcount((unsigned)(q.at<uchar>(i, b) ^ t.at<uchar>(j, b)));
                if (d < best) { best = d; bi = j; }
                if (d < bd[(size_t)j]) { bd[(size_t)j] = d; bwd[(size_t)j] = i; }
            }
            fwd[(size_t)i] = bi; fd[(size_t)i] = best;
        }
        for (int i = 0; i < q.rows; ++i)
            if (bwd[(size_t)fwd[(size_t)i]] == i) out.push_back(cv::DMatch(i, fwd[(size_t)i], 0, (float)fd[(size_t)i]));
        return out;
    }
    std::vector<cv::DMatch> performTracking(cv::Mat, cv::Mat, std::vector<cv::Point2f>&, std::vector<cv::Point2f>& features,
                                            std::vector<cv::KeyPoint>&, std::vector<cv::KeyPoint>& keyPoints, std::vector<double>&,
                                            std::vector<double>& detDists) {
        features = trkFeatures; keyPoints = trkKeyPoints; detDists = trkDetDists;
        return trkMatches;
    }
};

struct MatcherArgs {
    RansacArgs ransac;
    float dist[5];
    double radius, accept_ratio, dbscan_eps;
    double min_reproj_dist, min_euclid_dist;
    int remove_too_close, minimal_tracked_features;
    int ransac_verbose;
};
void configure(ScriptedMatcher& M, const MatcherArgs& a) {
    putslam::Matcher::MatcherParameters& p = M.matcherParameters;
    p.verbose = 0; p.VOVersion = 0; p.maxAngleBetweenFrames = 0;
    p.RANSACParams = to_params(a.ransac, a.ransac_verbose);
    p.OpenCVParams.detector = "ORB"; p.OpenCVParams.descriptor = "ORB";
    p.OpenCVParams.matchingXYZSphereRadius = a.radius; p.OpenCVParams.matchingXYZacceptRatioOfBestMatch = a.accept_ratio;
    p.OpenCVParams.DBScanEps = a.dbscan_eps;
    p.OpenCVParams.minimalReprojDistanceNewTrackingFeatures = a.min_reproj_dist;
    p.OpenCVParams.minimalEuclidDistanceNewTrackingFeatures = a.min_euclid_dist;
    p.OpenCVParams.removeTooCloseFeatures = a.remove_too_close;
    p.OpenCVParams.minimalTrackedFeatures = a.minimal_tracked_features;
    p.cameraMatrixMat = camera(a.ransac);
    p.distortionCoeffsMat = cv::Mat::zeros(1, 5, CV_32FC1);
    for (int i = 0; i < 5; ++i) p.distortionCoeffsMat.at<float>(i) = a.dist[i];
}
cv::Mat desc_rows(const uint8_t* d, int n) {
    cv::Mat m(n, 32, CV_8UC1);
    if (n) std::memcpy(m.data, d, (size_t)n * 32);
    return m;
}
}  // namespace

// Matcher::matchXYZ (public overload, newDetection = false -> the private one, matcher.cpp:542-798): guided matching of
// M map features against the N key points of the "previous" frame state, then RANSAC.
//   map_xyz  M x 3 double, map_desc M x 32, map_octave / map_detdist per map feature (its single ExtendedDescriptor)
//   cur_xyz  N x 3 float,  cur_desc N x 32, cur_octave, cur_detdist
// Outputs: T (row-major 4x4), inlier (map index, current index) pairs in order, the value matchXYZ returns, and the
// "MatchesXYZ - we found : n (Perfect matches = p)" line (verbose > 0, matcher.cpp:750-753).  Returns the pair count.
REF_API int ref_match_xyz(const double* map_xyz, const uint8_t* map_desc, const int* map_octave, const double* map_detdist, int M,
                          const float* cur_xyz, const uint8_t* cur_desc, const int* cur_octave, const double* cur_detdist, int N,
                          const MatcherArgs* a, int computation_number, uint64_t seed, int use_frame_ids, float* T_out,
                          int* pair_map, int* pair_cur, double* ratio_out, int* n_matches, int* n_perfect, int* hyp_used) {
    ScriptedMatcher Mx;
    configure(Mx, *a);
    Mx.matcherParameters.verbose = 1;
    std::vector<putslam::MapFeature> map((size_t)M);
    std::vector<int> frameIds;
    for (int j = 0; j < M; ++j) {
        putslam::MapFeature& f = map[(size_t)j];
        f.id = (unsigned)j;
        f.position = putslam::Vec3(map_xyz[3 * j], map_xyz[3 * j + 1], map_xyz[3 * j + 2]);
        putslam::ExtendedDescriptor e(cv::Point2f(0, 0), cv::Point2f(0, 0), f.position, desc_rows(map_desc + 32 * (size_t)j, 1),
                                      map_octave[j], map_detdist[j]);
        const unsigned pose = use_frame_ids ? (unsigned)(7 + j % 3) : 0u;
        f.descriptors[pose] = e;
        if (use_frame_ids) frameIds.push_back((int)pose);
    }
    Mx.prevDescriptors = desc_rows(cur_desc, N);
    Mx.prevFeatures3D = points(cur_xyz, N);
    Mx.prevKeyPoints.resize((size_t)N); Mx.prevDetDists.resize((size_t)N);
    Mx.prevFeaturesUndistorted.resize((size_t)N); Mx.prevFeaturesDistorted.resize((size_t)N);
    for (int i = 0; i < N; ++i) {
        Mx.prevKeyPoints[(size_t)i].octave = cur_octave[i];
        Mx.prevDetDists[(size_t)i] = cur_detdist[i];
        Mx.prevFeaturesUndistorted[(size_t)i] = cv::Point2f((float)i, 0.f);      // u carries the current index to the output
        Mx.prevFeaturesDistorted[(size_t)i] = cv::Point2f((float)i, 0.f);
    }
    // the filtered size RANSAC will see is not known before the guided matching has run: first pass with a throw-away stream
    int mf = -1;
    {
        CoutCapture cap;
        g_replay.active = false;
        Mx.matcherParameters.RANSACParams.verbose = 1;
        std::vector<putslam::MapFeature> found; Eigen::Matrix4f T;
        Mx.matchXYZ(map, 99, found, T, false, frameIds, computation_number);
        mf = (int)cap.after("RANSAC: matches.size() = ", -1.0);
    }
    CoutCapture cap;
    start_replay(seed, mf, a->ransac.used_pairs);
    Mx.matcherParameters.RANSACParams.verbose = 1;
    std::vector<putslam::MapFeature> found;
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const double ratio = Mx.matchXYZ(map, 99, found, T, false, frameIds, computation_number);
    g_replay.active = false;
    out_T(T, T_out);
    for (size_t k = 0; k < found.size(); ++k) { pair_map[k] = (int)found[k].id; pair_cur[k] = (int)found[k].u; }
    *ratio_out = ratio;
    *n_matches = (int)cap.after("MatchesXYZ - we found : ", -1.0);
    *n_perfect = (int)cap.after("(Perfect matches = ", -1.0);
    if (hyp_used) *hyp_used = mf < 0 ? 0 : (int)g_replay.hyp;
    return (int)found.size();
}

// Matcher::match (matcher.cpp:452-516): one VO step.  The previous frame state is (prev_desc, prev_xyz); the scripted
// detector returns cur_kp (x, y, octave) with descriptors cur_desc; DBScan, undistortion, back-projection, brute-force
// matching, RANSAC and pointInlierRatio are the reference's.  Outputs: T, inlier (queryIdx, trainIdx) pairs, the surviving
// key-point rows (class ids), their 3-D points and undistorted pixels, the return value.  Returns the inlier count.
REF_API int ref_match_vo(const uint8_t* prev_desc, const float* prev_xyz, int n_prev, const float* cur_kp_xy, const int* cur_octave,
                         const uint8_t* cur_desc, int n_cur, const uint16_t* depth, int W, int H, double depth_scale,
                         const MatcherArgs* a, uint64_t seed, float* T_out, int* inl_q, int* inl_t, double* ratio_out,
                         int* kept_ids, float* kept_xyz, float* kept_uv, int* n_kept, int* hyp_used) {
    auto prepare = [&](ScriptedMatcher& Mx) {
        configure(Mx, *a);
        Mx.prevDescriptors = desc_rows(prev_desc, n_prev);
        Mx.prevFeatures3D = points(prev_xyz, n_prev);
        Mx.prevKeyPoints.resize((size_t)n_prev);
        Mx.scriptDescriptors = desc_rows(cur_desc, n_cur);
        Mx.scriptKeyPoints.resize((size_t)n_cur);
        for (int i = 0; i < n_cur; ++i) {
            cv::KeyPoint& k = Mx.scriptKeyPoints[(size_t)i];
            k.pt = cv::Point2f(cur_kp_xy[2 * i], cur_kp_xy[2 * i + 1]); k.octave = cur_octave[i]; k.class_id = i;
        }
    };
    putslam::SensorFrame frame;
    frame.depthImage = cv::Mat(H, W, CV_16UC1, (void*)depth);
    frame.depthImageScale = depth_scale;
    int mf = -1;
    {
        ScriptedMatcher Mx; prepare(Mx);
        Mx.matcherParameters.RANSACParams.verbose = 1;
        CoutCapture cap;
        g_replay.active = false;
        Eigen::Matrix4f T; std::vector<cv::DMatch> inl;
        Mx.match(frame, T, inl);
        mf = (int)cap.after("RANSAC: matches.size() = ", -1.0);
    }
    ScriptedMatcher Mx; prepare(Mx);
    CoutCapture cap;
    start_replay(seed, mf, a->ransac.used_pairs);
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity(); std::vector<cv::DMatch> inl;
    const double ratio = Mx.match(frame, T, inl);
    g_replay.active = false;
    out_T(T, T_out);
    for (size_t k = 0; k < inl.size(); ++k) { inl_q[k] = inl[k].queryIdx; inl_t[k] = inl[k].trainIdx; }
    *ratio_out = ratio;
    *n_kept = (int)Mx.prevKeyPoints.size();
    for (size_t k = 0; k < Mx.prevKeyPoints.size(); ++k) {
        kept_ids[k] = Mx.prevKeyPoints[k].class_id;
        for (int c = 0; c < 3; ++c) kept_xyz[3 * k + c] = Mx.prevFeatures3D[k][c];
        kept_uv[2 * k] = Mx.prevFeaturesUndistorted[k].x; kept_uv[2 * k + 1] = Mx.prevFeaturesUndistorted[k].y;
    }
    if (hyp_used) *hyp_used = mf < 0 ? 0 : (int)g_replay.hyp;
    return (int)inl.size();
}

// Matcher::matchFeatureLoopClosure (matcher.cpp:802-861) on two feature sets given as (desc, xyz double) per feature.
REF_API int ref_loop_closure(const uint8_t* desc0, const double* xyz0, int n0, const uint8_t* desc1, const double* xyz1, int n1,
                             const MatcherArgs* a, uint64_t seed, float* T_out, int* pair0, int* pair1, double* ret_out, int* hyp_used) {
    auto build = [&](std::vector<putslam::MapFeature> sets[2]) {
        const uint8_t* d[2] = {desc0, desc1}; const double* x[2] = {xyz0, xyz1}; const int n[2] = {n0, n1};
        for (int s = 0; s < 2; ++s) {
            sets[s].resize((size_t)n[s]);
            for (int j = 0; j < n[s]; ++j) {
                putslam::MapFeature& f = sets[s][(size_t)j];
                f.id = (unsigned)j;
                const putslam::Vec3 p(x[s][3 * j], x[s][3 * j + 1], x[s][3 * j + 2]);
                f.descriptors[(unsigned)(3 + s)] = putslam::ExtendedDescriptor(cv::Point2f(0, 0), cv::Point2f((float)j, 0.f), p,
                                                                               desc_rows(d[s] + 32 * (size_t)j, 1), 0, 0.0);
            }
        }
    };
    int frames[2] = {3, 4};
    int mf = -1;
    {
        ScriptedMatcher Mx; configure(Mx, *a);
        Mx.matcherParameters.RANSACParams.verbose = 1;
        std::vector<putslam::MapFeature> sets[2]; build(sets);
        CoutCapture cap;
        g_replay.active = false;
        std::vector<std::pair<int, int> > pairs; Eigen::Matrix4f T;
        Mx.matchFeatureLoopClosure(sets, frames, pairs, T);
        mf = (int)cap.after("RANSAC: matches.size() = ", -1.0);
    }
    ScriptedMatcher Mx; configure(Mx, *a);
    std::vector<putslam::MapFeature> sets[2]; build(sets);
    CoutCapture cap;
    start_replay(seed, mf, a->ransac.used_pairs);
    std::vector<std::pair<int, int> > pairs; Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const double ret = Mx.matchFeatureLoopClosure(sets, frames, pairs, T);
    g_replay.active = false;
    out_T(T, T_out);
    for (size_t k = 0; k < pairs.size(); ++k) { pair0[k] = pairs[k].first; pair1[k] = pairs[k].second; }
    *ret_out = ret;
    if (hyp_used) *hyp_used = mf < 0 ? 0 : (int)g_replay.hyp;
    return (int)pairs.size();
}

// Matcher::removeTooCloseFeatures (matcher.cpp:886-974) and mergeTrackedFeatures (:95-130): the list edits of trackKLT.
// kept: indices of the surviving features; matches filtered in place (mq, mt, *m).  Returns the survivor count.
REF_API int ref_remove_too_close(const float* dist_xy, const float* undist_xy, const float* xyz, int n, int* mq, int* mt, int* m,
                                 double min_euclid, double min_reproj, int* kept) {
    ScriptedMatcher Mx;
    MatcherArgs a; std::memset(&a, 0, sizeof(a)); a.ransac.used_pairs = 3;
    configure(Mx, a);
    Mx.matcherParameters.OpenCVParams.minimalEuclidDistanceNewTrackingFeatures = min_euclid;
    Mx.matcherParameters.OpenCVParams.minimalReprojDistanceNewTrackingFeatures = min_reproj;
    std::vector<cv::Point2f> d((size_t)n), u((size_t)n); std::vector<cv::KeyPoint> kp((size_t)n); std::vector<double> dd((size_t)n);
    for (int i = 0; i < n; ++i) {
        d[(size_t)i] = cv::Point2f(dist_xy[2 * i], dist_xy[2 * i + 1]); u[(size_t)i] = cv::Point2f(undist_xy[2 * i], undist_xy[2 * i + 1]);
        kp[(size_t)i].class_id = i; dd[(size_t)i] = (double)i;
    }
    std::vector<Eigen::Vector3f> p = points(xyz, n);
    std::vector<cv::DMatch> matches((size_t)*m);
    for (int k = 0; k < *m; ++k) matches[(size_t)k] = cv::DMatch(mq[k], mt[k], 0, 0.f);
    Mx.removeTooCloseFeatures(d, u, p, kp, dd, matches);
    for (size_t k = 0; k < kp.size(); ++k) kept[k] = kp[k].class_id;
    for (size_t k = 0; k < matches.size(); ++k) { mq[k] = matches[k].queryIdx; mt[k] = matches[k].trainIdx; }
    *m = (int)matches.size();
    return (int)kp.size();
}
// added: indices (into the sandbox lists) of the features mergeTrackedFeatures appends, in order.  Returns their number.
REF_API int ref_merge_tracked(const float* undist_xy, int n, const float* sandbox_undist_xy, int ns, double min_reproj, int* added) {
    ScriptedMatcher Mx;
    MatcherArgs a; std::memset(&a, 0, sizeof(a)); a.ransac.used_pairs = 3;
    configure(Mx, a);
    Mx.matcherParameters.OpenCVParams.minimalReprojDistanceNewTrackingFeatures = min_reproj;
    std::vector<cv::Point2f> u((size_t)n), d((size_t)n), su((size_t)ns), sd((size_t)ns);
    std::vector<Eigen::Vector3f> p((size_t)n), sp((size_t)ns);
    std::vector<cv::KeyPoint> kp((size_t)n), skp((size_t)ns);
    std::vector<double> dd((size_t)n), sdd((size_t)ns);
    for (int i = 0; i < n; ++i) { u[(size_t)i] = cv::Point2f(undist_xy[2 * i], undist_xy[2 * i + 1]); kp[(size_t)i].class_id = -1; }
    for (int i = 0; i < ns; ++i) { su[(size_t)i] = cv::Point2f(sandbox_undist_xy[2 * i], sandbox_undist_xy[2 * i + 1]); skp[(size_t)i].class_id = i; }
    Mx.mergeTrackedFeatures(u, su, d, sd, p, sp, kp, skp, dd, sdd);
    int k = 0;
    for (size_t i = (size_t)n; i < kp.size(); ++i) added[k++] = kp[i].class_id;
    return k;
}

REF_API const char* ref_shim_model(void) {
    static char buf[256];
    std::snprintf(buf, sizeof(buf), "fixed_redux_tree=%d product_coeff_seq=%d umeyama_scale_lhs=%d jacobi_threshold32=%d jacobi_sweep_alt=%d umeyama_f64=%d",
                  SHIM_FIXED_REDUX_TREE, SHIM_PRODUCT_COEFF_SEQ, SHIM_UMEYAMA_SCALE_LHS, SHIM_JACOBI_THRESHOLD32, SHIM_JACOBI_SWEEP_ALT,
                  SHIM_UMEYAMA_F64);
    return buf;
}
