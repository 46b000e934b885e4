// C entry points that drive the REFERENCE's own Matcher (src/Matcher/matcher.cpp etc. compiled from /root/reference where
// they lie) with the B200 subclass of adapter/putslam_tree (MatcherB200 : public MatcherOpenCV) plugged into its virtual
// interface -- `make -C oracle ref` builds oracle/_ref/libref_tree.so from
//     reference : matcher.cpp  MatchingOnPatches.cpp  dbscan.cpp  RANSAC.cpp  RGBD.cpp  depthSensorModel.cpp  tinyxml2.cpp
//     ours      : adapter/putslam_tree/src/matcherB200.cpp  adapter/pslam_adapter.cpp (-DPSLAM_USE_REAL_HEADERS)
//     this file : a MatcherOpenCV whose constructors skip the XML files and whose own detect / describe / match / track
//                 abort (src/Matcher/matcherOpenCV.cpp needs OpenCV's features2d / xfeatures2d libraries, absent here;
//                 the point of the test is that the B200 overrides are what gets called), and the wrapper below.
// TEST INFRASTRUCTURE ONLY (tests/test_gpu_tree_integration.py): the reference's orchestration -- Matcher::detectInitFeatures,
// runVO -> match / trackKLT, DBScan, RGBD::*, RANSAC -- runs as compiled from its sources and reaches the GPU only through
// the four virtuals, exactly as a PUTSLAM build with Matcher/matcherB200.h would.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <opencv2/shim_cv.h>
#include "Matcher/matcherB200.h"
#include "TransformEst/transformEst.h"

#define TREE_API extern "C" __attribute__((visibility("default")))

// ---- MatcherOpenCV without OpenCV's feature libraries ------------------------------------------------------------------
MatcherOpenCV::MatcherOpenCV(void) : Matcher("OpenCV Matcher") {}
MatcherOpenCV::MatcherOpenCV(const std::string, const std::string) : Matcher("OpenCVMatcher") {}   // parameters are set by the wrapper
MatcherOpenCV::~MatcherOpenCV(void) {}
const std::string& MatcherOpenCV::getName() const { return name; }
void MatcherOpenCV::initVariables() {}
static void not_here(const char* what) { std::fprintf(stderr, "MatcherOpenCV::%s reached: the B200 override was bypassed\n", what); std::abort(); }
std::vector<cv::KeyPoint> MatcherOpenCV::detectFeatures(cv::Mat) { not_here("detectFeatures"); return {}; }
cv::Mat MatcherOpenCV::describeFeatures(cv::Mat, std::vector<cv::KeyPoint>&) { not_here("describeFeatures"); return cv::Mat(); }
std::vector<cv::DMatch> MatcherOpenCV::performMatching(cv::Mat, cv::Mat) { not_here("performMatching"); return {}; }
std::vector<cv::DMatch> MatcherOpenCV::performTracking(cv::Mat, cv::Mat, std::vector<cv::Point2f>&, std::vector<cv::Point2f>&,
                                                       std::vector<cv::KeyPoint>&, std::vector<cv::KeyPoint>&, std::vector<double>&,
                                                       std::vector<double>&) { not_here("performTracking"); return {}; }
namespace putslam { TransformEst* createG2OEstimator(void) { std::abort(); } }

// ---- the sample stream of RANSAC::getRandomMatches (see ref_frontend_wrap.cpp) -----------------------------------------
namespace {
struct Replay { bool active = false; uint64_t seed = 0; uint32_t m = 1, hyp = 0, blk = 0; int w = 4, got = 0; uint32_t r[4]; int chosen[4]; uint32_t lcg = 1u; } g;
void philox(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
    uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3], k0 = k_in[0], k1 = k_in[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
struct CoutCapture {
    std::streambuf* old; std::ostringstream os;
    CoutCapture() : old(std::cout.rdbuf(os.rdbuf())) {}
    ~CoutCapture() { std::cout.rdbuf(old); }
    double after(const std::string& key, double dflt) const {
        const std::string s = os.str(); const std::size_t p = s.rfind(key);
        return p == std::string::npos ? dflt : std::atof(s.c_str() + p + key.size());
    }
};
}  // namespace
extern "C" __attribute__((visibility("default"))) void srand(unsigned) {}
extern "C" __attribute__((visibility("default"))) int rand(void) {
    if (!g.active) { g.lcg = g.lcg * 1103515245u + 12345u; return (int)((g.lcg >> 1) & 0x7fffffff); }
    if (g.w == 4) { const uint32_t c[4] = {g.blk, g.hyp, 0u, 0u}, k[2] = {(uint32_t)g.seed, (uint32_t)(g.seed >> 32)}; philox(c, k, g.r); ++g.blk; g.w = 0; }
    const int idx = (int)(g.r[g.w++] % g.m);
    bool dup = false;
    for (int a = 0; a < g.got; ++a) if (g.chosen[a] == idx) dup = true;
    if (!dup) g.chosen[g.got++] = idx;
    if (g.got == 3) { ++g.hyp; g.blk = 0; g.w = 4; g.got = 0; }
    return idx;
}

struct TreeArgs {
    int vo_tracking;                       // Matcher::MatcherParameters::VOVERSION
    int error_version; double thr_e, thr_r, min_ratio; int min_matches;
    float fx, fy, cx, cy, dist[5];
    int grid_cols, grid_rows, maximal_tracked_features, minimal_tracked_features;
    double dbscan_eps;
    int win_size, max_levels, max_iter; float eps;
    double tracking_error_threshold, tracking_min_eig_threshold, min_reproj_dist, min_euclid_dist;
    int remove_too_close;
};

static void configure(MatcherB200& M, const TreeArgs& a, int ransac_verbose) {
    putslam::Matcher::MatcherParameters& p = M.matcherParameters;
    p.verbose = 0; p.VOVersion = a.vo_tracking; p.maxAngleBetweenFrames = 0;
    std::memset(&p.RANSACParams, 0, sizeof(p.RANSACParams));
    p.RANSACParams.verbose = ransac_verbose;
    p.RANSACParams.errorVersion = p.RANSACParams.errorVersionVO = p.RANSACParams.errorVersionMap = a.error_version;
    p.RANSACParams.inlierThresholdEuclidean = a.thr_e; p.RANSACParams.inlierThresholdReprojection = a.thr_r;
    p.RANSACParams.minimalInlierRatioThreshold = a.min_ratio; p.RANSACParams.minimalNumberOfMatches = a.min_matches;
    p.RANSACParams.usedPairs = 3;
    Matcher::parameters& o = p.OpenCVParams;
    o.detector = "ORB"; o.descriptor = "ORB";
    o.gridRows = a.grid_rows; o.gridCols = a.grid_cols; o.useInitialFlow = 0; o.winSize = a.win_size; o.maxLevels = a.max_levels;
    o.maxIter = a.max_iter; o.eps = a.eps; o.minimalTrackedFeatures = a.minimal_tracked_features;
    o.maximalTrackedFeatures = a.maximal_tracked_features; o.minimalReprojDistanceNewTrackingFeatures = a.min_reproj_dist;
    o.minimalEuclidDistanceNewTrackingFeatures = a.min_euclid_dist; o.DBScanEps = a.dbscan_eps;
    o.matchingXYZSphereRadius = 0.12; o.matchingXYZacceptRatioOfBestMatch = 0.55;
    o.trackingErrorThreshold = a.tracking_error_threshold; o.trackingMinEigThreshold = a.tracking_min_eig_threshold;
    o.trackingErrorType = 0; o.removeTooCloseFeatures = a.remove_too_close;
    p.cameraMatrixMat = cv::Mat::zeros(3, 3, CV_32FC1);
    p.cameraMatrixMat.at<float>(0, 0) = a.fx; p.cameraMatrixMat.at<float>(1, 1) = a.fy; p.cameraMatrixMat.at<float>(0, 2) = a.cx;
    p.cameraMatrixMat.at<float>(1, 2) = a.cy; p.cameraMatrixMat.at<float>(2, 2) = 1.f;
    p.distortionCoeffsMat = cv::Mat::zeros(1, 5, CV_32FC1);
    for (int i = 0; i < 5; ++i) p.distortionCoeffsMat.at<float>(i) = a.dist[i];
}

// One VO step of the reference: Matcher::detectInitFeatures(frame 0), then Matcher::runVO(frame 1) -- match() or trackKLT()
// by VOVersion -- on a MatcherB200.  Outputs: T (row-major 4x4), inlier (queryIdx, trainIdx) pairs, the returned ratio, and
// the feature state the call leaves behind (key points of frame 0 / frame 1 as x, y, octave; 3-D points of frame 1).
// Returns the inlier count, < 0 on failure.
TREE_API int tree_run_vo(const uint8_t* rgb0, const uint16_t* depth0, const uint8_t* rgb1, const uint16_t* depth1, int W, int H,
                         int channels, double depth_scale, const TreeArgs* a, uint64_t seed, float* T_out, int* inl_q, int* inl_t,
                         double* ratio_out, int cap, float* kp0 /* cap x 3 */, int* n_kp0, float* kp1 /* cap x 3 */, float* xyz1, int* n_kp1,
                         int* hyp_used) {
    putslam::SensorFrame f0, f1;
    const int type = channels == 3 ? CV_8UC3 : CV_8UC1;
    f0.rgbImage = cv::Mat(H, W, type, (void*)rgb0); f0.depthImage = cv::Mat(H, W, CV_16UC1, (void*)depth0); f0.depthImageScale = depth_scale;
    f1.rgbImage = cv::Mat(H, W, type, (void*)rgb1); f1.depthImage = cv::Mat(H, W, CV_16UC1, (void*)depth1); f1.depthImageScale = depth_scale;
    int mf = -1;
    {   // first pass: learn how many matches RANSAC's filter keeps (the reference prints it), with a throw-away sample stream
        MatcherB200 M("", "");
        configure(M, *a, 1);
        CoutCapture capture;
        g.active = false;
        M.detectInitFeatures(f0);
        Eigen::Matrix4f T; std::vector<cv::DMatch> inl;
        M.runVO(f1, T, inl);
        mf = (int)capture.after("RANSAC: matches.size() = ", -1.0);
    }
    putslam::Matcher* m = putslam::createMatcherB200("", "");           // the factory of the tree file
    MatcherB200& M = *static_cast<MatcherB200*>(m);
    configure(M, *a, 0);
    m->detectInitFeatures(f0);                                          // the reference's own code from here on
    Matcher::featureSet s0 = m->getFeatures();
    *n_kp0 = (int)s0.feature2D.size();
    for (int i = 0; i < *n_kp0 && i < cap; ++i) { kp0[3 * i] = s0.feature2D[(size_t)i].pt.x; kp0[3 * i + 1] = s0.feature2D[(size_t)i].pt.y; kp0[3 * i + 2] = (float)s0.feature2D[(size_t)i].octave; }
    g.active = true; g.seed = seed; g.m = (uint32_t)(mf > 0 ? mf : 1); g.hyp = 0; g.blk = 0; g.w = 4; g.got = 0;
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    std::vector<cv::DMatch> inl;
    const double ratio = m->runVO(f1, T, inl);
    g.active = false;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_out[4 * i + j] = T(i, j);
    *ratio_out = ratio;
    if (hyp_used) *hyp_used = mf < 0 ? 0 : (int)g.hyp;
    Matcher::featureSet s1 = m->getFeatures();
    *n_kp1 = (int)s1.feature2D.size();
    for (int i = 0; i < *n_kp1 && i < cap; ++i) {
        kp1[3 * i] = s1.feature2D[(size_t)i].pt.x; kp1[3 * i + 1] = s1.feature2D[(size_t)i].pt.y; kp1[3 * i + 2] = (float)s1.feature2D[(size_t)i].octave;
        for (int c = 0; c < 3; ++c) xyz1[3 * i + c] = s1.feature3D[(size_t)i][c];
    }
    const int n = (int)inl.size();
    for (int i = 0; i < n && i < cap; ++i) { inl_q[i] = inl[(size_t)i].queryIdx; inl_t[i] = inl[(size_t)i].trainIdx; }
    return n;
}
