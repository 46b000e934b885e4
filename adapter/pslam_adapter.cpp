// pslam_adapter.cpp -- see pslam_adapter.h.  Pure host C++; every numeric stage goes through the C ABI.
#include "pslam_adapter.h"

#include <algorithm>
#include <cmath>
#include <ctime>
#include <iostream>
#include <memory>
#include <set>
#include <unordered_map>

namespace putslam_b200 {
// bytes from one row of a cv::Mat to the next: cv::Mat::step[0] in OpenCV, step0 in the stand-in types of shim/
static inline size_t matRowBytes(const cv::Mat& m) {
#ifdef PSLAM_USE_REAL_HEADERS
    return (size_t)m.step[0];
#else
    return m.step0;
#endif
}


static void logError(pslam_ctx* c, const char* where, int code) {
    std::cerr << "[putslam_b200] " << where << " failed (" << code << "): " << (c ? pslam_last_error(c) : "no context")
              << std::endl;
}

Device::Device(int device) : device_(device) {}
Device::~Device() {
    if (ctx_) pslam_ctx_destroy(ctx_);
}
pslam_ctx* Device::ctx() {
    if (!ctx_) {
        const int r = pslam_ctx_create(device_, &ctx_);
        if (r != PSLAM_OK) {
            std::cerr << "[putslam_b200] pslam_ctx_create(" << device_ << ") failed (" << r
                      << "): no usable sm_100 GPU and no CPU fallback" << std::endl;
            ctx_ = nullptr;
        }
    }
    return ctx_;
}
Device& defaultDevice() {
    static Device d(0);
    return d;
}

static pslam_camera cameraFrom(const cv::Mat& K, const cv::Mat* dist) {
    pslam_camera cam;
    cam.fx = K.at<float>(0, 0); cam.fy = K.at<float>(1, 1); cam.cx = K.at<float>(0, 2); cam.cy = K.at<float>(1, 2);
    for (int i = 0; i < 5; ++i) cam.dist[i] = 0.f;
    if (dist && !dist->empty()) {
        const int n = std::min(5, dist->rows * dist->cols);
        for (int i = 0; i < n; ++i) cam.dist[i] = dist->ptr<float>(0)[i];
    }
    return cam;
}

// A cv::Mat with non-contiguous rows is packed first ("check isContinuous(), else copy").
static const uint8_t* contiguousBytes(const cv::Mat& m, size_t rowBytes, std::vector<uint8_t>& tmp) {
    if (m.empty()) return nullptr;
    if (m.isContinuous()) return m.data;
    tmp.resize(rowBytes * (size_t)m.rows);
    for (int r = 0; r < m.rows; ++r) std::memcpy(tmp.data() + rowBytes * r, m.ptr<uint8_t>(r), rowBytes);
    return tmp.data();
}

// ---- RGBD -----------------------------------------------------------------------------------------
namespace RGBD {
static std::vector<cv::Point2f> undistortImpl(const std::vector<cv::Point2f>& pts, const cv::Mat& K, const cv::Mat& dist) {
    if (pts.empty()) return std::vector<cv::Point2f>();   // RGBD.cpp:259-260
    pslam_ctx* c = defaultDevice().ctx();
    std::vector<cv::Point2f> out(pts.size());
    pslam_camera cam = cameraFrom(K, &dist);
    // undistortion needs no depth: a 1x1 dummy image keeps the entry point single
    uint16_t dummy = 0;
    std::vector<float> xyz(3 * pts.size());
    const int r = c ? pslam_backproject(c, &pts[0].x, (int)pts.size(), &dummy, 1, 1, 1, &cam, 1, 1.0, &out[0].x, xyz.data(),
                                        nullptr, nullptr, nullptr)
                    : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) { logError(c, "removeImageDistortion", r); return std::vector<cv::Point2f>(); }
    return out;
}
std::vector<cv::Point2f> removeImageDistortion(std::vector<cv::Point2f>& features, cv::Mat cameraMatrix, cv::Mat distCoeffs) {
    return undistortImpl(features, cameraMatrix, distCoeffs);
}
std::vector<cv::Point2f> removeImageDistortion(std::vector<cv::KeyPoint>& features, cv::Mat cameraMatrix, cv::Mat distCoeffs) {
    std::vector<cv::Point2f> pts(features.size());   // cv::KeyPoint::convert
    for (size_t i = 0; i < features.size(); ++i) pts[i] = features[i].pt;
    return undistortImpl(pts, cameraMatrix, distCoeffs);
}
std::vector<Eigen::Vector3f> keypoints2Dto3D(std::vector<cv::Point2f> undistortedFeatures2D, cv::Mat depthImage,
                                             cv::Mat cameraMatrix, double depthImageScale, int startingID) {
    const int n = (int)undistortedFeatures2D.size() - startingID;
    std::vector<Eigen::Vector3f> features3D((size_t)std::max(0, n));
    if (n <= 0) return features3D;
    pslam_ctx* c = defaultDevice().ctx();
    pslam_camera cam = cameraFrom(cameraMatrix, nullptr);
    const int stride = (int)(matRowBytes(depthImage) / sizeof(uint16_t));
    const int r = c ? pslam_backproject(c, &undistortedFeatures2D[(size_t)startingID].x, n, depthImage.ptr<uint16_t>(0),
                                        depthImage.cols, depthImage.rows, stride, &cam, 0, depthImageScale, nullptr,
                                        features3D[0].data(), nullptr, nullptr, nullptr)
                    : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) logError(c, "keypoints2Dto3D", r);
    return features3D;
}
}  // namespace RGBD

// ---- RANSAC ---------------------------------------------------------------------------------------
RANSAC::RANSAC(parameters p, cv::Mat cameraMatrix) : RANSACParams(p) {
    seed_ = (uint64_t)time((time_t*)0);   // the reference seeds from the clock (RANSAC.cpp:13); use setSeed for reproducibility
    RANSACParams.iterationCount = 487;    // computeRANSACIteration(0.20), RANSAC.cpp:30
    if (!cameraMatrix.empty()) {
        fx_ = cameraMatrix.at<float>(0, 0); fy_ = cameraMatrix.at<float>(1, 1);
        cx_ = cameraMatrix.at<float>(0, 2); cy_ = cameraMatrix.at<float>(1, 2);
    }
}

static pslam_ransac_params toAbi(const RANSAC::parameters& p, float fx, float fy, float cx, float cy) {
    pslam_ransac_params a;
    a.error_version = p.errorVersion;
    a.inlier_threshold_euclidean = p.inlierThresholdEuclidean;
    a.inlier_threshold_reprojection = p.inlierThresholdReprojection;
    a.minimal_inlier_ratio_threshold = p.minimalInlierRatioThreshold;
    a.minimal_number_of_matches = p.minimalNumberOfMatches;
    a.used_pairs = p.usedPairs;
    a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy;
    return a;
}

Eigen::Matrix4f RANSAC::estimateTransformation(std::vector<Eigen::Vector3f> prevFeatures, std::vector<Eigen::Vector3f> features,
                                               std::vector<cv::DMatch> matches, std::vector<cv::DMatch>& bestInlierMatches) {
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    const int m = (int)matches.size();
    std::vector<int> mq((size_t)m), mt((size_t)m), inl((size_t)std::max(1, m));
    for (int k = 0; k < m; ++k) { mq[k] = matches[k].queryIdx; mt[k] = matches[k].trainIdx; }
    int nInl = 0;
    pslam_ransac_params a = toAbi(RANSACParams, fx_, fy_, cx_, cy_);
    pslam_ctx* c = defaultDevice().ctx();
    const int r = c ? pslam_ransac_estimate(c, prevFeatures.empty() ? nullptr : prevFeatures[0].data(), (int)prevFeatures.size(),
                                            features.empty() ? nullptr : features[0].data(), (int)features.size(), mq.data(),
                                            mt.data(), m, &a, seed_, numHyp_, T.data(), inl.data(), &nInl, &bestRatio_, &hypUsed_)
                    : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) {   // reference failure value: identity + no inliers (RANSAC.cpp:78-79,162-163)
        logError(c, "RANSAC::estimateTransformation", r);
        bestInlierMatches.clear();
        return Eigen::Matrix4f::Identity();
    }
    bestInlierMatches.clear();
    for (int i = 0; i < nInl; ++i) bestInlierMatches.push_back(matches[(size_t)inl[i]]);
    return T;
}

double RANSAC::pointInlierRatio(std::vector<cv::DMatch>& inlierMatches, std::vector<cv::DMatch>& allMatches) {
    std::set<int> inlier, all;   // RANSAC.h:56-66 verbatim semantics
    for (auto& m : allMatches) all.insert(m.trainIdx);
    for (auto& in : inlierMatches) inlier.insert(in.trainIdx);
    return double(inlier.size()) / double(all.size());
}

// ---- Matcher ---------------------------------------------------------------------------------------
std::vector<cv::DMatch> MatcherB200::performMatching(cv::Mat prevDescriptors, cv::Mat descriptors) {
    std::vector<cv::DMatch> out;
    const int nq = prevDescriptors.rows, nt = descriptors.rows;
    if (nq == 0 || nt == 0) return out;
    std::vector<uint8_t> tq, tt;
    const uint8_t* q = contiguousBytes(prevDescriptors, (size_t)prevDescriptors.cols, tq);
    const uint8_t* t = contiguousBytes(descriptors, (size_t)descriptors.cols, tt);
    const int cap = std::min(nq, nt);
    std::vector<int> oq((size_t)cap), ot((size_t)cap);
    std::vector<float> od((size_t)cap);
    int n = 0;
    pslam_ctx* c = dev_.ctx();
    const int r = c ? pslam_match_bf_mutual(c, q, nq, t, nt, prevDescriptors.cols, oq.data(), ot.data(), od.data(), &n)
                    : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) { logError(c, "performMatching", r); return out; }
    out.reserve((size_t)n);
    for (int i = 0; i < n; ++i) out.push_back(cv::DMatch(oq[i], ot[i], 0, od[i]));   // imgIdx = 0 like BFMatcher
    return out;
}

std::vector<cv::DMatch> MatcherB200::performMatchingRatio(cv::Mat prevDescriptors, cv::Mat descriptors, float ratio) {
    std::vector<cv::DMatch> out;
    const int nq = prevDescriptors.rows, nt = descriptors.rows;
    if (nq == 0 || nt < 2) return out;
    std::vector<uint8_t> tq, tt;
    const uint8_t* q = contiguousBytes(prevDescriptors, (size_t)prevDescriptors.cols, tq);
    const uint8_t* t = contiguousBytes(descriptors, (size_t)descriptors.cols, tt);
    std::vector<int> idx(2 * (size_t)nq);
    std::vector<float> dist(2 * (size_t)nq);
    pslam_ctx* c = dev_.ctx();
    const int r = c ? pslam_match_knn2(c, q, nq, t, nt, prevDescriptors.cols, idx.data(), dist.data()) : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) { logError(c, "performMatchingRatio", r); return out; }
    for (int i = 0; i < nq; ++i)
        if (dist[2 * i] < ratio * dist[2 * i + 1]) out.push_back(cv::DMatch(i, idx[2 * i], 0, dist[2 * i]));
    return out;
}

// predicted pyramid level, all double, host libm -- matcher.cpp:641-651 / 682-692, matcher.h:26-28
static int predictedLevel(int detLevel, double detDist, double curDist) {
    static const double scaleFactor = 1.2;
    static const double logScaleFactor = std::log(scaleFactor);
    const int nLevels = 8;
    double detLevelScaleFactor = pow(scaleFactor, detLevel);
    double curLevelScaleFactor = detLevelScaleFactor * detDist / curDist;
    int curLevel = (int)std::ceil(std::log(curLevelScaleFactor) / logScaleFactor);
    curLevel = std::max(0, curLevel);
    curLevel = std::min(nLevels - 1, curLevel);
    return curLevel;
}

void MatcherB200::unpinMapSide() {
    pslam_ctx* c = dev_.ctx();
    for (const void* p : pinnedMap_) if (c) pslam_host_unregister(c, p);
    pinnedMap_.clear();
}

bool MatcherB200::pinMapSide(const MapSide& map) {
    unpinMapSide();
    pslam_ctx* c = dev_.ctx();
    if (!c || map.octave.empty()) return false;
    const size_t M = map.octave.size();
    struct R { const void* p; size_t bytes; } ranges[4] = {
        {map.xyz.data(), 24 * M}, {map.descriptors.data, map.descriptors.isContinuous() ? 32 * M : 0},
        {map.octave.data(), 4 * M}, {map.detDist.data(), 8 * M}};
    bool all = true;
    for (const R& r : ranges) {
        if (!r.p || !r.bytes) { all = false; continue; }
        if (pslam_host_register(c, r.p, r.bytes) == PSLAM_OK) pinnedMap_.push_back(r.p);
        else all = false;
    }
    return all;
}

double MatcherB200::matchXYZCore(const MapSide& map, cv::Mat currentPoseDescriptors,
                                 std::vector<Eigen::Vector3f>& currentPoseFeatures3D,
                                 std::vector<cv::KeyPoint>& currentPoseKeyPoints, std::vector<double>& currentPoseDetDists,
                                 double matchingXYZSphereRadius, double matchingXYZacceptRatioOfBestMatch, int computationNumber,
                                 const RANSAC::parameters& ransacParams, cv::Mat cameraMatrix,
                                 Eigen::Matrix4f& estimatedTransformation, std::vector<cv::DMatch>& matches,
                                 std::vector<cv::DMatch>& inlierMatches, bool xorDistance) {
    matches.clear();
    inlierMatches.clear();
    if (computationNumber > 1) {   // matcher.cpp:619-622
        matchingXYZSphereRadius += 0.02 * (computationNumber - 1);
        matchingXYZacceptRatioOfBestMatch = std::max(0.1, matchingXYZacceptRatioOfBestMatch - 0.05 * (computationNumber - 1));
    }
    const int N = (int)currentPoseKeyPoints.size(), M = (int)map.octave.size();
    // no matches possible -> -1.0 like matcher.cpp:755-756; also keeps &v[0][0] off empty vectors.  Inconsistent list sizes
    // (the reference asserts them, matcher.cpp:631-636) are refused instead of read out of bounds.
    if (N == 0 || M == 0) return -1.0;
    if ((int)currentPoseFeatures3D.size() != N || (int)currentPoseDetDists.size() != N || currentPoseDescriptors.rows != N ||
        currentPoseDescriptors.cols * (int)currentPoseDescriptors.elemSize() != 32 || map.descriptors.rows != M ||
        (int)map.xyz.size() != 3 * M || (int)map.detDist.size() != M) {
        logError(dev_.ctx(), "matchXYZ", PSLAM_ERR_ARG);
        return -1.0;
    }
    std::vector<int> curLevels, mapLevels, curOct;
    std::vector<float> mapXyz;
    if (hostLevels_) {
        curLevels.resize((size_t)N); mapLevels.resize((size_t)M); mapXyz.resize(3 * (size_t)M);
        for (int i = 0; i < N; ++i) {   // matcher.cpp:639-651: curDist = Vector3f::norm() (float) widened
            const float x = currentPoseFeatures3D[i][0], y = currentPoseFeatures3D[i][1], z = currentPoseFeatures3D[i][2];
            const float yy = y * y, zz = z * z;
            const double curDist = std::sqrt(x * x + (yy + zz));   // float expression -> sqrtf -> double
            curLevels[i] = predictedLevel(currentPoseKeyPoints[i].octave, currentPoseDetDists[i], curDist);
        }
        for (int j = 0; j < M; ++j) {   // matcher.cpp:682-692 (double norm of the double position), :665 (cast to float)
            const double px = map.xyz[3 * j], py = map.xyz[3 * j + 1], pz = map.xyz[3 * j + 2];
            const double curDist = std::sqrt(px * px + py * py + pz * pz);
            mapLevels[j] = predictedLevel(map.octave[j], map.detDist[j], curDist);
            mapXyz[3 * j] = (float)px; mapXyz[3 * j + 1] = (float)py; mapXyz[3 * j + 2] = (float)pz;
        }
    } else {
        curOct.resize((size_t)N);
        for (int i = 0; i < N; ++i) curOct[i] = currentPoseKeyPoints[i].octave;
    }
    std::vector<uint8_t> tm, tc;
    const uint8_t* md = contiguousBytes(map.descriptors, 32, tm);
    const uint8_t* cd = contiguousBytes(currentPoseDescriptors, 32, tc);
    RANSAC::parameters rp = ransacParams;
    rp.errorVersion = rp.errorVersionMap;   // matcher.cpp:760-761
    float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f;
    if (!cameraMatrix.empty()) {
        fx = cameraMatrix.at<float>(0, 0); fy = cameraMatrix.at<float>(1, 1);
        cx = cameraMatrix.at<float>(0, 2); cy = cameraMatrix.at<float>(1, 2);
    }
    pslam_ransac_params a = toAbi(rp, fx, fy, cx, cy);
    int cap = std::max(2048, 2 * N);   // typical yield is ~1 match per current keypoint; grown on truncation
    std::vector<int> mq, mt, inl;
    std::vector<float> mdist;
    pslam_frame_result res;
    pslam_ctx* c = dev_.ctx();
    int r = PSLAM_ERR_NO_DEVICE;
    for (int attempt = 0; c && attempt < 6; ++attempt) {   // the reference's match vector is unbounded: grow on truncation
        mq.resize((size_t)cap); mt.resize((size_t)cap); mdist.resize((size_t)cap); inl.resize((size_t)cap);
        if (hostLevels_)
            r = pslam_frame_to_map(c, mapXyz.data(), md, mapLevels.data(), M, &currentPoseFeatures3D[0][0], cd, curLevels.data(), N,
                                   matchingXYZSphereRadius, matchingXYZacceptRatioOfBestMatch, xorDistance ? 1 : 0, &a, seed_,
                                   numHyp_, cap, mq.data(), mt.data(), mdist.data(), inl.data(), &res);
        else
            r = pslam_frame_to_map_features(c, map.xyz.data(), md, map.octave.data(), map.detDist.data(), M,
                                            &currentPoseFeatures3D[0][0], cd, curOct.data(), currentPoseDetDists.data(), N,
                                            matchingXYZSphereRadius, matchingXYZacceptRatioOfBestMatch, xorDistance ? 1 : 0, &a,
                                            seed_, numHyp_, cap, mq.data(), mt.data(), mdist.data(), inl.data(), &res);
        if (r != PSLAM_ERR_CAPACITY) break;
        cap = res.n_matches + 16;
    }
    estimatedTransformation = Eigen::Matrix4f::Identity();
    if (r != PSLAM_OK) { logError(c, "matchXYZ", r); return -1.0; }
    if (res.n_matches <= 0) return -1.0;   // matcher.cpp:755-756
    matches.reserve((size_t)res.n_matches);
    inlierMatches.reserve((size_t)res.n_inliers);
    for (int k = 0; k < res.n_matches; ++k) matches.push_back(cv::DMatch(mq[k], mt[k], -1, mdist[k]));
    for (int k = 0; k < res.n_inliers; ++k) inlierMatches.push_back(matches[(size_t)inl[k]]);
    std::memcpy(estimatedTransformation.data(), res.T, sizeof(res.T));
    return res.inlier_ratio;   // == RANSAC::pointInlierRatio(inlierMatches, matches), matcher.cpp:797 (computed with a bitmap)
}

// ---- DBScan ------------------------------------------------------------------------------------------------------------
void DBScan::run(std::vector<cv::KeyPoint>& pts) {
    const int n = (int)pts.size();
    // (float)cv::norm(a.pt - b.pt) < eps : float differences, double sqrt of the double sum of squares, rounded to float,
    // compared in double (dbscan.cpp:90-92 and the `dist[..] < eps` tests)
    const auto close = [&](int a, int b) {
        const float dx = pts[(size_t)a].pt.x - pts[(size_t)b].pt.x, dy = pts[(size_t)a].pt.y - pts[(size_t)b].pt.y;
        return (double)(float)std::sqrt((double)dx * dx + (double)dy * dy) < eps_;
    };
    label_.assign((size_t)n, 0);                 // 0: not labelled yet
    std::vector<char> seen((size_t)n, 0);
    std::vector<int> frontier, fresh;
    int next = 1;
    for (int seed = 0; seed < n; ++seed) {
        if (seen[(size_t)seed]) continue;
        seen[(size_t)seed] = 1;
        frontier.clear();
        for (int k = 0; k < n; ++k) if (close(seed, k)) frontier.push_back(k);      // the seed itself and visited points count
        if ((int)frontier.size() < minPts_) { label_[(size_t)seed] = -1; continue; }
        label_[(size_t)seed] = next;
        // the frontier grows while it is walked; a point that was noise keeps its -1 (dbscan.cpp:45-46)
        for (size_t j = 0; j < frontier.size(); ++j) {
            const int x = frontier[j];
            if (!seen[(size_t)x]) {
                seen[(size_t)x] = 1;
                fresh.clear();
                for (int k = 0; k < n; ++k) if (!seen[(size_t)k] && close(x, k)) fresh.push_back(k);
                if ((int)fresh.size() >= minPts_) frontier.insert(frontier.end(), fresh.begin(), fresh.end());
            }
            if (label_[(size_t)x] == 0) label_[(size_t)x] = next;
        }
        ++next;
    }
    // the first perCluster_ members of every cluster stay, in list order; noise stays
    std::vector<int> taken((size_t)next, 0);
    size_t out = 0;
    for (int i = 0; i < n; ++i) {
        const int c = label_[(size_t)i];
        if (c > 0 && taken[(size_t)c]++ >= perCluster_) continue;
        if (out != (size_t)i) pts[out] = pts[(size_t)i];
        ++out;
    }
    pts.resize(out);
}

std::vector<cv::KeyPoint> MatcherB200::detectFeatures(cv::Mat rgbImage, int gridCols, int gridRows, int maximalTrackedFeatures) {
    return detectGrid(rgbImage, gridCols, gridRows, maximalTrackedFeatures, false);
}
std::vector<cv::KeyPoint> MatcherB200::detectFeaturesFAST(cv::Mat rgbImage, int gridCols, int gridRows, int maximalTrackedFeatures) {
    return detectGrid(rgbImage, gridCols, gridRows, maximalTrackedFeatures, true);
}

// MatcherOpenCV::detectFeatures (src/Matcher/matcherOpenCV.cpp:118-176) around the device detector
std::vector<cv::KeyPoint> MatcherB200::detectGrid(cv::Mat rgbImage, int gridCols, int gridRows, int maximalTrackedFeatures,
                                                  bool fast) {
    std::vector<cv::KeyPoint> raw_keypoints;
    pslam_ctx* c = dev_.ctx();
    if (!c || rgbImage.empty() || gridCols <= 0 || gridRows <= 0) {
        if (!c) logError(c, "detectFeatures", PSLAM_ERR_NO_DEVICE);
        return raw_keypoints;
    }
    const auto compare_response = [](const cv::KeyPoint& p1, const cv::KeyPoint& p2) { return p1.response > p2.response; };
    const int W = rgbImage.cols, H = rgbImage.rows, ch = rgbImage.channels();
    const int rowBytes = (int)matRowBytes(rgbImage);
    const int maximalFeaturesInROI = maximalTrackedFeatures * 3 / (gridCols * gridRows);
    lastFrameData_ = nullptr;   // set again below, only once the device really holds this frame
    // cv::ORB::create() keeps 500 per call (tied responses can add a few); cv::FAST has no budget: strict 3x3 maxima,
    // at most one per 2x2 pixels
    const int w = W / gridCols, h = H / gridRows;
    const int cap = fast ? w * h / 4 + 64 : 4096;
    std::vector<float> xy(2 * (size_t)cap), size(fast ? 1 : (size_t)cap), angle(fast ? 1 : (size_t)cap), response((size_t)cap);
    std::vector<int> octave(fast ? 1 : (size_t)cap);
    for (int k = 0; k < gridCols; k++) {
        for (int i = 0; i < gridRows; i++) {
            const int x0 = k * W / gridCols, y0 = i * H / gridRows;
            const unsigned char* roi = rgbImage.data + (size_t)y0 * rowBytes + (size_t)x0 * ch;
            int n = 0;
            const int r = fast ? pslam_fast_detect(c, roi, w, h, rowBytes, ch, /*COLOR_RGB2GRAY, :122*/ 1, 10, xy.data(), response.data(),
                                                   cap, &n)
                               : pslam_orb_detect(c, roi, w, h, rowBytes, ch, 1, 500, xy.data(), size.data(), angle.data(),
                                                  response.data(), octave.data(), cap, &n);
            if (r != PSLAM_OK) { logError(c, "detectFeatures", r); continue; }
            if (!fast && gridCols == 1 && gridRows == 1) {   // the whole frame went to the device in one piece: describeFeatures may reuse it
                lastFrameData_ = rgbImage.data; lastFrameRows_ = H; lastFrameCols_ = W; lastFrameStep_ = rowBytes; lastFrameCh_ = ch;
            }
            std::vector<cv::KeyPoint> keypointsInROI((size_t)n);
            for (int j = 0; j < n; ++j) {
                cv::KeyPoint& kp = keypointsInROI[(size_t)j];
                kp.pt = cv::Point2f(xy[2 * j], xy[2 * j + 1]);
                kp.response = response[j]; kp.class_id = -1;
                if (fast) { kp.size = 7.f; kp.angle = -1.f; kp.octave = 0; }
                else { kp.size = size[j]; kp.angle = angle[j]; kp.octave = octave[j]; }
            }
            std::sort(keypointsInROI.begin(), keypointsInROI.end(), compare_response);
            for (size_t j = 0; j < keypointsInROI.size() && (int)j < maximalFeaturesInROI; j++) {
                keypointsInROI[j].pt.x += float(x0);
                keypointsInROI[j].pt.y += float(y0);
                raw_keypoints.push_back(keypointsInROI[j]);
            }
        }
    }
    std::sort(raw_keypoints.begin(), raw_keypoints.end(), compare_response);
    if ((int)raw_keypoints.size() > maximalTrackedFeatures) raw_keypoints.resize((size_t)maximalTrackedFeatures);
    return raw_keypoints;
}

cv::Mat MatcherB200::describeFeatures(cv::Mat rgbImage, std::vector<cv::KeyPoint>& features) {
    cv::Mat descriptors;
    pslam_ctx* c = dev_.ctx();
    const int n = (int)features.size();
    if (!c || rgbImage.empty() || n == 0) {
        if (!c) logError(c, "describeFeatures", PSLAM_ERR_NO_DEVICE);
        features.clear();   // cv::ORB::compute leaves no keypoints when it cannot describe any
        return descriptors;
    }
    std::vector<float> xy(2 * (size_t)n), ang((size_t)n);
    std::vector<int> oct((size_t)n), order((size_t)n);
    for (int i = 0; i < n; ++i) {
        xy[2 * i] = features[i].pt.x; xy[2 * i + 1] = features[i].pt.y;
        oct[i] = features[i].octave; ang[i] = features[i].angle;
    }
    descriptors.create(n, 32, CV_8U);
    int nOut = 0;
    const int rowBytes = (int)matRowBytes(rgbImage);
    const bool resident = reuseFrame_ && lastFrameData_ == rgbImage.data && lastFrameRows_ == rgbImage.rows &&
                          lastFrameCols_ == rgbImage.cols && lastFrameStep_ == rowBytes && lastFrameCh_ == rgbImage.channels();
    const int r = pslam_orb_describe(c, resident ? nullptr : rgbImage.data, rgbImage.cols, rgbImage.rows, rowBytes,
                                     rgbImage.channels(), xy.data(),
                                     oct.data(), ang.data(), n, order.data(), &nOut, descriptors.data);
    if (r != PSLAM_OK) {
        logError(c, "describeFeatures", r);
        features.clear();
        lastFrameData_ = nullptr;
        return cv::Mat();
    }
    if (!resident) {   // the frame resident on the device is now THIS one (uploaded by the call above)
        lastFrameData_ = rgbImage.data; lastFrameRows_ = rgbImage.rows; lastFrameCols_ = rgbImage.cols; lastFrameStep_ = rowBytes;
        lastFrameCh_ = rgbImage.channels();
    }
    std::vector<cv::KeyPoint> kept((size_t)nOut);
    for (int k = 0; k < nOut; ++k) kept[(size_t)k] = features[(size_t)order[(size_t)k]];
    features.swap(kept);
    if (nOut == 0) return cv::Mat();
    cv::Mat out(nOut, 32, CV_8U);
    std::memcpy(out.data, descriptors.data, 32 * (size_t)nOut);
    return out;
}

std::vector<cv::DMatch> MatcherB200::performTracking(cv::Mat prevImg, cv::Mat img, std::vector<cv::Point2f>& prevFeatures,
                                                     std::vector<cv::Point2f>& features, std::vector<cv::KeyPoint>& prevKeyPoints,
                                                     std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& prevDetDists,
                                                     std::vector<double>& detDists) {
    std::vector<cv::DMatch> matches;
    pslam_ctx* c = dev_.ctx();
    const int n = (int)prevFeatures.size();
    const TrackingParams& tp = tracking_;
    const bool init = tp.useInitialFlow > 0;
    auto giveUp = [&](int code) {
        if (code != PSLAM_OK) logError(c, "performTracking", code);
        features.clear(); keyPoints.clear(); detDists.clear();
        lastTrackedData_ = nullptr;
        return matches;
    };
    if (!c) return giveUp(PSLAM_ERR_NO_DEVICE);
    if (prevImg.empty() || img.empty() || prevImg.rows != img.rows || prevImg.cols != img.cols || prevImg.channels() != img.channels() ||
        (init && (int)features.size() != n) || (int)prevKeyPoints.size() != n || (int)prevDetDists.size() != n)
        return giveUp(PSLAM_ERR_ARG);
    const int rowBytes = (int)matRowBytes(img), prevRowBytes = (int)matRowBytes(prevImg);
    if (prevRowBytes != rowBytes) return giveUp(PSLAM_ERR_ARG);
    std::vector<float> prevXY(2 * (size_t)n + 2), curXY(2 * (size_t)n + 2), err((size_t)n + 1);
    std::vector<unsigned char> status((size_t)n + 1);
    std::vector<int> kept((size_t)n + 1);
    for (int i = 0; i < n; ++i) {
        prevXY[2 * i] = prevFeatures[(size_t)i].x; prevXY[2 * i + 1] = prevFeatures[(size_t)i].y;
        if (init) { curXY[2 * i] = features[(size_t)i].x; curXY[2 * i + 1] = features[(size_t)i].y; }
    }
    const int flags = (init ? PSLAM_KLT_USE_INITIAL_FLOW : 0) | (tp.trackingErrorType > 0 ? PSLAM_KLT_GET_MIN_EIGENVALS : 0);
    const bool resident = reuseTracked_ && lastTrackedData_ == prevImg.data && lastTrackedRows_ == prevImg.rows &&
                          lastTrackedCols_ == prevImg.cols && lastTrackedStep_ == prevRowBytes && lastTrackedCh_ == prevImg.channels() &&
                          lastTrackedLevels_ >= tp.maxLevels;
    int nKept = 0;
    const int r = pslam_klt_perform_tracking(c, resident ? nullptr : prevImg.data, img.data, img.cols, img.rows, rowBytes,
                                             img.channels(), prevXY.data(), curXY.data(), n, tp.winSize, tp.maxLevels,
                                             3 /* CV_TERMCRIT_ITER | CV_TERMCRIT_EPS */, tp.maxIter, tp.eps, flags,
                                             tp.trackingMinEigThreshold, tp.trackingErrorThreshold,
                                             tp.minimalReprojDistanceNewTrackingFeatures, status.data(), err.data(), kept.data(),
                                             &nKept);
    if (r != PSLAM_OK) return giveUp(r);
    lastTrackedData_ = img.data; lastTrackedRows_ = img.rows; lastTrackedCols_ = img.cols; lastTrackedStep_ = rowBytes;
    lastTrackedCh_ = img.channels(); lastTrackedLevels_ = tp.maxLevels;
    features.resize((size_t)nKept); keyPoints.resize((size_t)nKept); detDists.resize((size_t)nKept);
    matches.reserve((size_t)nKept);
    for (int j = 0; j < nKept; ++j) {
        const int i = kept[(size_t)j];
        features[(size_t)j] = cv::Point2f(curXY[2 * i], curXY[2 * i + 1]);
        keyPoints[(size_t)j] = prevKeyPoints[(size_t)i];        // keyPoints = prevKeyPoints with the new positions (:238-242)
        keyPoints[(size_t)j].pt = features[(size_t)j];
        detDists[(size_t)j] = prevDetDists[(size_t)i];
        matches.push_back(cv::DMatch(i, j, 0));
    }
    return matches;
}

bool MatcherB200::uploadMapFeatures(int first, const MapSide& features, const std::vector<float>& viewAxis) {
    pslam_ctx* c = dev_.ctx();
    const int M = (int)features.octave.size();
    if (!c) return false;
    if ((int)features.xyz.size() != 3 * M || (int)features.detDist.size() != M || (int)viewAxis.size() != 3 * M ||
        features.descriptors.rows != M) {
        std::cerr << "[putslam_b200] uploadMapFeatures: inconsistent feature arrays" << std::endl;
        return false;
    }
    std::vector<uint8_t> tm;
    const uint8_t* md = contiguousBytes(features.descriptors, 32, tm);
    const int r = pslam_map_write(c, first, M, features.xyz.data(), md, features.octave.data(), features.detDist.data(),
                                  viewAxis.data());
    if (r != PSLAM_OK) { logError(c, "uploadMapFeatures", r); return false; }
    return true;
}

bool MatcherB200::updateMapPositions(int first, const std::vector<double>& xyz) {
    pslam_ctx* c = dev_.ctx();
    if (!c) return false;
    const int r = pslam_map_write(c, first, (int)(xyz.size() / 3), xyz.data(), nullptr, nullptr, nullptr, nullptr);
    if (r != PSLAM_OK) { logError(c, "updateMapPositions", r); return false; }
    return true;
}

bool MatcherB200::truncateMap(int nFeatures) {
    pslam_ctx* c = dev_.ctx();
    if (!c) return false;
    const int r = pslam_map_truncate(c, nFeatures);
    if (r != PSLAM_OK) { logError(c, "truncateMap", r); return false; }
    return true;
}

int MatcherB200::mapSize() {
    pslam_ctx* c = dev_.ctx();
    int n = 0;
    if (c) pslam_map_size(c, &n);
    return n;
}

double MatcherB200::matchXYZResident(const double cameraPose[16], const MapFilter& filter, cv::Mat currentPoseDescriptors,
                                     std::vector<Eigen::Vector3f>& currentPoseFeatures3D,
                                     std::vector<cv::KeyPoint>& currentPoseKeyPoints, std::vector<double>& currentPoseDetDists,
                                     double matchingXYZSphereRadius, double matchingXYZacceptRatioOfBestMatch,
                                     int computationNumber, const RANSAC::parameters& ransacParams, cv::Mat cameraMatrix,
                                     Eigen::Matrix4f& estimatedTransformation, std::vector<int>& keptFeatures,
                                     std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& inlierMatches,
                                     bool xorDistance) {
    matches.clear();
    inlierMatches.clear();
    keptFeatures.clear();
    estimatedTransformation = Eigen::Matrix4f::Identity();
    if (computationNumber > 1) {   // matcher.cpp:619-622
        matchingXYZSphereRadius += 0.02 * (computationNumber - 1);
        matchingXYZacceptRatioOfBestMatch = std::max(0.1, matchingXYZacceptRatioOfBestMatch - 0.05 * (computationNumber - 1));
    }
    pslam_ctx* c = dev_.ctx();
    if (!c) { logError(c, "matchXYZResident", PSLAM_ERR_NO_DEVICE); return -1.0; }
    const int N = (int)currentPoseKeyPoints.size(), M = mapSize();
    std::vector<int> curOct((size_t)N);
    for (int i = 0; i < N; ++i) curOct[i] = currentPoseKeyPoints[i].octave;
    std::vector<uint8_t> tc;
    const uint8_t* cd = contiguousBytes(currentPoseDescriptors, 32, tc);
    RANSAC::parameters rp = ransacParams;
    rp.errorVersion = rp.errorVersionMap;   // matcher.cpp:760-761
    float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f;
    if (!cameraMatrix.empty()) {
        fx = cameraMatrix.at<float>(0, 0); fy = cameraMatrix.at<float>(1, 1);
        cx = cameraMatrix.at<float>(0, 2); cy = cameraMatrix.at<float>(1, 2);
    }
    pslam_ransac_params a = toAbi(rp, fx, fy, cx, cy);
    pslam_map_prepare_params prep;
    prep.fx = filter.fx; prep.fy = filter.fy; prep.cx = filter.cx; prep.cy = filter.cy;
    prep.image_w = filter.imageW; prep.image_h = filter.imageH;
    prep.max_angle = filter.maxAngleBetweenFrames; prep.max_z = filter.maxZ;
    int cap = std::max(2048, 2 * N);
    std::vector<int> mq, mt, inl;
    std::vector<float> mdist;
    keptFeatures.resize((size_t)std::max(1, M));
    int nKept = 0;
    pslam_frame_result res;
    int r = PSLAM_ERR_NO_DEVICE;
    for (int attempt = 0; attempt < 6; ++attempt) {   // the reference's match vector is unbounded: grow on truncation
        mq.resize((size_t)cap); mt.resize((size_t)cap); mdist.resize((size_t)cap); inl.resize((size_t)cap);
        r = pslam_frame_to_resident_map(c, cameraPose, &prep, N ? &currentPoseFeatures3D[0][0] : nullptr, cd, curOct.data(),
                                        currentPoseDetDists.data(), N, matchingXYZSphereRadius,
                                        matchingXYZacceptRatioOfBestMatch, xorDistance ? 1 : 0, &a, seed_, numHyp_, cap,
                                        keptFeatures.data(), &nKept, nullptr, nullptr, mq.data(), mt.data(), mdist.data(),
                                        inl.data(), &res);
        if (r != PSLAM_ERR_CAPACITY) break;
        cap = res.n_matches + 16;
    }
    keptFeatures.resize((size_t)nKept);
    if (r != PSLAM_OK) { logError(c, "matchXYZResident", r); return -1.0; }
    if (res.n_matches <= 0) return -1.0;   // matcher.cpp:755-756
    matches.reserve((size_t)res.n_matches);
    inlierMatches.reserve((size_t)res.n_inliers);
    for (int k = 0; k < res.n_matches; ++k) matches.push_back(cv::DMatch(mq[k], mt[k], -1, mdist[k]));
    for (int k = 0; k < res.n_inliers; ++k) inlierMatches.push_back(matches[(size_t)inl[k]]);
    std::memcpy(estimatedTransformation.data(), res.T, sizeof(res.T));
    return res.inlier_ratio;
}

double MatcherB200::matchCore(cv::Mat prevDescriptors, const std::vector<Eigen::Vector3f>& prevFeatures3D, cv::Mat descriptors,
                              const std::vector<cv::KeyPoint>& keyPoints, cv::Mat depthImage, double depthImageScale,
                              cv::Mat cameraMatrix, cv::Mat distCoeffs, const RANSAC::parameters& ransacParams,
                              std::vector<cv::Point2f>& undistortedFeatures2D, std::vector<Eigen::Vector3f>& features3D,
                              std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& inlierMatches,
                              Eigen::Matrix4f& estimatedTransformation) {
    matches.clear(); inlierMatches.clear();
    estimatedTransformation = Eigen::Matrix4f::Identity();
    const int nPrev = prevDescriptors.rows, nCur = descriptors.rows;
    undistortedFeatures2D.assign((size_t)nCur, cv::Point2f());
    features3D.assign((size_t)nCur, Eigen::Vector3f());
    if (nCur == 0) return 0.0 / 0.0;   // pointInlierRatio of two empty sets
    std::vector<uint8_t> tp, tc;
    const uint8_t* pd = contiguousBytes(prevDescriptors, 32, tp);
    const uint8_t* cd = contiguousBytes(descriptors, 32, tc);
    std::vector<float> uv(2 * (size_t)nCur);
    for (int i = 0; i < nCur; ++i) { uv[2 * i] = keyPoints[(size_t)i].pt.x; uv[2 * i + 1] = keyPoints[(size_t)i].pt.y; }
    RANSAC::parameters rp = ransacParams;
    rp.errorVersion = rp.errorVersionVO;   // matcher.cpp:491-492
    pslam_camera cam = cameraFrom(cameraMatrix, &distCoeffs);
    pslam_ransac_params a = toAbi(rp, cam.fx, cam.fy, cam.cx, cam.cy);
    const int cap = std::max(1, std::min(nPrev, nCur));
    std::vector<int> mq((size_t)cap), mt((size_t)cap), inl((size_t)cap);
    std::vector<float> md((size_t)cap);
    pslam_frame_result res;
    pslam_ctx* c = dev_.ctx();
    const int stride = (int)(matRowBytes(depthImage) / sizeof(uint16_t));
    const int r = c ? pslam_frame_to_frame(c, pd, nPrev ? prevFeatures3D[0].data() : nullptr, nPrev, cd, uv.data(), nCur,
                                           depthImage.ptr<uint16_t>(0), depthImage.cols, depthImage.rows, stride, &cam,
                                           distCoeffs.empty() ? 0 : 1, depthImageScale, &a, seed_, numHyp_,
                                           features3D[0].data(), &undistortedFeatures2D[0].x, nullptr, mq.data(), mt.data(),
                                           md.data(), inl.data(), &res)
                    : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) { logError(c, "match", r); return 0.0; }
    matches.reserve((size_t)res.n_matches);
    for (int k = 0; k < res.n_matches; ++k) matches.push_back(cv::DMatch(mq[k], mt[k], 0, md[k]));
    for (int k = 0; k < res.n_inliers; ++k) inlierMatches.push_back(matches[(size_t)inl[k]]);
    std::memcpy(estimatedTransformation.data(), res.T, sizeof(res.T));
    return res.inlier_ratio;
}

double MatcherB200::trackKLTCore(cv::Mat prevRgbImage, cv::Mat rgbImage, const std::vector<cv::Point2f>& prevFeaturesDistorted,
                                 const std::vector<Eigen::Vector3f>& prevFeatures3D, const std::vector<cv::KeyPoint>& prevKeyPoints,
                                 const std::vector<double>& prevDetDists, cv::Mat depthImage, double depthImageScale,
                                 cv::Mat cameraMatrix, cv::Mat distCoeffs, const RANSAC::parameters& ransacParams,
                                 std::vector<cv::Point2f>& distortedFeatures2D, std::vector<cv::Point2f>& undistortedFeatures2D,
                                 std::vector<Eigen::Vector3f>& features3D, std::vector<cv::KeyPoint>& keyPoints,
                                 std::vector<double>& detDists, std::vector<cv::DMatch>& matches,
                                 std::vector<cv::DMatch>& inlierMatches, Eigen::Matrix4f& estimatedTransformation) {
    distortedFeatures2D.clear(); undistortedFeatures2D.clear(); features3D.clear(); keyPoints.clear(); detDists.clear();
    matches.clear(); inlierMatches.clear();
    estimatedTransformation = Eigen::Matrix4f::Identity();
    const int n = (int)prevFeaturesDistorted.size();
    if (n == 0) return 0.0;                                    // "No features so identity()" (matcher.cpp:146-148)
    pslam_ctx* c = dev_.ctx();
    const TrackingParams& tp = tracking_;
    if (!c) { logError(c, "trackKLT", PSLAM_ERR_NO_DEVICE); return 0.0; }
    if (prevRgbImage.empty() || rgbImage.empty() || depthImage.empty() || prevRgbImage.rows != rgbImage.rows ||
        prevRgbImage.cols != rgbImage.cols || prevRgbImage.channels() != rgbImage.channels() || depthImage.rows != rgbImage.rows ||
        depthImage.cols != rgbImage.cols || (int)prevFeatures3D.size() != n || (int)prevKeyPoints.size() != n ||
        (int)prevDetDists.size() != n || tp.useInitialFlow > 0) {   // trackKLT passes an empty `features`: no initial flow
        logError(c, "trackKLT", PSLAM_ERR_ARG);
        return 0.0;
    }
    const int rowBytes = (int)matRowBytes(rgbImage), prevRowBytes = (int)matRowBytes(prevRgbImage);
    const int stride = (int)(matRowBytes(depthImage) / sizeof(uint16_t));
    if (prevRowBytes != rowBytes) { logError(c, "trackKLT", PSLAM_ERR_ARG); return 0.0; }
    RANSAC::parameters rp = ransacParams;
    rp.errorVersion = rp.errorVersionVO;                       // matcher.cpp:196-197
    pslam_camera cam = cameraFrom(cameraMatrix, &distCoeffs);
    pslam_ransac_params a = toAbi(rp, cam.fx, cam.fy, cam.cx, cam.cy);
    std::vector<float> prevXY(2 * (size_t)n), curXY(2 * (size_t)n), err((size_t)n), und(2 * (size_t)n), xyz(3 * (size_t)n);
    std::vector<double> dd((size_t)n);
    std::vector<unsigned char> status((size_t)n);
    std::vector<int> kept((size_t)n), inl((size_t)n);
    for (int i = 0; i < n; ++i) { prevXY[2 * i] = prevFeaturesDistorted[(size_t)i].x; prevXY[2 * i + 1] = prevFeaturesDistorted[(size_t)i].y; }
    const bool resident = reuseTracked_ && lastTrackedData_ == prevRgbImage.data && lastTrackedRows_ == prevRgbImage.rows &&
                          lastTrackedCols_ == prevRgbImage.cols && lastTrackedStep_ == prevRowBytes &&
                          lastTrackedCh_ == prevRgbImage.channels() && lastTrackedLevels_ >= tp.maxLevels;
    int nKept = 0;
    pslam_frame_result res;
    const int r = pslam_klt_frame(c, resident ? nullptr : prevRgbImage.data, rgbImage.data, rgbImage.cols, rgbImage.rows, rowBytes,
                                  rgbImage.channels(), prevXY.data(), prevFeatures3D[0].data(), curXY.data(), n, tp.winSize,
                                  tp.maxLevels, 3, tp.maxIter, tp.eps, tp.trackingErrorType > 0 ? PSLAM_KLT_GET_MIN_EIGENVALS : 0,
                                  tp.trackingMinEigThreshold, tp.trackingErrorThreshold, tp.minimalReprojDistanceNewTrackingFeatures,
                                  depthImage.ptr<uint16_t>(0), stride, &cam, distCoeffs.empty() ? 0 : 1, depthImageScale, &a, seed_,
                                  numHyp_, status.data(), err.data(), kept.data(), &nKept, und.data(), xyz.data(), dd.data(),
                                  inl.data(), &res);
    if (r != PSLAM_OK) { logError(c, "trackKLT", r); lastTrackedData_ = nullptr; return 0.0; }
    lastTrackedData_ = rgbImage.data; lastTrackedRows_ = rgbImage.rows; lastTrackedCols_ = rgbImage.cols; lastTrackedStep_ = rowBytes;
    lastTrackedCh_ = rgbImage.channels(); lastTrackedLevels_ = tp.maxLevels;
    distortedFeatures2D.resize((size_t)nKept); undistortedFeatures2D.resize((size_t)nKept); features3D.resize((size_t)nKept);
    keyPoints.resize((size_t)nKept); detDists.resize((size_t)nKept); matches.reserve((size_t)nKept);
    for (int j = 0; j < nKept; ++j) {
        const int i = kept[(size_t)j];
        distortedFeatures2D[(size_t)j] = cv::Point2f(curXY[2 * i], curXY[2 * i + 1]);
        undistortedFeatures2D[(size_t)j] = cv::Point2f(und[2 * j], und[2 * j + 1]);
        features3D[(size_t)j] = Eigen::Vector3f(xyz[3 * j], xyz[3 * j + 1], xyz[3 * j + 2]);
        keyPoints[(size_t)j] = prevKeyPoints[(size_t)i];
        keyPoints[(size_t)j].pt = distortedFeatures2D[(size_t)j];
        detDists[(size_t)j] = prevDetDists[(size_t)i];          // performTracking carries the detection distance over (:243)
        matches.push_back(cv::DMatch(i, j, 0));
    }
    for (int k = 0; k < res.n_inliers; ++k) inlierMatches.push_back(matches[(size_t)inl[k]]);
    std::memcpy(estimatedTransformation.data(), res.T, sizeof(res.T));
    return nKept > 0 ? res.inlier_ratio : 0.0;
}

double MatcherB200::matchFeatureLoopClosureCore(cv::Mat descriptors0, const std::vector<Eigen::Vector3f>& points3D0,
                                                cv::Mat descriptors1, const std::vector<Eigen::Vector3f>& points3D1,
                                                const RANSAC::parameters& ransacParams, cv::Mat cameraMatrix,
                                                std::vector<std::pair<int, int>>& pairedFeatures,
                                                Eigen::Matrix4f& estimatedTransformation) {
    const int n0 = descriptors0.rows, n1 = descriptors1.rows;
    if (n0 < 10 || n1 < 10) {   // matcher.cpp:830-834
        std::cout << "Too few features :(" << std::endl;
        return 0;
    }
    std::vector<uint8_t> t0, t1;
    const uint8_t* d0 = contiguousBytes(descriptors0, 32, t0);
    const uint8_t* d1 = contiguousBytes(descriptors1, 32, t1);
    RANSAC::parameters rp = ransacParams;
    rp.errorVersion = rp.errorVersionMap;   // matcher.cpp:843-844
    float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f;
    if (!cameraMatrix.empty()) {
        fx = cameraMatrix.at<float>(0, 0); fy = cameraMatrix.at<float>(1, 1);
        cx = cameraMatrix.at<float>(0, 2); cy = cameraMatrix.at<float>(1, 2);
    }
    pslam_ransac_params a = toAbi(rp, fx, fy, cx, cy);
    const int cap = std::min(n0, n1);
    std::vector<int> mq((size_t)cap), mt((size_t)cap), inl((size_t)cap);
    std::vector<float> md((size_t)cap);
    pslam_frame_result res;
    pslam_ctx* c = dev_.ctx();
    const int r = c ? pslam_loop_closure_pair(c, d0, points3D0[0].data(), n0, d1, points3D1[0].data(), n1, &a, seed_, numHyp_,
                                              mq.data(), mt.data(), md.data(), inl.data(), &res)
                    : PSLAM_ERR_NO_DEVICE;
    estimatedTransformation = Eigen::Matrix4f::Identity();
    pairedFeatures.clear();
    if (r != PSLAM_OK) { logError(c, "matchFeatureLoopClosure", r); return -1.0; }
    if (res.n_matches <= 0) return -1.0;   // matcher.cpp:838-839
    std::memcpy(estimatedTransformation.data(), res.T, sizeof(res.T));
    for (int k = 0; k < res.n_inliers; ++k) pairedFeatures.push_back(std::make_pair(mq[(size_t)inl[k]], mt[(size_t)inl[k]]));
    return res.inlier_ratio;
}

// ---- Kabsch ----------------------------------------------------------------------------------------
Mat34& KabschEst::computeTransformation(const Eigen::MatrixXd& setA, const Eigen::MatrixXd& setB) {
    transformation.setIdentity();
    const int n = (int)setA.rows();
    if (n == 0) return transformation;   // kabschEst.cpp:28
    std::vector<double> A(3 * (size_t)n), B(3 * (size_t)n);   // Eigen is column-major: (r, c) -> row-major points
    for (int r = 0; r < n; ++r)
        for (int c = 0; c < 3; ++c) { A[3 * r + c] = setA(r, c); B[3 * r + c] = setB(r, c); }
    int off[2] = {0, n};
    double T[12];
    pslam_ctx* c = defaultDevice().ctx();
    const int r = c ? pslam_kabsch_batch(c, A.data(), B.data(), off, 1, T) : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) { logError(c, "KabschEst::computeTransformation", r); return transformation; }
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 3; ++row) transformation.m[4 * col + row] = T[3 * col + row];
    return transformation;
}

// ---- host-side steps of Matcher::trackKLT (see pslam_adapter.h) ------------------------------------------------------
namespace tracking {
namespace {
// cv::norm(Point2f) / the explicit sqrt(u*u + v*v) of the reference: float differences, double arithmetic
inline bool close2D(const cv::Point2f& a, const cv::Point2f& b, double thr) {
    const double u = a.x - b.x, v = a.y - b.y;
    return std::sqrt(u * u + v * v) < thr;
}
inline bool close3D(const Eigen::Vector3f& a, const Eigen::Vector3f& b, double thr) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return std::sqrt(x * x + y * y + z * z) < thr;
}
// Uniform grid over points already seen; cells are 0.1 % wider than the threshold, so two points whose (float-rounded)
// distance is below the threshold always lie in neighbouring cells.  The grid only proposes candidates -- hash
// collisions add some, never lose one -- and the caller evaluates the exact predicate.
template <int DIM>
struct Grid {
    bool usable;
    double inv;
    std::unordered_map<long long, std::vector<int>> cells;
    explicit Grid(double thr) : usable(thr > 0 && std::isfinite(thr)), inv(usable ? 1.0 / (thr * 1.001) : 0.0) {}
    bool cell(const float* p, long long* c) const {
        for (int d = 0; d < DIM; ++d) {
            const double q = (double)p[d] * inv;
            if (!(std::fabs(q) < 1e15)) return false;     // NaN / inf / absurdly far: never close to anything
            c[d] = (long long)std::floor(q);
        }
        return true;
    }
    static long long key(const long long* c) {
        long long k = 0;
        for (int d = 0; d < DIM; ++d) k = k * 1000003LL + c[d];
        return k;
    }
    void insert(const float* p, int id) {
        long long c[DIM];
        if (cell(p, c)) cells[key(c)].push_back(id);
    }
    template <typename F>
    bool any(const float* p, F pred) const {                // pred(id) over all candidates until one holds
        long long c[DIM], n[DIM];
        if (!cell(p, c)) return false;
        const int total = DIM == 2 ? 9 : 27;
        for (int t = 0; t < total; ++t) {
            int r = t;
            for (int d = 0; d < DIM; ++d) { n[d] = c[d] + (r % 3) - 1; r /= 3; }
            auto it = cells.find(key(n));
            if (it == cells.end()) continue;
            for (int id : it->second) if (pred(id)) return true;
        }
        return false;
    }
};
template <typename T>
void permute(std::vector<T>& v, const std::vector<size_t>& order) {   // v := v[order]
    std::vector<T> out;
    out.reserve(order.size());
    for (size_t i : order) out.push_back(v[i]);
    v.swap(out);
}
template <typename T>
void compact(std::vector<T>& v, const std::vector<char>& gone) {
    size_t o = 0;
    for (size_t i = 0; i < v.size(); ++i) if (!gone[i]) v[o++] = v[i];
    v.resize(o);
}
}  // namespace

std::set<int> removeTooCloseFeatures(std::vector<cv::Point2f>& distortedFeatures2D, std::vector<cv::Point2f>& undistortedFeatures2D,
                                     std::vector<Eigen::Vector3f>& features3D, std::vector<cv::KeyPoint>& keyPoints,
                                     std::vector<double>& detDists, std::vector<cv::DMatch>& matches, double minEuclid,
                                     double minReproj, bool bruteForce) {
    std::set<int> featuresToRemove;
    const size_t n = features3D.size();
    if (undistortedFeatures2D.size() != n || distortedFeatures2D.size() != n || keyPoints.size() != n || detDists.size() != n) {
        std::cerr << "putslam_b200: removeTooCloseFeatures: vectors differ in size" << std::endl;
        return featuresToRemove;
    }
    std::vector<char> gone(n, 0);
    // the grids need finite thresholds and coordinates; anything else (an infinite threshold, NaN / inf / absurdly far
    // positions) takes the reference's plain double loop
    Grid<2> g2(minReproj);
    Grid<3> g3(minEuclid);
    bool gridOK = !bruteForce && (g2.usable || !(minReproj > 0)) && (g3.usable || !(minEuclid > 0));
    for (size_t j = 0; gridOK && j < n; ++j) {
        const float p2[2] = {undistortedFeatures2D[j].x, undistortedFeatures2D[j].y};
        const float p3[3] = {features3D[j][0], features3D[j][1], features3D[j][2]};
        long long c[3];
        gridOK = (!g2.usable || g2.cell(p2, c)) && (!g3.usable || g3.cell(p3, c));
    }
    if (!gridOK) {
        for (size_t i = 0; i < n; ++i)
            for (size_t j = i + 1; j < n; ++j)
                if (close3D(features3D[i], features3D[j], minEuclid) || close2D(undistortedFeatures2D[i], undistortedFeatures2D[j], minReproj))
                    gone[j] = 1;
    } else {
        for (size_t j = 0; j < n; ++j) {
            const float p2[2] = {undistortedFeatures2D[j].x, undistortedFeatures2D[j].y};
            const float p3[3] = {features3D[j][0], features3D[j][1], features3D[j][2]};
            // every earlier feature counts, removed or not (the reference's loops do not skip removed ones)
            if (g2.usable && g2.any(p2, [&](int i) { return close2D(undistortedFeatures2D[(size_t)i], undistortedFeatures2D[j], minReproj); })) gone[j] = 1;
            else if (g3.usable && g3.any(p3, [&](int i) { return close3D(features3D[(size_t)i], features3D[j], minEuclid); })) gone[j] = 1;
            if (g2.usable) g2.insert(p2, (int)j);
            if (g3.usable) g3.insert(p3, (int)j);
        }
    }
    for (size_t j = 0; j < n; ++j) if (gone[j]) featuresToRemove.insert((int)j);
    compact(distortedFeatures2D, gone); compact(undistortedFeatures2D, gone); compact(features3D, gone);
    compact(keyPoints, gone); compact(detDists, gone);
    matches.erase(std::remove_if(matches.begin(), matches.end(),
                                 [&](const cv::DMatch& o) { return featuresToRemove.find(o.trainIdx) != featuresToRemove.end(); }),
                  matches.end());
    return featuresToRemove;
}

void mergeTrackedFeatures(std::vector<cv::Point2f>& undistortedFeatures2D, const std::vector<cv::Point2f>& featuresSandBoxUndistorted,
                          std::vector<cv::Point2f>& distortedFeatures2D, const std::vector<cv::Point2f>& featuresSandBoxDistorted,
                          std::vector<Eigen::Vector3f>& features3D, const std::vector<Eigen::Vector3f>& features3DSandBox,
                          std::vector<cv::KeyPoint>& keyPoints, const std::vector<cv::KeyPoint>& keyPointsSandBox,
                          std::vector<double>& detDists, const std::vector<double>& detDistsSandBox, double minReproj,
                          bool bruteForce) {
    const size_t m = featuresSandBoxUndistorted.size();
    if (featuresSandBoxDistorted.size() != m || features3DSandBox.size() != m || keyPointsSandBox.size() != m || detDistsSandBox.size() != m) {
        std::cerr << "putslam_b200: mergeTrackedFeatures: vectors differ in size" << std::endl;
        return;
    }
    Grid<2> g(minReproj);
    bool grid = !bruteForce && g.usable;
    long long cc[2];
    for (size_t j = 0; grid && j < undistortedFeatures2D.size(); ++j) {
        const float p[2] = {undistortedFeatures2D[j].x, undistortedFeatures2D[j].y};
        grid = g.cell(p, cc);
    }
    for (size_t i = 0; grid && i < m; ++i) {
        const float p[2] = {featuresSandBoxUndistorted[i].x, featuresSandBoxUndistorted[i].y};
        grid = g.cell(p, cc);
    }
    if (grid)
        for (size_t j = 0; j < undistortedFeatures2D.size(); ++j) {
            const float p[2] = {undistortedFeatures2D[j].x, undistortedFeatures2D[j].y};
            g.insert(p, (int)j);
        }
    for (size_t i = 0; i < m; ++i) {
        const cv::Point2f& c = featuresSandBoxUndistorted[i];
        const float p[2] = {c.x, c.y};
        bool addFeature = true;
        if (grid) {
            addFeature = !g.any(p, [&](int j) { return close2D(c, undistortedFeatures2D[(size_t)j], minReproj); });
        } else if (minReproj > 0 || minReproj != minReproj) {   // (a non-positive threshold rejects nothing)
            for (size_t j = 0; j < undistortedFeatures2D.size(); ++j)
                if (close2D(c, undistortedFeatures2D[j], minReproj)) { addFeature = false; break; }
        }
        if (addFeature) {
            if (grid) g.insert(p, (int)undistortedFeatures2D.size());
            undistortedFeatures2D.push_back(c);
            distortedFeatures2D.push_back(featuresSandBoxDistorted[i]);
            features3D.push_back(features3DSandBox[i]);
            keyPoints.push_back(keyPointsSandBox[i]);
            detDists.push_back(detDistsSandBox[i]);
        }
    }
}

std::vector<cv::KeyPoint> predictDescriptionLevels(std::vector<cv::Point2f>& distortedFeatures2D,
                                                   std::vector<cv::Point2f>& undistortedFeatures2D,
                                                   std::vector<Eigen::Vector3f>& features3D, std::vector<cv::KeyPoint>& keyPoints,
                                                   std::vector<double>& detDists) {
    const int nLevels = 8;                                   // Matcher::nLevels (matcher.h:27)
    std::vector<cv::KeyPoint> descKeyPoints = keyPoints;
    const size_t n = keyPoints.size();
    if (distortedFeatures2D.size() != n || undistortedFeatures2D.size() != n || features3D.size() != n || detDists.size() != n) {
        std::cerr << "putslam_b200: predictDescriptionLevels: vectors differ in size" << std::endl;
        return descKeyPoints;
    }
    for (size_t i = 0; i < n; ++i) {
        // std::sqrt of a float expression: float arithmetic, float square root, then widened (matcher.cpp:287-290)
        const float x = features3D[i][0], y = features3D[i][1], z = features3D[i][2];
        const double curDist = (double)std::sqrt(x * x + y * y + z * z);
        descKeyPoints[i].octave = predictedLevel(descKeyPoints[i].octave, detDists[i], curDist);
    }
    std::vector<size_t> order;
    order.reserve(n);
    for (int l = 0; l < nLevels; ++l)
        for (size_t i = 0; i < n; ++i) if (descKeyPoints[i].octave == l) order.push_back(i);
    permute(distortedFeatures2D, order); permute(undistortedFeatures2D, order); permute(features3D, order);
    permute(keyPoints, order); permute(detDists, order);
    return descKeyPoints;
}

void dropUndescribed(const std::vector<cv::KeyPoint>& descKeyPoints, std::vector<cv::Point2f>& distortedFeatures2D,
                     std::vector<cv::Point2f>& undistortedFeatures2D, std::vector<Eigen::Vector3f>& features3D,
                     std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& detDists) {
    const size_t n = keyPoints.size();
    if (descKeyPoints.size() == n) return;                   // matcher.cpp:342
    if (distortedFeatures2D.size() != n || undistortedFeatures2D.size() != n || features3D.size() != n || detDists.size() != n) {
        std::cerr << "putslam_b200: dropUndescribed: vectors differ in size" << std::endl;
        return;
    }
    std::vector<char> gone(n, 1);
    for (size_t i = 0, j = 0; i < n && j < descKeyPoints.size(); ++i)
        if (close2D(keyPoints[i].pt, descKeyPoints[j].pt, 0.0001)) { gone[i] = 0; ++j; }
    compact(distortedFeatures2D, gone); compact(undistortedFeatures2D, gone); compact(features3D, gone);
    compact(keyPoints, gone); compact(detDists, gone);
}
}  // namespace tracking

const Mat66& TransformEst::uncertaintyImpl(const Eigen::MatrixXd& setA, std::vector<Mat33>& ua, const Eigen::MatrixXd& setB,
                                           std::vector<Mat33>& ub, Mat34& T, int parametrization) {
    uncertainty.setZero();
    const int n = (int)setA.rows();
    pslam_ctx* c = defaultDevice().ctx();
    if (n == 0 || (int)setB.rows() != n || (int)ua.size() != n || (int)ub.size() != n) {
        logError(c, "TransformEst::computeUncertainty", PSLAM_ERR_ARG);
        return uncertainty;
    }
    std::vector<double> A(3 * (size_t)n), B(3 * (size_t)n), CA(9 * (size_t)n), CB(9 * (size_t)n);
    for (int r = 0; r < n; ++r) {
        for (int k = 0; k < 3; ++k) { A[3 * r + k] = setA(r, k); B[3 * r + k] = setB(r, k); }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { CA[9 * r + 3 * i + j] = ua[(size_t)r](i, j); CB[9 * r + 3 * i + j] = ub[(size_t)r](i, j); }
    }
    int off[2] = {0, n}, ok = 0;
    double T12[12];
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 3; ++row) T12[3 * col + row] = T.m[4 * col + row];
    const int r = c ? pslam_transform_uncertainty_batch(c, A.data(), B.data(), CA.data(), CB.data(), off, T12, 1, parametrization,
                                                        uncertainty.m, &ok)
                    : PSLAM_ERR_NO_DEVICE;
    if (r != PSLAM_OK) { logError(c, "TransformEst::computeUncertainty", r); uncertainty.setZero(); }
    else if (!ok) std::cerr << "putslam_b200: TransformEst::computeUncertainty: singular Hessian" << std::endl;
    return uncertainty;
}
const Mat66& TransformEst::computeUncertainty(const Eigen::MatrixXd& setA, std::vector<Mat33>& setAUncertainty,
                                              const Eigen::MatrixXd& setB, std::vector<Mat33>& setBUncertainty, Mat34& T) {
    return uncertaintyImpl(setA, setAUncertainty, setB, setBUncertainty, T, PSLAM_UNCERTAINTY_EULER);
}
const Mat66& TransformEst::computeUncertaintyG2O(const Eigen::MatrixXd& setA, std::vector<Mat33>& setAUncertainty,
                                                 const Eigen::MatrixXd& setB, std::vector<Mat33>& setBUncertainty, Mat34& T) {
    return uncertaintyImpl(setA, setAUncertainty, setB, setBUncertainty, T, PSLAM_UNCERTAINTY_QUATERNION);
}

const Mat66& TransformEst::computeUncertaintyStrasdat(const Eigen::MatrixXd& setA, const Eigen::MatrixXd& setB, Mat34& T) {
    uncertainty.setZero();
    for (int i = 0; i < 6; ++i) uncertainty.m[7 * i] = 1.0;
    double depthAv = 0;
    for (long i = 0; i < (long)setA.rows(); ++i) {
        depthAv += std::sqrt(std::pow(setA(i, 0), 2.0) + std::pow(setA(i, 1), 2.0) + std::pow(setA(i, 2), 2.0));
        depthAv += std::sqrt(std::pow(setB(i, 0), 2.0) + std::pow(setB(i, 1), 2.0) + std::pow(setB(i, 2), 2.0));
    }
    depthAv /= 2 * (double)setA.rows();
    for (int k = 0; k < 3; ++k) uncertainty.m[7 * k] = std::pow(T(k, 3) / depthAv, 2.0);
    return uncertainty;
}

static std::unique_ptr<KabschEst> kabsch;
TransformEst* createKabschEstimator(void) {
    kabsch.reset(new KabschEst());
    return kabsch.get();
}

}  // namespace putslam_b200
