// pslam_adapter.h -- C++ host side of the drop-in: PUTSLAM's own hot-path interfaces, re-implemented
// on top of the C ABI (include/pslam_b200.h).  Same class / function names, argument meaning, ownership
// and failure conventions as the reference, inside namespace putslam_b200 so that both can be linked
// into one binary during a migration (INTEGRATION.md shows the three-line change per call site).
//
//   putslam_b200::RANSAC                  <->  class RANSAC               include/putslam/TransformEst/RANSAC.h:20-198
//   putslam_b200::MatcherB200::performMatching  <->  MatcherOpenCV::performMatching  src/Matcher/matcherOpenCV.cpp:198-206
//   putslam_b200::MatcherB200::matchXYZCore     <->  loop nest of Matcher::matchXYZ  src/Matcher/matcher.cpp:606-767
//   putslam_b200::RGBD::*                 <->  namespace RGBD             include/putslam/RGBD/RGBD.h:38-73
//   putslam_b200::KabschEst               <->  putslam::KabschEst         include/putslam/TransformEst/kabschEst.h:21-41
//
// Error convention: the reference has no error codes on this path; failure is "identity
// transform + no inliers" / "no matches".  Any C-ABI error is logged to std::cerr and mapped to exactly
// those values.  There is no CPU fallback.
#pragma once
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../include/pslam_b200.h"
#include "shim/pslam_shim_types.h"

namespace putslam_b200 {

// One device context per Matcher instance (tracking thread / loop-closure thread), created lazily.
class Device {
public:
    explicit Device(int device = 0);
    ~Device();
    pslam_ctx* ctx();
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
private:
    int device_;
    pslam_ctx* ctx_ = nullptr;
};
Device& defaultDevice();   // process-wide context used by the free functions and by RANSAC objects

namespace RGBD {
// include/putslam/RGBD/RGBD.h:38-73 -- identical signatures (arguments by value like the reference)
std::vector<cv::Point2f> removeImageDistortion(std::vector<cv::Point2f>& features, cv::Mat cameraMatrix, cv::Mat distCoeffs);
std::vector<cv::Point2f> removeImageDistortion(std::vector<cv::KeyPoint>& features, cv::Mat cameraMatrix, cv::Mat distCoeffs);
std::vector<Eigen::Vector3f> keypoints2Dto3D(std::vector<cv::Point2f> undistortedFeatures2D, cv::Mat depthImage,
                                             cv::Mat cameraMatrix, double depthImageScale, int startingID = 0);
}  // namespace RGBD

class RANSAC {
public:
    enum ERROR_VERSION { EUCLIDEAN_ERROR, REPROJECTION_ERROR, EUCLIDEAN_AND_REPROJECTION_ERROR, MAHALANOBIS_ERROR, ADAPTIVE_ERROR };
    struct parameters {   // field for field RANSAC::parameters (RANSAC.h:23-31)
        int verbose;
        int errorVersion, errorVersionVO, errorVersionMap;
        double inlierThresholdEuclidean, inlierThresholdReprojection, inlierThresholdMahalanobis;
        double minimalInlierRatioThreshold;
        int minimalNumberOfMatches;
        int usedPairs;
        int iterationCount;
    };
    RANSAC(parameters RANSACParameters, cv::Mat cameraMatrix = cv::Mat());

    Eigen::Matrix4f estimateTransformation(std::vector<Eigen::Vector3f> prevFeatures, std::vector<Eigen::Vector3f> features,
                                           std::vector<cv::DMatch> matches, std::vector<cv::DMatch>& inlierMatches);
    static double pointInlierRatio(std::vector<cv::DMatch>& inlierMatches, std::vector<cv::DMatch>& allMatches);

    // Additions (defaults keep the reference's behaviour): the reference seeds rand() with time(0) per
    // object (RANSAC.cpp:13); here the seed is explicit and the sample stream is counter-based.
    void setSeed(uint64_t seed) { seed_ = seed; }
    void setFixedHypotheses(int n) { numHyp_ = n; }   // 0 = the reference's adaptive bound (487, shrinking)
    int hypothesesUsed() const { return hypUsed_; }
    double bestInlierRatio() const { return bestRatio_; }
private:
    parameters RANSACParams;
    float fx_ = 517.3f, fy_ = 516.5f, cx_ = 318.6f, cy_ = 255.3f;
    uint64_t seed_;
    int numHyp_ = 0, hypUsed_ = 0;
    double bestRatio_ = 0.0;
};

// Matching half of the Matcher facade.  In a PUTSLAM build this is used as
//   class MatcherB200 : public MatcherOpenCV { std::vector<cv::DMatch> performMatching(cv::Mat a, cv::Mat b) override
//       { return core.performMatching(a, b); } putslam_b200::MatcherB200 core; };
class MatcherB200 {
public:
    explicit MatcherB200(int device = 0) : dev_(device) {}
    // MatcherOpenCV::performMatching (src/Matcher/matcherOpenCV.cpp:198-206), NORM_HAMMING + crossCheck
    std::vector<cv::DMatch> performMatching(cv::Mat prevDescriptors, cv::Mat descriptors);
    // MatcherOpenCV::detectFeatures for detector == "ORB" (src/Matcher/matcherOpenCV.cpp:118-176): RGB -> gray, the image
    // cut into gridCols x gridRows cells, cv::ORB::create()->detect on every cell (pslam_orb_detect), per cell the best
    // maximalTrackedFeatures * 3 / (gridCols * gridRows) by response, all cells merged, sorted by response and cut to
    // maximalTrackedFeatures.  The sorts are std::sort with the reference's comparator (matcherOpenCV.h:84-87).
    std::vector<cv::KeyPoint> detectFeatures(cv::Mat rgbImage, int gridCols = 1, int gridRows = 1,
                                             int maximalTrackedFeatures = 500);
    // the same wrapper for detector == "FAST" (matcherOpenCV.cpp:60-61: cv::FastFeatureDetector::create(), threshold 10,
    // non-maximum suppression, 9_16): pslam_fast_detect per grid cell
    std::vector<cv::KeyPoint> detectFeaturesFAST(cv::Mat rgbImage, int gridCols = 1, int gridRows = 1,
                                                 int maximalTrackedFeatures = 500);
    // MatcherOpenCV::describeFeatures for descriptor == "ORB" (src/Matcher/matcherOpenCV.cpp:181-195):
    // cv::ORB::create()->compute(rgbImage, features, descriptors).  rgbImage: CV_8UC3 (converted like ORB does, with
    // COLOR_BGR2GRAY on the stored channel order) or CV_8UC1.  `features` is filtered and reordered exactly as
    // cv::ORB::compute does it (keypoints within 31 px of the border dropped, the rest regrouped by octave); returns
    // one 32-byte row per remaining feature.
    cv::Mat describeFeatures(cv::Mat rgbImage, std::vector<cv::KeyPoint>& features);
    // MatcherOpenCV::performTracking (src/Matcher/matcherOpenCV.cpp:209-300; virtual, matcher.h:415-422) in one
    // submission (pslam_klt_perform_tracking): pyramidal Lucas-Kanade of prevFeatures from prevImg to img, err threshold,
    // pairwise too-close removal; features / keyPoints / detDists come back compacted to the survivors, the result is
    // DMatch(index in prevFeatures, index in features, 0) like the reference's.  With useInitialFlow `features` must hold
    // one guess per previous feature on entry (OpenCV's requirement).  On a device error: no matches, empty outputs,
    // message on std::cerr.
    struct TrackingParams {   // MatcherParameters::OpenCVParams fields (matcher.h:44-61), shipped defaults
        int winSize = 7, maxLevels = 3, maxIter = 30;     // resources/putslammatcherOpenCVParameters.xml:64
        double eps = 0.01;
        int useInitialFlow = 0, trackingErrorType = 0;
        double trackingErrorThreshold = 25.0, trackingMinEigThreshold = 0.0;
        double minimalReprojDistanceNewTrackingFeatures = 3.0;
    };
    void setTrackingParams(const TrackingParams& p) { tracking_ = p; }
    std::vector<cv::DMatch> performTracking(cv::Mat prevImg, cv::Mat img, std::vector<cv::Point2f>& prevFeatures,
                                            std::vector<cv::Point2f>& features, std::vector<cv::KeyPoint>& prevKeyPoints,
                                            std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& prevDetDists,
                                            std::vector<double>& detDists);
    // Opt-in: Matcher::trackKLT passes prevRgbImage = the frame it tracked INTO one call earlier (matcher.cpp:151-152).
    // When on, and prevImg is the very Mat (data pointer, size, step) the last performTracking received as img, the
    // previous frame is not uploaded again: its pyramid is still in HBM.  Same promise as setReuseDetectedFrame.
    void setReuseTrackedFrame(bool on) { reuseTracked_ = on; }
    // Opt-in: when describeFeatures is called with the very Mat (same data pointer, size, step) the last full-frame
    // detectFeatures saw, skip the second upload and describe on the frame already in HBM.  The caller promises not to
    // modify the pixels between the two calls (Matcher::match does not, src/Matcher/matcher.cpp:457-467).
    void setReuseDetectedFrame(bool on) { reuseFrame_ = on; }
    // knnMatch(k=2) + Lowe ratio (north_star extension): matches with d1 < ratio * d2
    std::vector<cv::DMatch> performMatchingRatio(cv::Mat prevDescriptors, cv::Mat descriptors, float ratio);

    struct MapSide {   // SoA view of std::vector<MapFeature> (one ExtendedDescriptor per feature, SURVEY B#13)
        std::vector<double> xyz;        // M x 3, MapFeature.position (double)
        cv::Mat descriptors;            // M x 32 CV_8U
        std::vector<int> octave;        // ExtendedDescriptor.octave
        std::vector<double> detDist;    // ExtendedDescriptor.detDist
    };
    // The loop nest + RANSAC of Matcher::matchXYZ (src/Matcher/matcher.cpp:617-767).  Returns what matchXYZ returns
    // (pointInlierRatio, or -1.0 when there are no matches); fills matches / inlierMatches / estimatedTransformation.
    double matchXYZCore(const MapSide& map, cv::Mat currentPoseDescriptors, std::vector<Eigen::Vector3f>& currentPoseFeatures3D,
                        std::vector<cv::KeyPoint>& currentPoseKeyPoints, std::vector<double>& currentPoseDetDists,
                        double matchingXYZSphereRadius, double matchingXYZacceptRatioOfBestMatch, int computationNumber,
                        const RANSAC::parameters& ransacParams, cv::Mat cameraMatrix, Eigen::Matrix4f& estimatedTransformation,
                        std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& inlierMatches, bool xorDistance = false);
    // ---- resident map (pslam_map_* / pslam_frame_to_resident_map): the covisible features stay in HBM ----
    // uploadMapFeatures stores slots [first, first + M); viewAxis = M x 3 float, third column of the rotation of the
    // view that holds each feature's descriptor (what FeaturesMap::findNearestFrame compares, featuresMap.cpp:528-563).
    bool uploadMapFeatures(int first, const MapSide& features, const std::vector<float>& viewAxis);
    // positions only, e.g. after FeaturesMap::updateMap moved the features (xyz = count x 3)
    bool updateMapPositions(int first, const std::vector<double>& xyz);
    bool truncateMap(int nFeatures);
    int mapSize();
    struct MapFilter {          // getAndFilterFeaturesFromMap's parameters (src/PUTSLAM/PUTSLAM.cpp:624-674)
        double fx = 517.3, fy = 516.5, cx = 318.6, cy = 255.3, imageW = 640, imageH = 480;   // DepthSensorModel
        double maxAngleBetweenFrames = 0.6;                                                     // matcherParameters
        double maxZ = 5.0;                                                                      // PUTSLAM.cpp:662
    };
    // getAndFilterFeaturesFromMap's numeric part + matchXYZ in one submission.  cameraPose = Mat34::matrix() data
    // (column-major 4x4, camera -> global).  keptFeatures lists the slots that passed the filters; matches'
    // queryIdx index that list, as the reference's index the filtered mapFeatures vector.
    double matchXYZResident(const double cameraPose[16], const MapFilter& filter, cv::Mat currentPoseDescriptors,
                            std::vector<Eigen::Vector3f>& currentPoseFeatures3D, std::vector<cv::KeyPoint>& currentPoseKeyPoints,
                            std::vector<double>& currentPoseDetDists, double matchingXYZSphereRadius,
                            double matchingXYZacceptRatioOfBestMatch, int computationNumber,
                            const RANSAC::parameters& ransacParams, cv::Mat cameraMatrix,
                            Eigen::Matrix4f& estimatedTransformation, std::vector<int>& keptFeatures,
                            std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& inlierMatches, bool xorDistance = false);
    // The numeric core of Matcher::match (src/Matcher/matcher.cpp:452-516, lines 470-496) in one device submission:
    // performMatching(prevDescriptors, descriptors) -> removeImageDistortion + keypoints2Dto3D on the current keypoints
    // -> RANSAC(prevFeatures3D, features3D, matches).  Returns pointInlierRatio (matcher.cpp:515).  Outputs are what
    // match() keeps for the next frame (undistortedFeatures2D, features3D) plus matches / inliers / transform.
    double matchCore(cv::Mat prevDescriptors, const std::vector<Eigen::Vector3f>& prevFeatures3D, cv::Mat descriptors,
                     const std::vector<cv::KeyPoint>& keyPoints, cv::Mat depthImage, double depthImageScale,
                     cv::Mat cameraMatrix, cv::Mat distCoeffs, const RANSAC::parameters& ransacParams,
                     std::vector<cv::Point2f>& undistortedFeatures2D, std::vector<Eigen::Vector3f>& features3D,
                     std::vector<cv::DMatch>& matches, std::vector<cv::DMatch>& inlierMatches,
                     Eigen::Matrix4f& estimatedTransformation);
    // The numeric core of Matcher::trackKLT (src/Matcher/matcher.cpp:133-207, lines 151-207; removeTooCloseFeatures off as
    // shipped) in one device submission (pslam_klt_frame): performTracking(prevRgbImage, rgbImage, prevFeaturesDistorted,
    // ...) -> removeImageDistortion + keypoints2Dto3D on the survivors -> RANSAC(prevFeatures3D, features3D, matches),
    // errorVersion = errorVersionVO (:196-197).  Outputs are the vectors trackKLT works on afterwards: the compacted
    // distorted / undistorted positions, 3-D points, key points and detDists, the DMatch(i, j, 0) list, the inliers and
    // the transformation.  Returns what trackKLT returns: pointInlierRatio(inlierMatches, matches), 0 without matches.
    // Uses the parameters of setTrackingParams and the frame reuse of setReuseTrackedFrame.
    double trackKLTCore(cv::Mat prevRgbImage, cv::Mat rgbImage, const std::vector<cv::Point2f>& prevFeaturesDistorted,
                        const std::vector<Eigen::Vector3f>& prevFeatures3D, const std::vector<cv::KeyPoint>& prevKeyPoints,
                        const std::vector<double>& prevDetDists, cv::Mat depthImage, double depthImageScale, cv::Mat cameraMatrix,
                        cv::Mat distCoeffs, const RANSAC::parameters& ransacParams, std::vector<cv::Point2f>& distortedFeatures2D,
                        std::vector<cv::Point2f>& undistortedFeatures2D, std::vector<Eigen::Vector3f>& features3D,
                        std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& detDists, std::vector<cv::DMatch>& matches,
                        std::vector<cv::DMatch>& inlierMatches, Eigen::Matrix4f& estimatedTransformation);
    // Matcher::matchFeatureLoopClosure (src/Matcher/matcher.cpp:802-861) after its MapFeature gathering loop (:809-827):
    // descriptors / 3-D points of the two frames -> paired feature indices, transform, and the value it returns
    // (0 for fewer than 10 features, -1.0 for no matches, else pointInlierRatio).
    double matchFeatureLoopClosureCore(cv::Mat descriptors0, const std::vector<Eigen::Vector3f>& points3D0, cv::Mat descriptors1,
                                       const std::vector<Eigen::Vector3f>& points3D1, const RANSAC::parameters& ransacParams,
                                       cv::Mat cameraMatrix, std::vector<std::pair<int, int>>& pairedFeatures,
                                       Eigen::Matrix4f& estimatedTransformation);
    void setSeed(uint64_t s) { seed_ = s; }
    void setFixedHypotheses(int n) { numHyp_ = n; }
    // true: predicted pyramid levels computed here with the host libm exactly like matcher.cpp:639-651,682-692 and handed
    // to the device; false (default): computed on the device (pslam_frame_to_map_features), same values, ~0.2 ms less
    // host time per frame
    void setHostLevels(bool on) { hostLevels_ = on; }
    // Page-lock the buffers of a map side that is handed to matchXYZCore frame after frame (pslam_host_register): its
    // descriptors, positions, octaves and detDists then go to the device straight from the MapSide, without the staging
    // copy (0.34 MB per frame at 5000 features).  Call again after the MapSide changed size; unpinMapSide before it dies.
    bool pinMapSide(const MapSide& map);
    void unpinMapSide();
private:
    std::vector<const void*> pinnedMap_;
    Device dev_;
    uint64_t seed_ = 0x5eed5eedULL;
    int numHyp_ = 0;
    bool hostLevels_ = false;
    std::vector<cv::KeyPoint> detectGrid(cv::Mat rgbImage, int gridCols, int gridRows, int maximalTrackedFeatures, bool fast);
    bool reuseFrame_ = false;
    TrackingParams tracking_;
    bool reuseTracked_ = false;
    const unsigned char* lastTrackedData_ = nullptr;   // img of the last performTracking
    int lastTrackedRows_ = 0, lastTrackedCols_ = 0, lastTrackedStep_ = 0, lastTrackedCh_ = 0, lastTrackedLevels_ = 0;
    const unsigned char* lastFrameData_ = nullptr;   // frame of the last 1 x 1 detectFeatures
    int lastFrameRows_ = 0, lastFrameCols_ = 0, lastFrameStep_ = 0, lastFrameCh_ = 0;
};

// DBScan (include/putslam/Matcher/dbscan.h:14-41, src/Matcher/dbscan.cpp): the de-clustering pass the reference runs on
// every detected keypoint list (Matcher::detectInitFeatures / match / matchXYZ, src/Matcher/matcher.cpp:24-26,459-461,
// 561-563) -- of every group of keypoints chained together by distances below eps, only the first featuresFromCluster (in
// list order) survive.  Same constructor and run() as the reference class.  It is a sequential graph walk over a few
// hundred keypoints (the cluster a border keypoint joins depends on the visiting order), so it stays on the host; the
// pairwise distances are evaluated on demand instead of through the reference's N x N matrix.
class DBScan {
public:
    DBScan(double eps = 10, int minPts = 2, int featuresFromCluster = 1) : eps_(eps), minPts_(minPts), perCluster_(featuresFromCluster) {}
    void run(std::vector<cv::KeyPoint>& clusteringSet);
    // cluster label per input keypoint of the last run(): -1 noise, > 0 cluster id (in order of creation)
    const std::vector<int>& labels() const { return label_; }
private:
    double eps_;
    int minPts_, perCluster_;
    std::vector<int> label_;
};

// Host-side steps of Matcher::trackKLT around the device calls (src/Matcher/matcher.cpp:96-131,262-322,886-974).  They are
// sequential, order-dependent list edits on a few hundred features and stay on the host like the reference's; the O(N^2)
// scans are replaced by uniform-grid neighbour queries that evaluate the reference's exact predicate on the candidates,
// so the results are identical (tests/test_abi_cpu.py compares them with the brute-force form and a numpy restatement).
namespace tracking {
// Matcher::removeTooCloseFeatures (matcher.cpp:886-974): feature j goes when some i < j (removed or not) is closer than
// minimalEuclidDistance in 3-D or than minimalReprojDistance in the undistorted image; the five vectors are compacted,
// matches whose trainIdx was removed are erased (trainIdx is NOT renumbered -- as in the reference).  Returns the set.
std::set<int> removeTooCloseFeatures(std::vector<cv::Point2f>& distortedFeatures2D, std::vector<cv::Point2f>& undistortedFeatures2D,
                                     std::vector<Eigen::Vector3f>& features3D, std::vector<cv::KeyPoint>& keyPoints,
                                     std::vector<double>& detDists, std::vector<cv::DMatch>& matches,
                                     double minimalEuclidDistanceNewTrackingFeatures, double minimalReprojDistanceNewTrackingFeatures,
                                     bool bruteForce = false);
// Matcher::mergeTrackedFeatures (matcher.cpp:96-131): every newly detected feature that is not closer than
// minimalReprojDistance (undistorted image) to a feature already in the list -- tracked or just added -- is appended.
void mergeTrackedFeatures(std::vector<cv::Point2f>& undistortedFeatures2D, const std::vector<cv::Point2f>& featuresSandBoxUndistorted,
                          std::vector<cv::Point2f>& distortedFeatures2D, const std::vector<cv::Point2f>& featuresSandBoxDistorted,
                          std::vector<Eigen::Vector3f>& features3D, const std::vector<Eigen::Vector3f>& features3DSandBox,
                          std::vector<cv::KeyPoint>& keyPoints, const std::vector<cv::KeyPoint>& keyPointsSandBox,
                          std::vector<double>& detDists, const std::vector<double>& detDistsSandBox,
                          double minimalReprojDistanceNewTrackingFeatures, bool bruteForce = false);
// matcher.cpp:281-340 up to the describeFeatures call: descKeyPoints = keyPoints with octave := the pyramid level predicted
// from the detection distance and the current distance (host libm, like the reference), then all vectors regrouped by that
// level, stably.  Returns descKeyPoints in the ORIGINAL order, as the reference holds it at that point: describeFeatures
// (cv::ORB::compute) regroups it by octave itself, which is exactly the order the five vectors were just given.
std::vector<cv::KeyPoint> predictDescriptionLevels(std::vector<cv::Point2f>& distortedFeatures2D,
                                                   std::vector<cv::Point2f>& undistortedFeatures2D,
                                                   std::vector<Eigen::Vector3f>& features3D, std::vector<cv::KeyPoint>& keyPoints,
                                                   std::vector<double>& detDists);
// matcher.cpp:341-380, the "unlucky case" after describeFeatures: cv::ORB::compute drops key points too close to the border,
// so descKeyPoints may have come back shorter; the five vectors keep exactly the features whose position still appears
// in descKeyPoints (sequential two-pointer walk, |pt - pt| < 0.0001 as in the reference).  No-op when nothing was dropped.
void dropUndescribed(const std::vector<cv::KeyPoint>& descKeyPoints, std::vector<cv::Point2f>& distortedFeatures2D,
                     std::vector<cv::Point2f>& undistortedFeatures2D, std::vector<Eigen::Vector3f>& features3D,
                     std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& detDists);
}  // namespace tracking

// putslam::TransformEst / KabschEst (transformEst.h:16-26, kabschEst.h:21-41).  Mat34 is
// Eigen::Transform<double,3,Affine>; its 4x4 column-major matrix is exposed here as double[16].
struct Mat34 {
    double m[16];
    Mat34() { setIdentity(); }
    void setIdentity() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    double operator()(int r, int c) const { return m[4 * c + r]; }
};
struct Mat33 {   // Eigen::Matrix<double,3,3>, column-major
    double m[9];
    Mat33() { for (int i = 0; i < 9; ++i) m[i] = 0.0; }
    double& operator()(int r, int c) { return m[3 * c + r]; }
    double operator()(int r, int c) const { return m[3 * c + r]; }
};
struct Mat66 {   // Eigen::Matrix<double,6,6>, column-major
    double m[36];
    Mat66() { setZero(); }
    void setZero() { for (int i = 0; i < 36; ++i) m[i] = 0.0; }
    double operator()(int r, int c) const { return m[6 * c + r]; }
};
class TransformEst {
public:
    virtual const std::string& getName() const = 0;
    virtual Mat34& computeTransformation(const Eigen::MatrixXd& setA, const Eigen::MatrixXd& setB) = 0;
    virtual ~TransformEst() {}
    // transformEst.h:29-144: 6 x 6 covariance of (x, y, z, roll, pitch, yaw) of `transformation` (setA ~ R setB + t) from the
    // per-point covariances; returns a reference to the member, like the reference.  Zero matrix + message on std::cerr when
    // the device call fails or the Hessian is singular.
    virtual const Mat66& computeUncertainty(const Eigen::MatrixXd& setA, std::vector<Mat33>& setAUncertainty,
                                            const Eigen::MatrixXd& setB, std::vector<Mat33>& setBUncertainty, Mat34& transformation);
    // transformEst.h:147-272: the same over (x, y, z, qx, qy, qz)
    virtual const Mat66& computeUncertaintyG2O(const Eigen::MatrixXd& setA, std::vector<Mat33>& setAUncertainty,
                                               const Eigen::MatrixXd& setB, std::vector<Mat33>& setBUncertainty, Mat34& transformation);
    // transformEst.h:343-356 (demoKabsch.cpp:351,438): identity with (t / mean point distance)^2 on the translation
    // diagonal.  A 2n-term host sum -- there is nothing to offload.
    virtual const Mat66& computeUncertaintyStrasdat(const Eigen::MatrixXd& setA, const Eigen::MatrixXd& setB, Mat34& transformation);
protected:
    Mat34 transformation;
    Mat66 uncertainty;
private:
    const Mat66& uncertaintyImpl(const Eigen::MatrixXd& setA, std::vector<Mat33>& ua, const Eigen::MatrixXd& setB,
                                 std::vector<Mat33>& ub, Mat34& T, int parametrization);
};
class KabschEst : public TransformEst {
public:
    KabschEst() : name("Kabsch Estimator") {}
    const std::string& getName() const override { return name; }
    Mat34& computeTransformation(const Eigen::MatrixXd& setA, const Eigen::MatrixXd& setB) override;
private:
    const std::string name;
};
TransformEst* createKabschEstimator(void);   // kabschEst.cpp:70-73: singleton, a second call replaces the first

}  // namespace putslam_b200
