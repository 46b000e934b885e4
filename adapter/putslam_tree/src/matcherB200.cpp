/** @file matcherB200.cpp -- the file a PUTSLAM maintainer adds as src/Matcher/matcherB200.cpp (see Matcher/matcherB200.h) */
#include "Matcher/matcherB200.h"

using namespace putslam;

/// separate instances for the tracking matcher and the loop-closure matcher, a second create call replaces the first
/// one -- as matcherOpenCV.cpp:17-47 does with its two file-static unique_ptrs
MatcherB200::Ptr matcherClassB200;
MatcherB200::Ptr loopClosingMatcherClassB200;

std::vector<cv::KeyPoint> MatcherB200::detectFeatures(cv::Mat rgbImage) {
    const Matcher::parameters& o = matcherParameters.OpenCVParams;
    if (o.detector == "ORB")
        return core.detectFeatures(rgbImage, o.gridCols, o.gridRows, o.maximalTrackedFeatures);
    if (o.detector == "FAST")
        return core.detectFeaturesFAST(rgbImage, o.gridCols, o.gridRows, o.maximalTrackedFeatures);
    return MatcherOpenCV::detectFeatures(rgbImage);              // SURF / SIFT: not this path
}

cv::Mat MatcherB200::describeFeatures(cv::Mat rgbImage, std::vector<cv::KeyPoint>& features) {
    if (matcherParameters.OpenCVParams.descriptor != "ORB") return MatcherOpenCV::describeFeatures(rgbImage, features);
    return core.describeFeatures(rgbImage, features);
}

std::vector<cv::DMatch> MatcherB200::performMatching(cv::Mat prevDescriptors, cv::Mat descriptors) {
    const std::string& d = matcherParameters.OpenCVParams.descriptor;
    if (d == "SURF" || d == "SIFT") return MatcherOpenCV::performMatching(prevDescriptors, descriptors);   // float L2 descriptors
    return core.performMatching(prevDescriptors, descriptors);
}

std::vector<cv::DMatch> MatcherB200::performTracking(cv::Mat prevImg, cv::Mat img, std::vector<cv::Point2f>& prevFeatures,
                                                     std::vector<cv::Point2f>& features, std::vector<cv::KeyPoint>& prevKeyPoints,
                                                     std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& prevDetDists,
                                                     std::vector<double>& detDists) {
    const Matcher::parameters& o = matcherParameters.OpenCVParams;
    putslam_b200::MatcherB200::TrackingParams tp;
    tp.winSize = o.winSize; tp.maxLevels = o.maxLevels; tp.maxIter = o.maxIter; tp.eps = o.eps;
    tp.useInitialFlow = o.useInitialFlow; tp.trackingErrorType = o.trackingErrorType;
    tp.trackingErrorThreshold = o.trackingErrorThreshold; tp.trackingMinEigThreshold = o.trackingMinEigThreshold;
    tp.minimalReprojDistanceNewTrackingFeatures = o.minimalReprojDistanceNewTrackingFeatures;
    core.setTrackingParams(tp);
    return core.performTracking(prevImg, img, prevFeatures, features, prevKeyPoints, keyPoints, prevDetDists, detDists);
}

putslam::Matcher* putslam::createMatcherB200(const std::string _parametersFile, const std::string _grabberParametersFile) {
    matcherClassB200.reset(new MatcherB200(_parametersFile, _grabberParametersFile));
    return matcherClassB200.get();
}

putslam::Matcher* putslam::createloopClosingMatcherB200(const std::string _parametersFile, const std::string _grabberParametersFile) {
    loopClosingMatcherClassB200.reset(new MatcherB200(_parametersFile, _grabberParametersFile));
    return loopClosingMatcherClassB200.get();
}
