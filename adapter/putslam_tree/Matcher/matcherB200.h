/** @file matcherB200.h -- the file a PUTSLAM maintainer adds as include/putslam/Matcher/matcherB200.h
 *
 * MatcherB200 derives from the reference's own MatcherOpenCV (include/putslam/Matcher/matcherOpenCV.h:31-87) and overrides
 * the four virtuals of the hot path (include/putslam/Matcher/matcher.h:405-422) with the B200 implementations behind
 * adapter/pslam_adapter.h.  Everything else of Matcher -- runVO, match, matchXYZ, trackKLT, the parameter loading -- is
 * inherited unchanged, so Tracker / Map / PoseGraph call the new path through the reference's own dispatch.
 *
 * `override` makes the compiler check every signature against the reference's declarations: this file is compiled
 * against the reference's real matcher.h / matcherOpenCV.h by `make -C adapter tree` (tests/test_boundary_conformance_cpu.py).
 */
#ifndef MATCHERB200_H_INCLUDED
#define MATCHERB200_H_INCLUDED

#include "Matcher/matcherOpenCV.h"
#include "pslam_adapter.h"

namespace putslam {
/// factories with the reference's signatures (matcherOpenCV.h:17-22): tracking matcher and loop-closure matcher
Matcher* createMatcherB200(const std::string _parametersFile, const std::string _grabberParametersFile);
Matcher* createloopClosingMatcherB200(const std::string _parametersFile, const std::string _grabberParametersFile);
}

class MatcherB200 : public MatcherOpenCV {
public:
    typedef std::unique_ptr<MatcherB200> Ptr;

    MatcherB200(const std::string _parametersFile, const std::string _grabberParametersFile, int device = 0)
        : MatcherOpenCV(_parametersFile, _grabberParametersFile), core(device) {}

    /// MatcherOpenCV::detectFeatures (matcherOpenCV.cpp:118-176) for detector == "ORB" / "FAST"
    std::vector<cv::KeyPoint> detectFeatures(cv::Mat rgbImage) override;

    /// MatcherOpenCV::describeFeatures (matcherOpenCV.cpp:181-195) for descriptor == "ORB"
    cv::Mat describeFeatures(cv::Mat rgbImage, std::vector<cv::KeyPoint>& features) override;

    /// MatcherOpenCV::performMatching (matcherOpenCV.cpp:198-206): cv::BFMatcher(NORM_HAMMING, crossCheck = true)
    std::vector<cv::DMatch> performMatching(cv::Mat prevDescriptors, cv::Mat descriptors) override;

    /// MatcherOpenCV::performTracking (matcherOpenCV.cpp:209-300): calcOpticalFlowPyrLK + thresholds + too-close rule
    std::vector<cv::DMatch> performTracking(cv::Mat prevImg, cv::Mat img, std::vector<cv::Point2f>& prevFeatures,
                                            std::vector<cv::Point2f>& features, std::vector<cv::KeyPoint>& prevKeyPoints,
                                            std::vector<cv::KeyPoint>& keyPoints, std::vector<double>& prevDetDists,
                                            std::vector<double>& detDists) override;

    putslam_b200::MatcherB200 core;   ///< owns one pslam_ctx: one per Matcher instance, as PUTSLAM runs one per thread
};

#endif  // MATCHERB200_H_INCLUDED
