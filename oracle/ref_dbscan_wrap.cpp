// C entry point around the REFERENCE's DBScan class (src/Matcher/dbscan.cpp, compiled from /root/reference by
// `make -C oracle ref` into oracle/_ref/libref_dbscan.so).  TEST INFRASTRUCTURE ONLY.
#include <vector>

#include "Matcher/dbscan.h"

extern "C" __attribute__((visibility("default")))
int orc_ref_dbscan(const float* xy, int n, double eps, int minPts, int featuresFromCluster, int* kept /* cap n */) {
    std::vector<cv::KeyPoint> kps((size_t)n);
    for (int i = 0; i < n; ++i) { kps[(size_t)i].pt = cv::Point2f(xy[2 * i], xy[2 * i + 1]); kps[(size_t)i].class_id = i; }
    DBScan d(eps, minPts, featuresFromCluster);
    d.run(kps);
    for (size_t k = 0; k < kps.size(); ++k) kept[k] = kps[k].class_id;
    return (int)kps.size();
}
