// dbscan_cli.cpp -- runs putslam_b200::DBScan on a keypoint list from a file (host only, no GPU): the CPU test-suite
// compares it with the reference's own DBScan compiled from the reference tree.
// usage: dbscan_cli <xy.bin (n x 2 float32)> <eps> <minPts> <featuresFromCluster> <kept.bin (int32 indices)> [labels.bin]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "pslam_adapter.h"

int main(int argc, char** argv) {
    if (argc < 6) { std::fprintf(stderr, "usage: dbscan_cli xy.bin eps minPts featuresFromCluster kept.bin [labels.bin]\n"); return 2; }
    std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
    if (!f) return 2;
    const size_t bytes = (size_t)f.tellg();
    std::vector<float> xy(bytes / 4);
    f.seekg(0);
    f.read((char*)xy.data(), (std::streamsize)bytes);
    std::vector<cv::KeyPoint> kps(xy.size() / 2);
    for (size_t i = 0; i < kps.size(); ++i) { kps[i].pt = cv::Point2f(xy[2 * i], xy[2 * i + 1]); kps[i].class_id = (int)i; }
    putslam_b200::DBScan d(atof(argv[2]), atoi(argv[3]), atoi(argv[4]));
    d.run(kps);
    std::vector<int> kept;
    for (const cv::KeyPoint& k : kps) kept.push_back(k.class_id);
    std::ofstream o(argv[5], std::ios::binary);
    o.write((const char*)kept.data(), (std::streamsize)(kept.size() * sizeof(int)));
    if (argc > 6) {
        std::ofstream l(argv[6], std::ios::binary);
        l.write((const char*)d.labels().data(), (std::streamsize)(d.labels().size() * sizeof(int)));
    }
    return 0;
}
