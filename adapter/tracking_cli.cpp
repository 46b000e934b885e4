// tracking_cli.cpp -- runs the host-side steps of Matcher::trackKLT (putslam_b200::tracking, pslam_adapter.h) on arrays from
// a directory (host only, no GPU): the CPU test-suite compares them with numpy restatements of the reference functions
// and the grid-accelerated form with the brute-force one.
// usage: tracking_cli <dir> remove <minEuclid> <minReproj> <brute 0|1>
//        tracking_cli <dir> merge  <minReproj> <brute 0|1>
//        tracking_cli <dir> levels
//        tracking_cli <dir> undescribed   (the features whose key point survived describeFeatures)
//        tracking_cli <dir> strasdat      (TransformEst::computeUncertaintyStrasdat, host arithmetic)
// inputs (float32 unless noted): und.bin n x 2, dist.bin n x 2, xyz.bin n x 3, oct.bin int32 n, det.bin float64 n,
// matches.bin int32 m x 2 (remove), s_und.bin / s_dist.bin / s_xyz.bin / s_oct.bin / s_det.bin (merge);
// outputs: id.bin (int32: original index of every surviving / appended feature, sandbox features numbered from n),
// removed.bin, matches_out.bin, desc_oct.bin (levels: predicted octave per ORIGINAL index)
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "pslam_adapter.h"

using namespace putslam_b200;

static std::string g_dir;
template <typename T>
static std::vector<T> rd(const std::string& name) {
    std::ifstream f(g_dir + "/" + name, std::ios::binary | std::ios::ate);
    if (!f) { std::cerr << "missing " << name << std::endl; exit(2); }
    const size_t n = (size_t)f.tellg();
    std::vector<T> v(n / sizeof(T));
    f.seekg(0);
    f.read((char*)v.data(), (std::streamsize)n);
    return v;
}
template <typename T>
static void wr(const std::string& name, const std::vector<T>& v) {
    std::ofstream f(g_dir + "/" + name, std::ios::binary);
    f.write((const char*)v.data(), (std::streamsize)(v.size() * sizeof(T)));
}
struct Lists {
    std::vector<cv::Point2f> und, dist;
    std::vector<Eigen::Vector3f> xyz;
    std::vector<cv::KeyPoint> kp;
    std::vector<double> det;
};
static Lists load(const std::string& prefix, int idBase) {
    Lists L;
    auto u = rd<float>(prefix + "und.bin"), d = rd<float>(prefix + "dist.bin"), x = rd<float>(prefix + "xyz.bin");
    auto o = rd<int>(prefix + "oct.bin");
    L.det = rd<double>(prefix + "det.bin");
    const size_t n = o.size();
    L.und.resize(n); L.dist.resize(n); L.xyz.resize(n); L.kp.resize(n);
    for (size_t i = 0; i < n; ++i) {
        L.und[i] = cv::Point2f(u[2 * i], u[2 * i + 1]); L.dist[i] = cv::Point2f(d[2 * i], d[2 * i + 1]);
        L.xyz[i] = Eigen::Vector3f(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
        L.kp[i].pt = L.dist[i]; L.kp[i].octave = o[i]; L.kp[i].class_id = idBase + (int)i;
    }
    return L;
}
static int dump(const Lists& L) {
    const size_t n = L.kp.size();
    if (L.und.size() != n || L.dist.size() != n || L.xyz.size() != n || L.det.size() != n) { std::cerr << "sizes differ" << std::endl; return 3; }
    std::vector<int> id;
    std::vector<float> row;
    for (size_t i = 0; i < n; ++i) {
        id.push_back(L.kp[i].class_id);
        row.push_back(L.und[i].x); row.push_back(L.und[i].y); row.push_back(L.dist[i].x); row.push_back(L.dist[i].y);
        row.push_back(L.xyz[i][0]); row.push_back(L.xyz[i][1]); row.push_back(L.xyz[i][2]); row.push_back((float)L.det[i]);
    }
    wr("id.bin", id); wr("rows.bin", row);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::cerr << "usage: tracking_cli <dir> remove|merge|levels ..." << std::endl; return 2; }
    g_dir = argv[1];
    const std::string op = argv[2];
    if (op == "strasdat") {   // A.bin, B.bin: n x 3 float64 row-major; T.bin: 4 x 4 float64 row-major -> U.bin 6 x 6 row-major
        auto A = rd<double>("A.bin"), B = rd<double>("B.bin"), T = rd<double>("T.bin");
        const long n = (long)(A.size() / 3);
        Eigen::MatrixXd setA(n, 3), setB(n, 3);
        for (long r = 0; r < n; ++r) for (int c = 0; c < 3; ++c) { setA(r, c) = A[3 * r + c]; setB(r, c) = B[3 * r + c]; }
        Mat34 trans;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) trans.m[4 * j + i] = T[4 * i + j];
        KabschEst est;
        const Mat66& u = est.computeUncertaintyStrasdat(setA, setB, trans);
        std::vector<double> out(36);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[6 * i + j] = u(i, j);
        wr("U.bin", out);
        return 0;
    }
    Lists L = load("", 0);
    if (op == "remove" && argc >= 6) {
        auto mm = rd<int>("matches.bin");
        std::vector<cv::DMatch> matches;
        for (size_t k = 0; k + 1 < mm.size(); k += 2) matches.push_back(cv::DMatch(mm[k], mm[k + 1], 0));
        std::set<int> gone = tracking::removeTooCloseFeatures(L.dist, L.und, L.xyz, L.kp, L.det, matches, atof(argv[3]), atof(argv[4]),
                                                              atoi(argv[5]) != 0);
        wr("removed.bin", std::vector<int>(gone.begin(), gone.end()));
        std::vector<int> mo;
        for (const cv::DMatch& m : matches) { mo.push_back(m.queryIdx); mo.push_back(m.trainIdx); }
        wr("matches_out.bin", mo);
        return dump(L);
    }
    if (op == "merge" && argc >= 5) {
        Lists S = load("s_", (int)L.kp.size());
        tracking::mergeTrackedFeatures(L.und, S.und, L.dist, S.dist, L.xyz, S.xyz, L.kp, S.kp, L.det, S.det, atof(argv[3]), atoi(argv[4]) != 0);
        return dump(L);
    }
    if (op == "levels") {
        std::vector<cv::KeyPoint> desc = tracking::predictDescriptionLevels(L.dist, L.und, L.xyz, L.kp, L.det);
        std::vector<int> oc;
        for (const cv::KeyPoint& k : desc) oc.push_back(k.octave);
        wr("desc_oct.bin", oc);
        return dump(L);
    }
    if (op == "undescribed") {   // desc_xy.bin: positions of the key points describeFeatures returned (m x 2 float32)
        auto dxy = rd<float>("desc_xy.bin");
        std::vector<cv::KeyPoint> desc(dxy.size() / 2);
        for (size_t j = 0; j < desc.size(); ++j) desc[j].pt = cv::Point2f(dxy[2 * j], dxy[2 * j + 1]);
        tracking::dropUndescribed(desc, L.dist, L.und, L.xyz, L.kp, L.det);
        return dump(L);
    }
    std::cerr << "bad arguments" << std::endl;
    return 2;
}
