// selftest.cpp -- drives the C++ adapter exactly the way PUTSLAM's call sites do (Matcher::match,
// Matcher::matchXYZ, demoKabsch) on inputs written by tests/test_gpu_adapter.py, and writes the results
// back as raw arrays for comparison with the oracle.  usage: adapter_selftest <dir>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
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
static cv::Mat descMat(std::vector<uint8_t>& raw) { return cv::Mat((int)(raw.size() / 32), 32, CV_8U, raw.data()); }
static void wrMatches(const std::string& name, const std::vector<cv::DMatch>& m) {
    std::vector<int> q, t, img; std::vector<float> d;
    for (auto& x : m) { q.push_back(x.queryIdx); t.push_back(x.trainIdx); img.push_back(x.imgIdx); d.push_back(x.distance); }
    wr(name + "_q.bin", q); wr(name + "_t.bin", t); wr(name + "_img.bin", img); wr(name + "_d.bin", d);
}

// ---- Matcher::trackKLT's tracking step (matcher.cpp:151-158): performTracking over a three-frame sequence ----
// frame 0 -> 1 uploads both frames; 1 -> 2 passes the Mat it tracked into as prevImg (prevRgbImage), whose pyramid is
// still resident; the third call (useInitialFlow, min-eigenvalue error) goes 0 -> 2 with fresh uploads.
static int kltSection() {
    auto dims = rd<int>("klt_dims.bin");                            // H, W, channels
    auto f0 = rd<uint8_t>("klt_f0.bin"), f1 = rd<uint8_t>("klt_f1.bin"), f2 = rd<uint8_t>("klt_f2.bin");
    auto xy = rd<float>("klt_xy.bin");
    const int type = dims[2] == 3 ? CV_8UC3 : CV_8U;
    cv::Mat m0(dims[0], dims[1], type, f0.data()), m1(dims[0], dims[1], type, f1.data()), m2(dims[0], dims[1], type, f2.data());
    MatcherB200 matcher(0);
    matcher.setReuseTrackedFrame(true);
    std::vector<cv::Point2f> prev(xy.size() / 2), cur;
    std::vector<cv::KeyPoint> prevKp(prev.size()), curKp;
    std::vector<double> prevDet(prev.size()), curDet;
    for (size_t i = 0; i < prev.size(); ++i) {
        prev[i] = cv::Point2f(xy[2 * i], xy[2 * i + 1]);
        prevKp[i].pt = prev[i]; prevKp[i].class_id = (int)i; prevKp[i].octave = (int)(i % 5); prevDet[i] = 0.25 * (double)i;
    }
    auto dump = [&](const std::string& tag, const std::vector<cv::DMatch>& m) {
        wrMatches(tag, m);
        std::vector<float> p; std::vector<int> ids, octs;
        for (size_t j = 0; j < cur.size(); ++j) {
            p.push_back(cur[j].x); p.push_back(cur[j].y); p.push_back(curKp[j].pt.x); p.push_back(curKp[j].pt.y);
            ids.push_back(curKp[j].class_id); octs.push_back(curKp[j].octave);
        }
        wr(tag + "_xy.bin", p); wr(tag + "_id.bin", ids); wr(tag + "_oct.bin", octs); wr(tag + "_det.bin", curDet);
    };
    std::vector<cv::DMatch> m01 = matcher.performTracking(m0, m1, prev, cur, prevKp, curKp, prevDet, curDet);
    dump("klt01", m01);
    std::vector<cv::Point2f> prev1 = cur;
    std::vector<cv::KeyPoint> prevKp1 = curKp;
    std::vector<double> prevDet1 = curDet;
    std::vector<cv::DMatch> m12 = matcher.performTracking(m1, m2, prev1, cur, prevKp1, curKp, prevDet1, curDet);
    dump("klt12", m12);
    MatcherB200::TrackingParams tp;
    tp.useInitialFlow = 1; tp.trackingErrorType = 1; tp.trackingErrorThreshold = 1e9; tp.trackingMinEigThreshold = 1e-3;
    tp.minimalReprojDistanceNewTrackingFeatures = 1.5; tp.winSize = 9; tp.maxLevels = 2;
    matcher.setTrackingParams(tp);
    cur = prev;                                                     // the guess: no motion
    std::vector<cv::DMatch> m02 = matcher.performTracking(m0, m2, prev, cur, prevKp, curKp, prevDet, curDet);
    dump("klt02", m02);
    // ---- the fused frame (MatcherB200::trackKLTCore == Matcher::trackKLT lines 151-207), shipped parameters ----
    std::ifstream probe(g_dir + "/klt_depth1.bin", std::ios::binary);
    if (probe) {
        auto z1 = rd<uint16_t>("klt_depth1.bin");
        auto pxyz = rd<float>("klt_prev_xyz.bin");
        auto camv = rd<float>("klt_cam.bin");                       // fx fy cx cy
        float Kf[9] = {camv[0], 0, camv[2], 0, camv[1], camv[3], 0, 0, 1};
        float Df[5] = {-0.0410f, 0.3286f, 0.0087f, 0.0051f, -0.5643f};
        cv::Mat K(3, 3, CV_32FC1, Kf), D(1, 5, CV_32FC1, Df);
        cv::Mat depth1(dims[0], dims[1], CV_16U, z1.data());
        std::vector<Eigen::Vector3f> prev3D(prev.size());
        for (size_t i = 0; i < prev3D.size(); ++i) prev3D[i] = Eigen::Vector3f(pxyz[3 * i], pxyz[3 * i + 1], pxyz[3 * i + 2]);
        RANSAC::parameters rp;
        rp.verbose = 0; rp.errorVersion = 2; rp.errorVersionVO = 0; rp.errorVersionMap = 0;
        rp.inlierThresholdEuclidean = 0.04; rp.inlierThresholdReprojection = 2.0; rp.inlierThresholdMahalanobis = 9.0;
        rp.minimalInlierRatioThreshold = 0.2; rp.minimalNumberOfMatches = 15; rp.usedPairs = 3; rp.iterationCount = 0;
        MatcherB200 fused(0);
        fused.setSeed(99);
        std::vector<cv::Point2f> dist2D, und2D; std::vector<Eigen::Vector3f> f3D; std::vector<cv::DMatch> mF, inF;
        Eigen::Matrix4f TF;
        const double ratio = fused.trackKLTCore(m0, m1, prev, prev3D, prevKp, prevDet, depth1, 5000.0, K, D, rp, dist2D, und2D, f3D,
                                                curKp, curDet, mF, inF, TF);
        cur = dist2D;
        dump("kltf", mF);
        wrMatches("kltf_inliers", inF);
        std::vector<float> u, x;
        for (size_t j = 0; j < und2D.size(); ++j) { u.push_back(und2D[j].x); u.push_back(und2D[j].y); x.push_back(f3D[j][0]); x.push_back(f3D[j][1]); x.push_back(f3D[j][2]); }
        wr("kltf_und.bin", u); wr("kltf_xyz.bin", x);
        wr("kltf_T.bin", std::vector<float>(TF.data(), TF.data() + 16));
        wr("kltf_ratio.bin", std::vector<double>{ratio});
    }
    std::cout << "adapter_selftest klt ok: " << m01.size() << " / " << m12.size() << " / " << m02.size() << " tracked" << std::endl;
    return 0;
}

// ---- demoKabsch.cpp:1020-1028: trans = est->computeTransformation(setA, setB); est->computeUncertainty(...) ----
static int uncSection() {
    auto A = rd<double>("kabsch_A.bin"), B = rd<double>("kabsch_B.bin");   // row-major n x 3, B ~ R A + t
    auto ca = rd<double>("unc_CA.bin"), cb = rd<double>("unc_CB.bin");     // n x 3 x 3 row-major
    const long n = (long)(A.size() / 3);
    Eigen::MatrixXd setA(n, 3), setB(n, 3);
    for (long r = 0; r < n; ++r) for (int c = 0; c < 3; ++c) { setA(r, c) = A[3 * r + c]; setB(r, c) = B[3 * r + c]; }
    std::vector<Mat33> ua((size_t)n), ub((size_t)n);
    for (long r = 0; r < n; ++r)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) { ua[(size_t)r](i, j) = ca[9 * r + 3 * i + j]; ub[(size_t)r](i, j) = cb[9 * r + 3 * i + j]; }
    TransformEst* est = createKabschEstimator();
    Mat34 trans = est->computeTransformation(setA, setB);
    wr("unc_T.bin", std::vector<double>(trans.m, trans.m + 16));
    // the estimator maps A onto B (B ~ R A + t), so B plays the role of computeUncertainty's setA
    const Mat66& u1 = est->computeUncertainty(setB, ub, setA, ua, trans);
    wr("unc_euler.bin", std::vector<double>(u1.m, u1.m + 36));
    const Mat66& u2 = est->computeUncertaintyG2O(setB, ub, setA, ua, trans);
    wr("unc_quat.bin", std::vector<double>(u2.m, u2.m + 36));
    std::cout << "adapter_selftest unc ok" << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: adapter_selftest <dir> [klt|unc]" << std::endl; return 2; }
    g_dir = argv[1];
    if (argc > 2 && std::string(argv[2]) == "klt") return kltSection();
    if (argc > 2 && std::string(argv[2]) == "unc") return uncSection();
    float Kf[9] = {517.3f, 0, 318.6f, 0, 516.5f, 255.3f, 0, 0, 1};
    float Df[5] = {-0.0410f, 0.3286f, 0.0087f, 0.0051f, -0.5643f};
    cv::Mat K(3, 3, CV_32FC1, Kf), D(1, 5, CV_32FC1, Df);

    // ---- Matcher::match path (matcher.cpp:470-496): performMatching -> undistort -> back-project -> RANSAC ----
    auto d1 = rd<uint8_t>("desc1.bin"), d2 = rd<uint8_t>("desc2.bin");
    auto uv1 = rd<float>("uv1.bin"), uv2 = rd<float>("uv2.bin");
    auto z1 = rd<uint16_t>("depth1.bin"), z2 = rd<uint16_t>("depth2.bin");
    MatcherB200 matcher(0);
    std::vector<cv::DMatch> matches = matcher.performMatching(descMat(d1), descMat(d2));
    wrMatches("vo_matches", matches);
    std::vector<cv::KeyPoint> kp1(uv1.size() / 2), kp2(uv2.size() / 2);
    for (size_t i = 0; i < kp1.size(); ++i) kp1[i].pt = cv::Point2f(uv1[2 * i], uv1[2 * i + 1]);
    for (size_t i = 0; i < kp2.size(); ++i) kp2[i].pt = cv::Point2f(uv2[2 * i], uv2[2 * i + 1]);
    std::vector<cv::Point2f> und1 = RGBD::removeImageDistortion(kp1, K, D), und2 = RGBD::removeImageDistortion(kp2, K, D);
    cv::Mat depth1(480, 640, CV_16U, z1.data()), depth2(480, 640, CV_16U, z2.data());
    std::vector<Eigen::Vector3f> p1 = RGBD::keypoints2Dto3D(und1, depth1, K, 5000.0), p2 = RGBD::keypoints2Dto3D(und2, depth2, K, 5000.0);
    std::vector<float> und2f, p2f;
    for (auto& p : und2) { und2f.push_back(p.x); und2f.push_back(p.y); }
    for (auto& p : p2) { p2f.push_back(p[0]); p2f.push_back(p[1]); p2f.push_back(p[2]); }
    wr("vo_und2.bin", und2f); wr("vo_xyz2.bin", p2f);
    RANSAC::parameters rp;
    rp.verbose = 0; rp.errorVersion = rp.errorVersionVO = rp.errorVersionMap = 0;
    rp.inlierThresholdEuclidean = 0.04; rp.inlierThresholdReprojection = 2.0; rp.inlierThresholdMahalanobis = 9.0;
    rp.minimalInlierRatioThreshold = 0.2; rp.minimalNumberOfMatches = 15; rp.usedPairs = 3; rp.iterationCount = 0;
    rp.errorVersion = rp.errorVersionVO;   // matcher.cpp:491-492
    RANSAC ransac(rp, K);
    ransac.setSeed(4242);
    std::vector<cv::DMatch> inliers;
    Eigen::Matrix4f T = ransac.estimateTransformation(p1, p2, matches, inliers);
    wrMatches("vo_inliers", inliers);
    wr("vo_T.bin", std::vector<float>(T.data(), T.data() + 16));
    std::vector<double> ratio{RANSAC::pointInlierRatio(inliers, matches), (double)ransac.hypothesesUsed(), ransac.bestInlierRatio()};
    wr("vo_ratio.bin", ratio);

    // ---- the same VO step fused into one submission (MatcherB200::matchCore) ----
    {
        std::vector<cv::Point2f> undF; std::vector<Eigen::Vector3f> x2F; std::vector<cv::DMatch> mF, inF; Eigen::Matrix4f TF;
        matcher.setSeed(4242);
        const double rF = matcher.matchCore(descMat(d1), p1, descMat(d2), kp2, depth2, 5000.0, K, D, rp, undF, x2F, mF, inF, TF);
        wrMatches("vof_matches", mF); wrMatches("vof_inliers", inF);
        wr("vof_T.bin", std::vector<float>(TF.data(), TF.data() + 16));
        std::vector<float> xf;
        for (auto& p : x2F) { xf.push_back(p[0]); xf.push_back(p[1]); xf.push_back(p[2]); }
        wr("vof_xyz2.bin", xf);
        wr("vof_ratio.bin", std::vector<double>{rF});
    }

    // ---- Matcher::matchXYZ path (matcher.cpp:606-797) ----
    MatcherB200::MapSide map;
    map.xyz = rd<double>("map_xyz.bin");
    auto mdesc = rd<uint8_t>("map_desc.bin");
    map.descriptors = descMat(mdesc);
    map.octave = rd<int>("map_octave.bin");
    map.detDist = rd<double>("map_detdist.bin");
    auto cxyz = rd<float>("cur_xyz.bin");
    auto cdesc = rd<uint8_t>("cur_desc.bin");
    auto coct = rd<int>("cur_octave.bin");
    std::vector<double> cdet = rd<double>("cur_detdist.bin");
    std::vector<Eigen::Vector3f> cur3D(cxyz.size() / 3);
    std::vector<cv::KeyPoint> curKp(cur3D.size());
    for (size_t i = 0; i < cur3D.size(); ++i) { cur3D[i] = Eigen::Vector3f(cxyz[3 * i], cxyz[3 * i + 1], cxyz[3 * i + 2]); curKp[i].octave = coct[i]; }
    for (int cn = 1; cn <= 2; ++cn) {
        matcher.setHostLevels(cn == 2);   // first call: levels on the device, retry: host libm levels (both must agree with the oracle)
        Eigen::Matrix4f Tm;
        std::vector<cv::DMatch> mm, mi;
        matcher.setSeed(77);
        const double r = matcher.matchXYZCore(map, descMat(cdesc), cur3D, curKp, cdet, 0.12, 0.55, cn, rp, K, Tm, mm, mi);
        const std::string tag = "map" + std::to_string(cn);
        wrMatches(tag + "_matches", mm); wrMatches(tag + "_inliers", mi);
        wr(tag + "_T.bin", std::vector<float>(Tm.data(), Tm.data() + 16));
        wr(tag + "_ratio.bin", std::vector<double>{r});
    }

    // ---- the same frame against the map resident in HBM (identity pose, every view axis = optical axis) ----
    {
        std::vector<float> axes(3 * map.octave.size(), 0.f);
        for (size_t j = 0; j < map.octave.size(); ++j) axes[3 * j + 2] = 1.f;
        const size_t half = map.octave.size() / 2;          // two uploads: the second appends
        MatcherB200::MapSide a, b;
        a.xyz.assign(map.xyz.begin(), map.xyz.begin() + 3 * half); b.xyz.assign(map.xyz.begin() + 3 * half, map.xyz.end());
        a.octave.assign(map.octave.begin(), map.octave.begin() + half); b.octave.assign(map.octave.begin() + half, map.octave.end());
        a.detDist.assign(map.detDist.begin(), map.detDist.begin() + half); b.detDist.assign(map.detDist.begin() + half, map.detDist.end());
        a.descriptors = cv::Mat((int)half, 32, CV_8U, map.descriptors.data);
        b.descriptors = cv::Mat((int)(map.octave.size() - half), 32, CV_8U, map.descriptors.data + 32 * half);
        bool ok = matcher.uploadMapFeatures(0, a, std::vector<float>(axes.begin(), axes.begin() + 3 * half));
        ok = ok && matcher.uploadMapFeatures((int)half, b, std::vector<float>(axes.begin() + 3 * half, axes.end()));
        if (!ok || matcher.mapSize() != (int)map.octave.size()) { std::cerr << "resident map upload failed" << std::endl; return 3; }
        const double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        MatcherB200::MapFilter filt;
        filt.maxZ = 1e9;
        matcher.setHostLevels(false);
        matcher.setSeed(77);
        Eigen::Matrix4f Tm;
        std::vector<cv::DMatch> mm, mi;
        std::vector<int> kept;
        const double r = matcher.matchXYZResident(eye, filt, descMat(cdesc), cur3D, curKp, cdet, 0.12, 0.55, 1, rp, K, Tm, kept, mm, mi);
        wrMatches("mapr_matches", mm); wrMatches("mapr_inliers", mi);
        wr("mapr_T.bin", std::vector<float>(Tm.data(), Tm.data() + 16));
        wr("mapr_ratio.bin", std::vector<double>{r});
        wr("mapr_kept.bin", kept);
    }

    // ---- describeFeatures (matcherOpenCV.cpp:181-195): ORB descriptors for provided keypoints, colour image ----
    {
        std::ifstream probe(g_dir + "/orb_bgr.bin", std::ios::binary);
        if (probe) {
            auto bgr = rd<uint8_t>("orb_bgr.bin");
            auto dims = rd<int>("orb_dims.bin");                  // H, W
            auto kxy = rd<float>("orb_xy.bin"); auto koct = rd<int>("orb_octave.bin"); auto kang = rd<float>("orb_angle.bin");
            cv::Mat img(dims[0], dims[1], CV_8UC3, bgr.data());
            std::vector<cv::KeyPoint> feats(koct.size());
            for (size_t i = 0; i < feats.size(); ++i) {
                feats[i].pt = cv::Point2f(kxy[2 * i], kxy[2 * i + 1]); feats[i].octave = koct[i]; feats[i].angle = kang[i];
                feats[i].class_id = (int)i;                        // to recover the permutation
            }
            cv::Mat d = matcher.describeFeatures(img, feats);
            std::vector<int> order(feats.size());
            for (size_t i = 0; i < feats.size(); ++i) order[i] = feats[i].class_id;
            wr("orb_order.bin", order);
            wr("orb_desc.bin", std::vector<uint8_t>(d.data, d.data + 32 * (size_t)d.rows));
        }
    }

    // ---- detectFeatures (matcherOpenCV.cpp:118-176) on a colour frame, 1 x 1 (shipped config) and 2 x 2 grids ----
    {
        std::ifstream probe(g_dir + "/det_rgb.bin", std::ios::binary);
        if (probe) {
            auto rgb = rd<uint8_t>("det_rgb.bin");
            auto dims = rd<int>("det_dims.bin");
            cv::Mat img(dims[0], dims[1], CV_8UC3, rgb.data());
            for (int gridn = 1; gridn <= 2; ++gridn) {
                std::vector<cv::KeyPoint> kps = matcher.detectFeatures(img, gridn, gridn, 500);
                std::vector<float> rows;
                std::vector<int> octs;
                for (const cv::KeyPoint& k : kps) {
                    rows.push_back(k.pt.x); rows.push_back(k.pt.y); rows.push_back(k.size); rows.push_back(k.angle);
                    rows.push_back(k.response); octs.push_back(k.octave);
                }
                wr("det" + std::to_string(gridn) + "_kp.bin", rows);
                wr("det" + std::to_string(gridn) + "_octave.bin", octs);
            }
            {   // detector == "FAST"
                std::vector<cv::KeyPoint> kps = matcher.detectFeaturesFAST(img, 1, 1, 500);
                std::vector<float> rows;
                for (const cv::KeyPoint& k : kps) {
                    rows.push_back(k.pt.x); rows.push_back(k.pt.y); rows.push_back(k.size); rows.push_back(k.angle);
                    rows.push_back(k.response); rows.push_back((float)k.octave);
                }
                wr("detfast_kp.bin", rows);
            }
        }
    }

    // ---- demoKabsch path (demoKabsch.cpp:1020): createKabschEstimator()->computeTransformation(A, B) ----
    auto A = rd<double>("kabsch_A.bin"), B = rd<double>("kabsch_B.bin");   // row-major n x 3
    const long n = (long)(A.size() / 3);
    Eigen::MatrixXd setA(n, 3), setB(n, 3);
    for (long r = 0; r < n; ++r) for (int c = 0; c < 3; ++c) { setA(r, c) = A[3 * r + c]; setB(r, c) = B[3 * r + c]; }
    TransformEst* est = createKabschEstimator();
    Mat34& Tk = est->computeTransformation(setA, setB);
    wr("kabsch_T.bin", std::vector<double>(Tk.m, Tk.m + 16));
    Mat34& Te = est->computeTransformation(Eigen::MatrixXd(0, 3), Eigen::MatrixXd(0, 3));
    wr("kabsch_T_empty.bin", std::vector<double>(Te.m, Te.m + 16));
    std::cout << "adapter_selftest ok: " << matches.size() << " VO matches, " << inliers.size() << " inliers, est " << est->getName() << std::endl;
    return 0;
}
