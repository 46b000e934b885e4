// Stand-in for the OpenCV C++ headers -- TEST INFRASTRUCTURE ONLY (nothing in the product includes this file).
//
// OpenCV's C++ headers/libraries do not exist in this image (only the Python cv2 module does).  To compile the
// REFERENCE's own translation units where they lie (oracle/Makefile, target ref -> oracle/_ref/), this header declares
// the small part of the cv:: API those files touch, with OpenCV's documented semantics:
//   Point_/Point3_/Vec/Rect/Size/Scalar/KeyPoint/DMatch   core/types.hpp (Point_<float> -> Point_<int> rounds with
//                                                         saturate_cast<int> = cvRound, half to even)
//   Mat                                                   a reference-counted 2-D array: at<T>, row, push_back, ROI,
//                                                         operator- (saturating for CV_8U), clone
//   norm(Mat, NORM_HAMMING | NORM_L2), norm(Point2f)      core: popcount of the bytes / sqrt of the double sum
//   undistortPoints                                       imgproc: 5 fixed-point iterations of the inverse Brown model
//                                                         in double (pinned bit for bit against cv2 4.13 through
//                                                         tests/golden/undistort_cv2.npz -- tests/test_ref_build_cpu.py
//                                                         checks this restatement against the same golden file)
//   imshow / waitKey / drawKeypoints / drawMatches / circle / cvtColor   no-ops (debug display, never on the path)
#ifndef PSLAM_REF_SHIM_CV_H
#define PSLAM_REF_SHIM_CV_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_RGB2GRAY 7
#define CV_BGR2GRAY 6

namespace cv {

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7 };

inline int cvRound(double v) { return (int)std::lrint(v); }       // round half to even (default rounding mode)

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> operator Point_<U>() const;
};
template <> template <> inline Point_<float>::operator Point_<int>() const { return Point_<int>(cvRound(x), cvRound(y)); }
template <> template <> inline Point_<int>::operator Point_<float>() const { return Point_<float>((float)x, (float)y); }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
// core/types.hpp: norm(Point_<T>) = std::sqrt((double)pt.x*pt.x + (double)pt.y*pt.y)
template <typename T> double norm(const Point_<T>& pt) { return std::sqrt((double)pt.x * pt.x + (double)pt.y * pt.y); }

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(); }
    template <typename U> Vec(const Vec<U, N>& o) { for (int i = 0; i < N; ++i) val[i] = (T)o.val[i]; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<int, 3> Vec3i;

struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
struct TermCriteria { enum { COUNT = 1, EPS = 2 }; int type, maxCount; double epsilon; TermCriteria(int t = 0, int m = 0, double e = 0) : type(t), maxCount(m), epsilon(e) {} };

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f p, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(p), size(s), angle(a), response(r), octave(o), class_id(c) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
    // features2d: KeyPoint::convert(keypoints, points2f): the .pt of every keypoint, in order
    static void convert(const std::vector<KeyPoint>& kps, std::vector<Point2f>& pts, const std::vector<int>& = std::vector<int>()) {
        pts.resize(kps.size());
        for (std::size_t i = 0; i < kps.size(); ++i) pts[i] = kps[i].pt;
    }
    // convert(points2f, keypoints, size = 1, response = 1, octave = 0, class_id = -1)
    static void convert(const std::vector<Point2f>& pts, std::vector<KeyPoint>& kps, float size = 1, float response = 1, int octave = 0, int class_id = -1) {
        kps.resize(pts.size());
        for (std::size_t i = 0; i < pts.size(); ++i) kps[i] = KeyPoint(pts[i], size, -1, response, octave, class_id);
    }
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

struct MatStep {                                   // cv::MatStep: step[0] = bytes per row, converts to size_t
    std::size_t p[2];
    MatStep(std::size_t s = 0) { p[0] = s; p[1] = 0; }
    operator std::size_t() const { return p[0]; }
    std::size_t& operator[](int i) { return p[i]; }
    const std::size_t& operator[](int i) const { return p[i]; }
    MatStep& operator=(std::size_t s) { p[0] = s; return *this; }
};

class Mat {
    std::shared_ptr<std::vector<uchar> > buf;     // null for headers over user memory
    int type_;
public:
    uchar* data;
    int rows, cols;
    MatStep step;                                  // bytes per row
    Mat() : type_(0), data(nullptr), rows(0), cols(0), step(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* p) : type_(type), data((uchar*)p), rows(r), cols(c), step((std::size_t)c * elemSize()) {}
    Mat(int r, int c, int type, void* p, std::size_t step_) : type_(type), data((uchar*)p), rows(r), cols(c), step(step_) {}
    // Mat(std::vector<Point2f>): an N x 1 CV_32FC2 header over the vector's memory (no copy)
    explicit Mat(const std::vector<Point2f>& v) : type_(CV_32FC2), data((uchar*)v.data()), rows((int)v.size()), cols(1), step(sizeof(Point2f)) {}
    Mat(const Mat& m, const Rect& roi) : buf(m.buf), type_(m.type_), data(m.data + (std::size_t)roi.y * m.step + (std::size_t)roi.x * m.elemSize()),
                                         rows(roi.height), cols(roi.width), step(m.step) {}
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = (std::size_t)c * elemSize();
        buf = std::make_shared<std::vector<uchar> >((std::size_t)r * step);
        data = buf->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }          // the vector is value-initialised
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    std::size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return (std::size_t)sz[depth()]; }
    std::size_t elemSize() const { return elemSize1() * (std::size_t)channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (std::size_t)cols * elemSize() || rows <= 1; }
    std::size_t total() const { return (std::size_t)rows * (std::size_t)cols; }
    Size size() const { return Size(cols, rows); }
    template <typename T> T& at(int i, int j) { return *(T*)(data + (std::size_t)i * step + (std::size_t)j * sizeof(T)); }
    template <typename T> const T& at(int i, int j) const { return *(const T*)(data + (std::size_t)i * step + (std::size_t)j * sizeof(T)); }
    // at(i): element i of a single-row or single-column (continuous) matrix
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }
    template <typename T> T* ptr(int i = 0, int j = 0) { return (T*)(data + (std::size_t)i * step) + j; }
    template <typename T> const T* ptr(int i = 0, int j = 0) const { return (const T*)(data + (std::size_t)i * step) + j; }
    Mat row(int i) const { Mat m(*this); m.data = data + (std::size_t)i * step; m.rows = 1; return m; }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int i = 0; i < rows; ++i) std::memcpy(m.data + (std::size_t)i * m.step, data + (std::size_t)i * step, (std::size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& o) const { o = clone(); }
    // push_back(Mat): append the rows of m (same type and width); an empty matrix adopts m's shape
    void push_back(const Mat& m) {
        if (m.empty()) return;
        if (empty()) { *this = m.clone(); return; }
        assert(m.type_ == type_ && m.cols == cols);
        Mat n(rows + m.rows, cols, type_);
        for (int i = 0; i < rows; ++i) std::memcpy(n.data + (std::size_t)i * n.step, data + (std::size_t)i * step, n.step);
        for (int i = 0; i < m.rows; ++i) std::memcpy(n.data + (std::size_t)(rows + i) * n.step, m.data + (std::size_t)i * m.step, n.step);
        *this = n;
    }
};
inline void swap(Mat& a, Mat& b) { std::swap(a, b); }

// MatExpr a - b evaluated into a Mat: cv::subtract, which SATURATES for CV_8U (saturate_cast<uchar>(a - b))
inline Mat operator-(const Mat& a, const Mat& b) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    Mat r(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int i = 0; i < a.rows; ++i) {
        if (a.depth() == CV_8U) {
            const uchar *pa = a.ptr<uchar>(i), *pb = b.ptr<uchar>(i); uchar* pr = r.ptr<uchar>(i);
            for (int k = 0; k < n; ++k) { int v = (int)pa[k] - (int)pb[k]; pr[k] = (uchar)(v < 0 ? 0 : v); }
        } else if (a.depth() == CV_32F) {
            const float *pa = a.ptr<float>(i), *pb = b.ptr<float>(i); float* pr = r.ptr<float>(i);
            for (int k = 0; k < n; ++k) pr[k] = pa[k] - pb[k];
        } else assert(!"shim: operator- for this depth");
    }
    return r;
}
// cv::norm(Mat, normType): NORM_HAMMING = number of set bits over all bytes (CV_8U only); NORM_L2 = sqrt(sum v^2) in double
inline double norm(const Mat& m, int normType = NORM_L2) {
    const int n = m.cols * m.channels();
    if (normType == NORM_HAMMING) {
        assert(m.depth() == CV_8U);
        int d = 0;
        for (int i = 0; i < m.rows; ++i) { const uchar* p = m.ptr<uchar>(i); for (int k = 0; k < n; ++k) d += __builtin_popcount((unsigned)p[k]); }
        return (double)d;
    }
    assert(normType == NORM_L2);
    double s = 0;
    for (int i = 0; i < m.rows; ++i)
        for (int k = 0; k < n; ++k) {
            double v = m.depth() == CV_32F ? (double)m.ptr<float>(i)[k] : m.depth() == CV_8U ? (double)m.ptr<uchar>(i)[k] : m.ptr<double>(i)[k];
            s += v * v;
        }
    return std::sqrt(s);
}

// imgproc undistortPoints(src, dst, cameraMatrix, distCoeffs) without R / P: normalised coordinates.
// cvUndistortPointsInternal as pinned against cv2 4.13: all in double, intrinsics and the coefficients
// (k1 k2 p1 p2 k3) widened from CV_32F, 5 iterations without an epsilon test, result stored as float.
inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& D) {
    const int n = (int)src.total();
    dst.create(n, 1, CV_32FC2);
    const double fx = K.at<float>(0, 0), fy = K.at<float>(1, 1), cx = K.at<float>(0, 2), cy = K.at<float>(1, 2);
    const double ifx = 1. / fx, ify = 1. / fy;
    double k[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5 && i < (int)D.total(); ++i) k[i] = D.at<float>(i);
    for (int i = 0; i < n; ++i) {
        const Point2f p = src.rows == n ? src.at<Point2f>(i, 0) : src.at<Point2f>(0, i);
        double x = ((double)p.x - cx) * ifx, y = ((double)p.y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = 1. / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        Vec2f& o = dst.at<Vec2f>(i, 0);
        o[0] = (float)x; o[1] = (float)y;
    }
}

// ---- features2d types that appear as members of MatcherOpenCV (include/putslam/Matcher/matcherOpenCV.h:71-73); only
// declared, never used: matcherOpenCV.cpp itself (ORB / SURF / SIFT creation) is not compiled into oracle/_ref
template <typename T> using Ptr = std::shared_ptr<T>;
class Feature2D { public: virtual ~Feature2D() {} };
typedef Feature2D FeatureDetector;
typedef Feature2D DescriptorExtractor;
class BFMatcher { public: explicit BFMatcher(int = NORM_L2, bool = false) {} };

// ---- display / colour helpers the reference calls only for debugging -----------------------------------------------
inline void cvtColor(const Mat& src, Mat& dst, int) { dst.create(src.rows, src.cols, CV_8UC1); }   // result never read (RGBD.cpp:151-152)
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
inline void drawKeypoints(const Mat&, const std::vector<KeyPoint>&, Mat&) {}
inline void drawMatches(const Mat&, const std::vector<KeyPoint>&, const Mat&, const std::vector<KeyPoint>&, const std::vector<DMatch>&, Mat&) {}
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) {}

}  // namespace cv
#endif
