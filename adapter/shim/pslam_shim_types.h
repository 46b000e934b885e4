// pslam_shim_types.h -- minimal, layout-compatible stand-ins for the few OpenCV / Eigen types that cross
// the PUTSLAM hot-path boundary.  This image has neither OpenCV C++ nor Eigen headers, so the adapter is
// compiled and tested against these; in a real PUTSLAM build define PSLAM_USE_REAL_HEADERS and the genuine
// headers are used instead (the adapter only touches the members declared here).
//
//   cv::DMatch        {int queryIdx, trainIdx, imgIdx; float distance}        (OpenCV core/types.hpp)
//   cv::Point2f       {float x, y}
//   cv::KeyPoint      {Point2f pt; float size, angle, response; int octave, class_id}
//   cv::Mat           rows, cols, data, step[0], isContinuous(), type(), channels(), ptr<T>(row), at<T>(r, c)
//   Eigen::Vector3f   3 packed floats;   Eigen::Matrix4f  16 floats, column-major
//   Eigen::MatrixXd   dynamic, column-major double, rows()/cols()/data()/operator()(r, c)
#pragma once
#ifdef PSLAM_USE_REAL_HEADERS
#include <Eigen/Dense>
#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
#else
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace cv {
enum { CV_8U_ = 0, CV_16U_ = 2, CV_32F_ = 5, CV_8UC3_ = 16 };   // OpenCV's CV_MAKETYPE values
#ifndef CV_8U
#define CV_8U cv::CV_8U_
#define CV_16U cv::CV_16U_
#define CV_32F cv::CV_32F_
#define CV_32FC1 cv::CV_32F_
#define CV_8UC1 cv::CV_8U_
#define CV_8UC3 cv::CV_8UC3_
#endif
struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1;
    float distance = 3.402823466e+38f;
    DMatch() {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int i, float d) : queryIdx(q), trainIdx(t), imgIdx(i), distance(d) {}
};
struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
// Reference-counted dense 2-D array, enough of cv::Mat for descriptor / depth / camera matrices.
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step0 = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t step = 0) : rows(r), cols(c), data((unsigned char*)ext), type_(type) {
        step0 = step ? step : (size_t)c * elemSize();
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type; step0 = (size_t)c * elemSize();
        store_ = std::shared_ptr<std::vector<unsigned char>>(new std::vector<unsigned char>(step0 * (size_t)r));
        data = store_->data();
    }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_8U_ ? 1 : (type_ == CV_16U_ ? 2 : (type_ == CV_8UC3_ ? 3 : 4)); }
    int channels() const { return type_ == CV_8UC3_ ? 3 : 1; }
    bool isContinuous() const { return step0 == (size_t)cols * elemSize(); }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + step0 * (size_t)r); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + step0 * (size_t)r); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
private:
    int type_ = 0;
    std::shared_ptr<std::vector<unsigned char>> store_;
};
}  // namespace cv

namespace Eigen {
struct Vector3f {
    float v[3] = {0, 0, 0};
    Vector3f() {}
    Vector3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
    const float* data() const { return v; }
    float* data() { return v; }
};
struct Matrix4f {  // column-major like Eigen's default
    float m[16];
    Matrix4f() { std::memset(m, 0, sizeof(m)); }
    static Matrix4f Identity() { Matrix4f r; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
    float& operator()(int r, int c) { return m[4 * c + r]; }
    const float& operator()(int r, int c) const { return m[4 * c + r]; }
    float* data() { return m; }
    const float* data() const { return m; }
};
class MatrixXd {  // column-major dynamic double matrix
public:
    MatrixXd() {}
    MatrixXd(long r, long c) : r_(r), c_(c), d_((size_t)(r * c), 0.0) {}
    long rows() const { return r_; }
    long cols() const { return c_; }
    double& operator()(long r, long c) { return d_[(size_t)(c * r_ + r)]; }
    const double& operator()(long r, long c) const { return d_[(size_t)(c * r_ + r)]; }
    const double* data() const { return d_.data(); }
    double* data() { return d_.data(); }
private:
    long r_ = 0, c_ = 0;
    std::vector<double> d_;
};
}  // namespace Eigen
static_assert(sizeof(cv::DMatch) == 16, "cv::DMatch layout");
static_assert(sizeof(Eigen::Vector3f) == 12, "Eigen::Vector3f layout");
static_assert(sizeof(Eigen::Matrix4f) == 64, "Eigen::Matrix4f layout");
#endif  // PSLAM_USE_REAL_HEADERS
