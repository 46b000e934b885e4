// Stand-in for the reference's include/putslam/Defs/opencv.h (which pulls in all of OpenCV), used ONLY to compile the
// reference's own src/Matcher/dbscan.cpp into oracle/_ref/ (see oracle/Makefile, target ref).  OpenCV's C++ headers do
// not exist in this image; dbscan.cpp needs exactly three things from them, declared here with OpenCV's semantics:
//   cv::Point2f with operator-            (core/types.hpp: component-wise float subtraction)
//   cv::norm(const Point2f&) -> double     (core/types.hpp: std::sqrt((double)pt.x*pt.x + (double)pt.y*pt.y))
//   cv::KeyPoint {pt, size, angle, response, octave, class_id}
// TEST INFRASTRUCTURE: nothing in the product includes this file.
#ifndef PSLAM_REF_SHIM_OPENCV_H
#define PSLAM_REF_SHIM_OPENCV_H
#include <algorithm>   // OpenCV's core headers pull these in; dbscan.cpp relies on it for std::remove_if
#include <cmath>
namespace cv {
struct Point2f {
    float x, y;
    Point2f() : x(0), y(0) {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
inline Point2f operator-(const Point2f& a, const Point2f& b) { return Point2f(a.x - b.x, a.y - b.y); }
inline double norm(const Point2f& pt) { return std::sqrt((double)pt.x * pt.x + (double)pt.y * pt.y); }
struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
}  // namespace cv
#endif
