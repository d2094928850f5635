// cv_compat.h -- the few OpenCV / Eigen types that appear in the reference's FeatureExtractor / FeatureMatcher
// signatures, as layout-compatible PODs, so the host mirror builds without OpenCV.  Define AFV_USE_OPENCV to use
// the real cv:: types instead (same memory layout: cv::KeyPoint is 28 bytes {pt.x, pt.y, size, angle, response,
// octave, class_id}; descriptors are continuous row-major CV_8U / CV_32F).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef AFV_USE_OPENCV
#include <opencv2/core.hpp>
namespace afvcv = cv;
#else
namespace afvcv {
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");
enum { CV_8U = 0, CV_32F = 5 };
// minimal continuous row-major matrix (what the reference passes around as cv::Mat descriptors / gray images)
struct Mat {
    int rows = 0, cols = 0, type_ = CV_8U;
    std::vector<uint8_t> buf;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    void create(int r, int c, int t) { rows = r; cols = c; type_ = t; buf.assign((size_t)r * c * elemSize(), 0); }
    void release() { rows = cols = 0; buf.clear(); }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
    size_t step() const { return (size_t)cols * elemSize(); }
    bool empty() const { return rows == 0 || cols == 0; }
    uint8_t* data() { return buf.data(); }
    const uint8_t* data() const { return buf.data(); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(buf.data() + (size_t)r * step()); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(buf.data() + (size_t)r * step()); }
};
}  // namespace afvcv
#endif
namespace afv_host {
struct mat2f { float m[4]; static mat2f scaledIdentity(float s) { mat2f r; r.m[0] = s; r.m[1] = 0; r.m[2] = 0; r.m[3] = s; return r; } };  // Eigen::Matrix2f stand-in
}
