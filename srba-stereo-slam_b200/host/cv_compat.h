// cv_compat.h -- the few OpenCV value types the reference's keyframe classes store (cv::KeyPoint, cv::DMatch,
// cv::Mat of CV_8U descriptors).  When the pipeline is built against real OpenCV define SRBA_B200_WITH_OPENCV and
// the real headers are used; the layouts are identical (28-byte KeyPoint, 16-byte DMatch; reference
// src/srba-stereo-slam_utils.cpp:49-50,75), which is what makes the GPU library's output buffers drop-in.
#pragma once
#ifdef SRBA_B200_WITH_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
namespace cv {
struct Point2f { float x, y; Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};
// Minimal ref-counted 2-D byte matrix (only what the reference touches: rows, cols, type(), row copies, at<uint8_t>).
class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0) {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c), buf_(new std::vector<uint8_t>((size_t)r * c)) {}
    int type() const { return CV_8U; }
    bool empty() const { return rows == 0 || cols == 0; }
    uint8_t* ptr(int r = 0) { return buf_ ? buf_->data() + (size_t)r * cols : nullptr; }
    const uint8_t* ptr(int r = 0) const { return buf_ ? buf_->data() + (size_t)r * cols : nullptr; }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(ptr(r))[c]; }
    template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(ptr(r))[c]; }
    void copyRowTo(int r, Mat& dst, int dr) const { std::memcpy(dst.ptr(dr), ptr(r), cols); }
    void create(int r, int c, int /*type*/) { rows = r; cols = c; buf_.reset(new std::vector<uint8_t>((size_t)r * c)); }
private:
    std::shared_ptr<std::vector<uint8_t>> buf_;
};
}  // namespace cv
#endif
static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint must be 28 bytes (sfe_keypoint)");
static_assert(sizeof(cv::DMatch) == 16, "cv::DMatch must be 16 bytes (sfe_dmatch)");
