// cv_compat.h -- the few OpenCV 2.4 types and calls the reference's hot-path callers use (cv::KeyPoint, cv::DMatch, cv::Mat of
// CV_8U descriptors / CV_32F point lists, cv::BFMatcher(NORM_HAMMING).match, cv::findFundamentalMat(FM_RANSAC)).  When the
// pipeline is built against real OpenCV define SRBA_B200_WITH_OPENCV and the real headers are used; the layouts are
// identical (28-byte KeyPoint, 16-byte DMatch; reference src/srba-stereo-slam_utils.cpp:49-50,75), which is what makes the
// GPU library's output buffers drop-in.  Without OpenCV:
//   cv::BFMatcher::match        -> the GPU brute-force matcher (sfe_match_bf) through a hook the engine installs
//                                  (reference call: CSRBAStereoSLAMEstimator.cpp:1447-1452)
//   cv::findFundamentalMat      -> host/fundamental_ransac.h (reference call: CSRBAStereoSLAMEstimator.cpp:2043)
// so the reference's data-association functions compile and run unchanged (the test-suite compiles them from the reference checkout and does exactly that).
#pragma once
#ifdef SRBA_B200_WITH_OPENCV
#include <opencv2/calib3d/calib3d.hpp>
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>
#include "fundamental_ransac.h"
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
typedef unsigned char uchar;
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
template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };
// Minimal ref-counted, single-channel, continuous 2-D matrix: what the reference touches (rows, cols, type(), row(i),
// copyTo, ptr, at<T>, create, element iteration).  row() returns a VIEW that shares the buffer, as in OpenCV.
class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_8U), ofs_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), type_(CV_8U), ofs_(0) { create(r, c, type); }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_8U ? 1 : (type_ == CV_32F ? 4 : 8); }
    bool empty() const { return rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    uint8_t* ptr(int r = 0) { return buf_ ? buf_->data() + ofs_ + (size_t)r * cols * elemSize() : nullptr; }
    const uint8_t* ptr(int r = 0) const { return buf_ ? buf_->data() + ofs_ + (size_t)r * cols * elemSize() : nullptr; }
    template <typename T> T& at(int r, int c) { return reinterpret_cast<T*>(ptr(r))[c]; }
    template <typename T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(ptr(r))[c]; }
    Mat row(int r) const { Mat v; v.rows = 1; v.cols = cols; v.type_ = type_; v.buf_ = buf_; v.ofs_ = ofs_ + (size_t)r * cols * elemSize(); return v; }
    // copies into `dst`'s storage when the sizes agree (the row(i).copyTo(m.row(k)) idiom, EST.cpp:1357-1358), else reallocates
    void copyTo(const Mat& dst) const
    {
        Mat& d = const_cast<Mat&>(dst);
        if (d.rows != rows || d.cols != cols || d.type_ != type_ || !d.buf_) d.create(rows, cols, type_);
        if (!empty()) std::memcpy(d.ptr(0), ptr(0), (size_t)rows * cols * elemSize());
    }
    void copyRowTo(int r, Mat& dst, int dr) const { std::memcpy(dst.ptr(dr), ptr(r), cols * elemSize()); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; ofs_ = 0;
        buf_.reset(new std::vector<uint8_t>((size_t)r * c * elemSize()));
    }
    template <typename T> T* begin() { return reinterpret_cast<T*>(ptr(0)); }
    template <typename T> T* end() { return reinterpret_cast<T*>(ptr(0)) + (size_t)rows * cols; }
    template <typename T> const T* begin() const { return reinterpret_cast<const T*>(ptr(0)); }
    template <typename T> const T* end() const { return reinterpret_cast<const T*>(ptr(0)) + (size_t)rows * cols; }
private:
    int type_; size_t ofs_;
    std::shared_ptr<std::vector<uint8_t>> buf_;
};
template <typename T> using MatIterator_ = T*;

// ---- brute-force matcher --------------------------------------------------------------------------------------------------
enum { NORM_HAMMING = 6 };
// (query, nq, train, nt, out[nq]) -> number of matches written (0 when the train set is empty); BFMatcher semantics:
// lowest train index on ties, queryIdx == row.  The CUDA-backed engine installs the GPU implementation.
typedef int (*TBruteForceHook)(const uint8_t* query, int nq, const uint8_t* train, int nt, DMatch* out);
inline TBruteForceHook& bruteForceHook() { static TBruteForceHook h = nullptr; return h; }
class BFMatcher {
public:
    BFMatcher(int normType = NORM_HAMMING, bool crossCheck = false)
    {
        if (normType != NORM_HAMMING || crossCheck) throw std::logic_error("cv::BFMatcher (b200): only NORM_HAMMING without cross check (EST.cpp:1447)");
    }
    void match(const Mat& queryDescriptors, const Mat& trainDescriptors, std::vector<DMatch>& matches) const
    {
        matches.clear();
        if (queryDescriptors.rows == 0 || trainDescriptors.rows == 0) return;
        if (queryDescriptors.cols != 32 || trainDescriptors.cols != 32 || queryDescriptors.type() != CV_8U)
            throw std::logic_error("cv::BFMatcher (b200): 256-bit CV_8U descriptors only");
        if (!bruteForceHook()) throw std::runtime_error("cv::BFMatcher (b200): no GPU context yet (there is no CPU fallback)");
        matches.resize(queryDescriptors.rows);
        const int n = bruteForceHook()(queryDescriptors.ptr(0), queryDescriptors.rows, trainDescriptors.ptr(0), trainDescriptors.rows, matches.data());
        if (n < 0) throw std::runtime_error("cv::BFMatcher (b200): the GPU matcher failed");
        matches.resize(n);
    }
};

// ---- fundamental matrix -----------------------------------------------------------------------------------------------------
enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 };
// points: N x 2 CV_32F (the reference's layout, EST.cpp:2021-2022).  Only FM_RANSAC with N >= 15 is implemented (the
// reference guards the call with that count, EST.cpp:1591); returns the 3x3 CV_64F matrix (empty if no model was found).
inline Mat findFundamentalMat(const Mat& points1, const Mat& points2, int method, double param1, double param2, std::vector<uchar>& mask)
{
    if (method != FM_RANSAC) throw std::logic_error("cv::findFundamentalMat (b200): FM_RANSAC only (EST.cpp:2043)");
    if (points1.rows != points2.rows || points1.cols != 2 || points2.cols != 2 || points1.type() != CV_32F || points2.type() != CV_32F)
        throw std::logic_error("cv::findFundamentalMat (b200): N x 2 CV_32F point lists expected");
    const int n = points1.rows;
    mask.assign(n, 0);
    double F[9];
    const int good = n > 0 ? findFundamentalMatRansac(points1.begin<float>(), points2.begin<float>(), n, param1, param2, mask.data(), F, nullptr) : 0;
    Mat out;
    if (good > 0) { out.create(3, 3, CV_64F); std::memcpy(out.ptr(0), F, sizeof(F)); }
    return out;
}
}  // namespace cv
#endif
static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint must be 28 bytes (sfe_keypoint)");
static_assert(sizeof(cv::DMatch) == 16, "cv::DMatch must be 16 bytes (sfe_dmatch)");
