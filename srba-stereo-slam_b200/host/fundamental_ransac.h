// fundamental_ransac.h -- cv::findFundamentalMat(points1, points2, FM_RANSAC, param1, param2, mask) as OpenCV 2.4 runs it,
// for the data-association filter CSRBAStereoSLAMEstimator::m_detect_outliers_with_F (reference
// CSRBAStereoSLAMEstimator.cpp:2015-2055, call at :2043 with param1 = max_y_diff_epipolar, param2 = ransac_fit_prob).
// OpenCV is not vendored by the reference; this restates the published 2.4 algorithm (modules/calib3d/src/fundam.cpp,
// modelest.cpp):
//   * RANSAC over 7-point minimal sets, at most 2000 iterations, the iteration count re-estimated from the inlier ratio
//     (cvRANSACUpdateNumIters), NO final refit (2.4 has it commented out), needs >= 15 points (else the caller skips it,
//     EST.cpp:1591);
//   * sampling by cv::RNG -- multiply-with-carry, state = state*4164903690 + (state >> 32), seeded with (uint64)-1 for
//     every call -- `rng % count` with duplicate rejection and the "no three collinear points" partial-subset check;
//   * 7-point solver on raw pixel coordinates: null space of the 7x9 epipolar system, det(l*F1 + (1-l)*F2) = 0 solved as
//     a cubic (trigonometric / Cardano form), each real root scaled so that F33 = 1;
//   * error = max of the two squared point-to-epipolar-line distances (stored as float), inlier iff err <= param1^2;
//     a model replaces the best one only if it has strictly more inliers (and >= 7).
// Parity: the inlier masks are pinned statistically against cv2 4.13's findFundamentalMat (same RNG, same minimal
// solver family) in tests/test_host_geometry.py; bit-level RNG/SVD parity with a real 2.4 build is unpinned (the null-
// space basis OpenCV's SVD returns orders the <= 3 cubic roots, which only matters when two roots tie on inlier count).
#pragma once
#include <cstdint>

// cv::RNG (OpenCV core): the generator behind cvRandInt / RNG::next.
struct CvRngMwc {
    uint64_t state;
    explicit CvRngMwc(uint64_t seed = 0xffffffffffffffffull) : state(seed ? seed : 0xffffffffffffffffull) {}
    uint32_t next() { state = (uint64_t)(uint32_t)state * 4164903690u + (state >> 32); return (uint32_t)state; }
};

// Up to 3 fundamental matrices (row-major, 9 doubles each) through 7 correspondences; returns their number.
int fundamental7Point(const double* m1_xy, const double* m2_xy, double* F_out /* 27 */);

// points: n x 2 floats (x, y) each.  mask[n] receives 1 for inliers.  Returns the inlier count of the winning model
// (0: no model found; the mask is then all zeros, which the reference's update loop reads as "everything rejected").
int findFundamentalMatRansac(const float* points1_xy, const float* points2_xy, int n, double param1, double param2,
                             uint8_t* mask, double* F_out /* 9, may be NULL */, int* iterations_out /* may be NULL */);
// The same with the OpenCV generation selectable, as sfe_config.pyramid_mode / blur_mode do for the extraction:
//   SRBA_FM_CV24  what the reference links (default above): <= 2000 iterations, partial-subset collinearity check while
//                 drawing (300 attempts);
//   SRBA_FM_CV4   OpenCV >= 3: <= 1000 iterations, whole-sample check after drawing (10000 attempts).  Bit-identical inlier
//                 masks to cv2 4.13 on every golden case -- this is what pins the shared machinery (RNG, 7-point solver,
//                 error, iteration update); the two generations differ only in the lines named here.
enum { SRBA_FM_CV24 = 0, SRBA_FM_CV4 = 1 };
int findFundamentalMatRansacEx(const float* points1_xy, const float* points2_xy, int n, double param1, double param2,
                               uint8_t* mask, double* F_out, int* iterations_out, int generation);
