// kf_serialization.h -- the reference's keyframe stream formats, byte-compatible:
//   keypoints + descriptors   CSRBAStereoSLAMEstimator.cpp:2440-2484 (load), :2517-2560 (dump):
//       size_t n; n x {float x, y, response, size, angle; int octave, class_id}; int rows, cols, type; uchar data[rows*cols]
//   matches + ids             :2486-2512 (load), :2591-2616 (dump):
//       size_t n; n x {size_t id; int queryIdx, trainIdx; float distance; int imgIdx}
// Used to interchange keyframes (golden vectors, shards between ranks/runs).  Note the field ORDER differs from the
// in-memory cv::KeyPoint (response before size) and cv::DMatch (distance before imgIdx) -- kept as in the reference.
#pragma once
#include <fstream>
#include <iostream>
#include "CStereoSLAMKF.h"

bool dumpKeypointsToStream(std::ostream& stream, const TKeyPointList& keypoints, const cv::Mat& descriptors);
bool loadKeypointsFromStream(std::istream& stream, TKeyPointList& keypoints, cv::Mat& descriptors);
bool dumpMatchesToStream(std::ostream& stream, const TDMatchList& matches, const std::vector<size_t>& matches_ids);
bool loadMatchesFromStream(std::istream& stream, TDMatchList& matches, std::vector<size_t>& matches_ids);
// Options block of the state file, CSRBAStereoSLAMEstimator.cpp:2411-2438 (load), :2562-2589 (dump): 20 fields written
// with sizeof() each, in this order: n_levels, n_feats (size_t), min_ORB_distance (int), detect_method (enum),
// detect_fast_th (int), non_max_supp_method (enum), matching_options.epipolar_TH, max_y_diff_epipolar (double),
// matching_options.maxORB_dist, max_orb_distance_da, ransac_fit_prob (double), matching_options.{max_disp, min_disp,
// enable_robust_1to1_match}, residual_th, max_translation, max_rotation (double), non_maximal_suppression (bool),
// updated_matches_th, up_matches_th_plus (size_t).
bool dumpOptionsToStream(std::ostream& stream, const TSRBAStereoSLAMOptions& options);
bool loadOptionsFromStream(std::istream& stream, TSRBAStereoSLAMOptions& options);
// One keyframe = left keypoints, right keypoints, matches (the order m_save_state uses, :2633-2650 in the #if 0 block).
bool dumpKeyframeToStream(std::ostream& stream, const CStereoSLAMKF& kf);
bool loadKeyframeFromStream(std::istream& stream, CStereoSLAMKF& kf);
