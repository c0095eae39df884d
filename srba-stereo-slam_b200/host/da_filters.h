// da_filters.h -- host-side data-association steps immediately downstream of the GPU keyframe match
// (SURVEY.md 8(f) rows 1-2), restated bug-compatibly from the reference so that keyframe decisions match:
//   direction-histogram filter        CSRBAStereoSLAMEstimator.cpp:1883-1946 (m_detect_outliers_with_direction)
//   ORB distance + uniqueness filter   :1951-2010 (m_detect_outliers_with_orb_distance) -- INCLUDING the counter bug:
//       two commented-out lines make the "best distance so far" slot receive `++stage1_counter` (:2003-2007)
//   tracking-info record               :1657-1669 (t_kf_da_info, srba-stereo-slam_utils.h:55-66)
//   match-id assignment                :572-617
//   observation assembly               :666-699 + projectMatchTo3D (srba-stereo-slam_utils.h:558-574)
// Small N (a few hundred matches), sequential by nature: host code, no GPU.
#pragma once
#include <cstddef>
#include <limits>
#include <set>
#include <utility>
#include <vector>
#include "CStereoSLAMKF.h"

static const size_t INVALID_IDX = static_cast<size_t>(-1);

struct TVectorDAMatchInfo {                                              // EST.h:100-114
    enum TDAMatchStatus { sTRACKED = 0, sNON_TRACKED, sREJ_SLOPE, sREJ_ORB, sREJ_FUND_MATRIX, sREJ_CHANGE_POSE, sREJ_CONSISTENCY };
    std::vector<TDAMatchStatus> m_status;
    std::vector<cv::DMatch> m_matches;
    explicit TVectorDAMatchInfo(size_t s) { m_status.resize(s, sTRACKED); m_matches.resize(s); }
    size_t size() const { return m_status.size(); }
};

struct t_kf_da_info {                                                    // srba-stereo-slam_utils.h:55-66
    size_t kf_idx; size_t tracked_matches;
    std::vector<std::pair<size_t, double>> tracking_info;
    t_kf_da_info() : kf_idx(INVALID_IDX), tracked_matches(0) {}
};
typedef std::vector<t_kf_da_info> TVectorKfsDaInfo;

// EST.cpp:1883-1946.  `offset` is the image height (EST.cpp:1486).
void detectOutliersWithDirection(TVectorDAMatchInfo& this_matches, const size_t& offset, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf);
// EST.cpp:1951-2010.
void detectOutliersWithOrbDistance(TVectorDAMatchInfo& this_matches, double max_orb_distance_da, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf);
// EST.cpp:1427-1429 + 1657-1669: out_da from the statuses.
void fillTrackingInfo(const TVectorDAMatchInfo& this_matches, const CStereoSLAMKF& other_kf, t_kf_da_info& out_da);
// The GPU match + the two host filters in the reference's order (EST.cpp:1413-1515, 1644-1669); F-matrix and
// change-in-pose filters (:1589-1642) need OpenCV calib3d / the stereo-vo solver and are left to the host pipeline.
void internalDataAssociation(sfe_context* ctx, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, bool filter_by_direction,
                             bool filter_by_orb_distance, double max_orb_distance_da, size_t image_rows, t_kf_da_info& out_da);
// EST.cpp:572-617: ids of the new keyframe's matches (tracked -> reuse, new -> ++last_match_id).  `sorted` lists
// the daInfo indices by decreasing tracked count (EST.cpp:461-470).  Returns {tracked, new}.
std::pair<size_t, size_t> assignMatchIds(CStereoSLAMKF& new_kf, const TVectorKfsDaInfo& daInfo, const std::vector<size_t>& sorted,
                                         const t_vector_kf& keyframes, size_t& last_match_id);

// The adaptive FAST / ORB threshold loop around processNewImagePair, verbatim from EST.cpp:263-315 (octave 0).
// Returns the number of processNewImagePair calls made (>= 1); throws std::runtime_error where the reference throws.
struct TAdaptiveThOptions { bool orb_adaptive_fast_th = true; int adaptive_th_min_matches = 100; int detect_fast_th = 20; };   // kitti.ini:113-115
int processWithAdaptiveThresholds(rso::CStereoOdometryEstimator& vo, rso::CStereoOdometryEstimator::TStereoOdometryRequest& request,
                                  rso::CStereoOdometryEstimator::TStereoOdometryResult& result, const TAdaptiveThOptions& opt);

struct TStereoCameraParams { double cul, cvl, fl, cur, fr, baseline; };
struct TObservation { size_t feat_id; float lx, ly, rx, ry; double X, Y, Z; };
// UTL.h:558-574 (without the camera_pose_on_robot composition, which needs MRPT's CPose3DRotVec).
void projectMatchTo3D(double ul, double vl, double ur, const TStereoCameraParams& cam, double& X, double& Y, double& Z);
// EST.cpp:135-161, 666-699.
void assembleObservations(const CStereoSLAMKF& kf, const TStereoCameraParams& cam, std::vector<TObservation>& out);
