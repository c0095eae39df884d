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
#include <functional>
#include "CBoWManager.h"
#include "CStereoSLAMKF.h"
#include "fundamental_ransac.h"

static const size_t INVALID_IDX = static_cast<size_t>(-1);

struct TLoopClosureInfo {                                                // srba-stereo-slam_utils.h:71-82
    std::vector<TKeyFrameID> similar_kfs; TKeyFrameID lc_id; std::vector<mrpt::poses::CPose3D> similar_kfs_poses;
    TLoopClosureInfo() : lc_id((TKeyFrameID)INVALID_KF_ID) {}
};

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
// EST.cpp:2015-2055: left-left fundamental matrix by RANSAC (fundamental_ransac.h = cv::findFundamentalMat FM_RANSAC),
// points1 = the OTHER keyframe's left keypoints, points2 = THIS keyframe's; non-inliers become sREJ_FUND_MATRIX.
void detectOutliersWithF(TVectorDAMatchInfo& this_matches, const size_t& num_tracked, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf,
                         double max_y_diff_epipolar, double ransac_fit_prob);
// EST.cpp:2113-2177: change in pose between the keyframes through voEngine.getChangeInPose (use_custom_initial_pose
// toggled around the call); pairs whose residual exceeds residual_th -- or every pair, if the solver reports an invalid
// result -- become sREJ_CHANGE_POSE.
void detectOutliersWithChangeInPose(rso::CStereoOdometryEstimator& voEngine, TVectorDAMatchInfo& this_matches, const CStereoSLAMKF& this_kf,
                                    const CStereoSLAMKF& other_kf, const mrpt::poses::CPose3DRotVec& kf_ini_rel_pose,
                                    const mrpt::utils::TStereoCamera& stereo_camera, double residual_th);
// CSRBAStereoSLAMEstimator::m_internal_data_association in full (EST.cpp:1413-1727): GPU brute-force match, the four
// filters in order under the srba_options switches, the ">= 15 inputs" guards (:1591, :1619) whose failure leaves the
// record empty, optional if_raw_match / if_match_after debug files (general_options.debug).
void internalDataAssociation(rso::CStereoOdometryEstimator& voEngine, const TSRBAStereoSLAMOptions& srba_options, const CStereoSLAMKF& this_kf,
                             const CStereoSLAMKF& other_kf, t_kf_da_info& out_da,
                             const mrpt::poses::CPose3DRotVec& kf_ini_rel_pose = mrpt::poses::CPose3DRotVec());
// The same from a ready brute-force result (what the GPU match returned): everything after EST.cpp:1452.
void internalDataAssociationFromMatches(rso::CStereoOdometryEstimator& voEngine, const TSRBAStereoSLAMOptions& srba_options, const CStereoSLAMKF& this_kf,
                                        const CStereoSLAMKF& other_kf, const std::vector<cv::DMatch>& bf_matches, t_kf_da_info& out_da,
                                        const mrpt::poses::CPose3DRotVec& kf_ini_rel_pose, std::vector<int>* status_out = nullptr);
// CSRBAStereoSLAMEstimator::m_data_association (EST.cpp:1341-1411): one record per similar keyframe, in list order.
void dataAssociation(rso::CStereoOdometryEstimator& voEngine, const TSRBAStereoSLAMOptions& srba_options, const t_vector_kf& keyframes,
                     const CStereoSLAMKF& kf, const TLoopClosureInfo& lc_info, TVectorKfsDaInfo& out_da);

// CSRBAStereoSLAMEstimator::m_get_similar_kfs (EST.cpp:1737-1878).  The SRBA state it reads is passed as two callbacks:
//   topo_distance(from, to)      spanning_tree.sym.next_edge[from][to].distance, or numeric_limits<size_t>::max() when
//                                `to` is not in `from`'s tree (:1801-1812); may throw when `from` has no tree (:1817)
//   tree_pose(root, id, out)     create_complete_spanning_tree(root, ..)[id].pose, false if absent (:1852-1857)
// Returns "potential loop closure found".
typedef std::function<size_t(TKeyFrameID, TKeyFrameID)> TTopoDistanceFn;
typedef std::function<bool(TKeyFrameID, TKeyFrameID, mrpt::poses::CPose3D&)> TTreePoseFn;
bool getSimilarKfs(const TKeyFrameID& newKfId, const QueryResults& dbQueryResults, const TSRBAStereoSLAMOptions& srba_options, size_t submap_size,
                   const mrpt::poses::CPose3DRotVec& incr_pose_from_last_kf, const TTopoDistanceFn& topo_distance, const TTreePoseFn& tree_pose,
                   TLoopClosureInfo& out);

// Round-1 form: the GPU match + the two cheap filters only.
void internalDataAssociation(sfe_context* ctx, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, bool filter_by_direction,
                             bool filter_by_orb_distance, double max_orb_distance_da, size_t image_rows, t_kf_da_info& out_da);
// EST.cpp:572-617: ids of the new keyframe's matches (tracked -> reuse, new -> ++last_match_id).  `sorted` lists
// the daInfo indices by decreasing tracked count (EST.cpp:461-470).  Returns {tracked, new}.
// Optional outputs: per_kf_tracked[i] = matches whose id came from daInfo[i] (vectorNumberTrackedFeats, :577, :598);
// da_dist[m] = the distance the da_dist debug file shows for match m (:602, :614), -1 for a new feature.
std::pair<size_t, size_t> assignMatchIds(CStereoSLAMKF& new_kf, const TVectorKfsDaInfo& daInfo, const std::vector<size_t>& sorted,
                                         const t_vector_kf& keyframes, size_t& last_match_id, std::vector<int>* per_kf_tracked = nullptr,
                                         std::vector<double>* da_dist = nullptr);

// The adaptive FAST / ORB threshold loop around processNewImagePair, verbatim from EST.cpp:263-315 (octave 0).
// Returns the number of processNewImagePair calls made (>= 1); throws std::runtime_error where the reference throws.
struct TAdaptiveThOptions { bool orb_adaptive_fast_th = true; int adaptive_th_min_matches = 100; int detect_fast_th = 20; };   // kitti.ini:113-115
int processWithAdaptiveThresholds(rso::CStereoOdometryEstimator& vo, rso::CStereoOdometryEstimator::TStereoOdometryRequest& request,
                                  rso::CStereoOdometryEstimator::TStereoOdometryResult& result, const TAdaptiveThOptions& opt);

struct TStereoCameraParams { double cul, cvl, fl, cur, fr, baseline; };
struct TObservation { size_t feat_id; float lx, ly, rx, ry; double X, Y, Z; };     // new_kf_observation_t: obs.feat_id, left_px, right_px, setRelPos
// UTL.h:558-574: the closed form on plain numbers ...
void projectMatchTo3D(double ul, double vl, double ur, const TStereoCameraParams& cam, double& X, double& Y, double& Z);
// ... and with the reference's signature.
mrpt::math::TPoint3D projectMatchTo3D(const double& ul, const double& vl, const double& ur, const mrpt::utils::TStereoCamera& stereoCamera);
// EST.cpp:135-161, 666-699: one observation per stereo match, landmark initialised at
// camera_pose_on_robot_rvt + projectMatchTo3D(...) (:158, :698).
void assembleObservations(const CStereoSLAMKF& kf, const TStereoCameraParams& cam, std::vector<TObservation>& out);      // camera frame only
void assembleObservations(const CStereoSLAMKF& kf, const TSRBAStereoSLAMOptions& srba_options, std::vector<TObservation>& out);
