#!/usr/bin/env python
"""oracle/_ref: the reference's OWN host-side data-association code, compiled here from where it lies.

TEST INFRASTRUCTURE ONLY.  The hot-path arithmetic of famoreno/srba-stereo-slam lives in un-vendored libraries (OpenCV 2.4,
stereo-vo, DBoW2) and cannot be built; but everything the reference itself wrote downstream of the keyframe match IS in
/root/reference/src/CSRBAStereoSLAMEstimator.cpp and compiles from that one file once the few external types it touches
exist.  This recipe takes the functions below, by line range, UNCHANGED from the reference source (nothing is copied into
the repository: the generated translation unit goes to oracle/_ref/, which is git-ignored), wraps them in a class that has
the members they use, and builds oracle/_ref/libref_da.so:

  m_data_association                      CSRBAStereoSLAMEstimator.cpp:1341-1411
  m_internal_data_association             :1413-1727
  m_get_similar_kfs                       :1737-1878   (over a stub of the SRBA state it reads)
  m_detect_outliers_with_direction        :1883-1946
  m_detect_outliers_with_orb_distance     :1951-2010   (the `= ++stage1_counter` bug included, naturally)
  m_detect_outliers_with_F                :2015-2055
  m_detect_outliers_with_change_in_pose   :2113-2177
  match-id assignment block               :572-617     (inside performStereoSLAM; wrapped with the locals it uses)
  observation-list block                  :666-704     (idem) + projectMatchTo3D, srba-stereo-slam_utils.h:558-574
  updateTranslation/RotationThreshold     srba-stereo-slam_utils.cpp:153-171
  + the call sites of the stereo front-end:  initialize :1102-1144, :1163-1167 (vocabulary, camera pose, options, engine
    parameters, forced modes), first frame :111-118, the adaptive-threshold loop and pose accumulation :258-330, candidate
    keyframe + BoW query :413-427 -- wrapped as member functions; they RUN on the GPU in tests/test_gpu_sequence.py.

They compile against the PRODUCT's drop-in headers (srba-stereo-slam_b200/host: cv_compat.h, mrpt_lite.h, CStereoSLAMKF.h,
stereo_odometry_b200.h, srba_options.h) -- which is also the compile-time proof that those headers are source compatible
with the reference -- with cv::BFMatcher backed by a plain CPU loop and cv::findFundamentalMat / getChangeInPose by the
product's host implementations (their own pins: tests/test_host_geometry.py).  tests/test_ref_da.py then feeds the same
keyframes to these functions and to the product's (host/da_filters.cpp, stereo_slam_frontend.cpp) and diffs the results.

Usage: python oracle/ref_build.py   (no-op with a message when /root/reference is absent, e.g. on the GPU box, where the
prebuilt .so that travelled with the snapshot is used).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
HOST = os.path.join(ROOT, "srba-stereo-slam_b200", "host")

# (file, first line, last line, text the first line must contain)
RANGES = {
    "data_association": ("CSRBAStereoSLAMEstimator.cpp", 1341, 1411, "void CSRBAStereoSLAMEstimator::m_data_association("),
    "internal_da": ("CSRBAStereoSLAMEstimator.cpp", 1413, 1727, "void CSRBAStereoSLAMEstimator::m_internal_data_association("),
    "similar_kfs": ("CSRBAStereoSLAMEstimator.cpp", 1737, 1878, "bool CSRBAStereoSLAMEstimator::m_get_similar_kfs("),
    "direction": ("CSRBAStereoSLAMEstimator.cpp", 1883, 1946, "void CSRBAStereoSLAMEstimator::m_detect_outliers_with_direction"),
    "orb_distance": ("CSRBAStereoSLAMEstimator.cpp", 1951, 2010, "void CSRBAStereoSLAMEstimator::m_detect_outliers_with_orb_distance"),
    "fund": ("CSRBAStereoSLAMEstimator.cpp", 2015, 2055, "void CSRBAStereoSLAMEstimator::m_detect_outliers_with_F"),
    "pose": ("CSRBAStereoSLAMEstimator.cpp", 2113, 2177, "void CSRBAStereoSLAMEstimator::m_detect_outliers_with_change_in_pose"),
    "ids_block": ("CSRBAStereoSLAMEstimator.cpp", 572, 617, "vector<int> vectorNumberTrackedFeats( numberSimilarKfs, 0 );"),
    "obs_block": ("CSRBAStereoSLAMEstimator.cpp", 666, 704, "mySRBA::TNewKeyFrameInfo       newKFInfo;"),
    # the call sites of the stereo front-end itself: compile-time conformance + a GPU run of the reference's loop
    "init_a": ("CSRBAStereoSLAMEstimator.cpp", 1102, 1144, "m_bow_manager.loadVocabularyFromConfigFile(config, \"SRBA_GENERAL\", \"voc_filename\" );"),
    "init_b": ("CSRBAStereoSLAMEstimator.cpp", 1163, 1167, "m_max_rotation_limit = srba_options.max_rotation"),
    "first_a": ("CSRBAStereoSLAMEstimator.cpp", 111, 115, "m_odom_request.stereo_imgs = stImgs;"),
    "first_b": ("CSRBAStereoSLAMEstimator.cpp", 118, 118, "m_bow_manager.insertIntoDB( new_kf );"),
    "next": ("CSRBAStereoSLAMEstimator.cpp", 258, 330, "m_odom_request.stereo_imgs            = stImgs;"),
    "candidate": ("CSRBAStereoSLAMEstimator.cpp", 413, 427, "m_keyframes.push_back( CStereoSLAMKF() );"),
    "project": ("srba-stereo-slam_utils.h", 558, 574, "inline mrpt::math::TPoint3D projectMatchTo3D("),
    "thresholds": ("srba-stereo-slam_utils.cpp", 153, 171, "double updateTranslationThreshold("),
}

PRELUDE = r'''
// GENERATED by oracle/ref_build.py -- do not commit.  Reference functions below are verbatim line ranges of /root/reference/src.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <deque>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "CBoWManager.h"
#include "CStereoSLAMKF.h"
#include "da_filters.h"
#include "srba_options.h"
using namespace std;
using namespace mrpt;
using namespace mrpt::poses;
using namespace cv;

// ---- what the reference's headers / libraries would have provided --------------------------------------------------------
namespace mrpt { namespace system { namespace os {
inline FILE* fopen(const std::string& n, const char* m) { return ::fopen(n.c_str(), m); }
inline FILE* fopen(const char* n, const char* m) { return ::fopen(n, m); }
using ::fprintf; using ::fclose;
} } }
#define GENERATE_NAME_WITH_2KF_OUT(_STR,_ID1,_ID2) mrpt::format("%s/%s_kf%04d_with_kf%04d.txt", general_options.out_dir.c_str(), #_STR, (int)(_ID1), (int)(_ID2))
#define DUMP_VECTORLIKE(_v) if( _v.size() > 0 ) { for(size_t k = 0; k < _v.size(); ++k) cout << #_v << "[" << k << "] = " << _v[k] << ", "; cout << endl; }
typedef vector< pair<int,float> > t_vector_pair_idx_distance;                 // srba-stereo-slam_utils.h:509
typedef uint64_t topo_dist_t; typedef size_t TLandmarkID;
struct TSpanTreeEntry { TKeyFrameID next; topo_dist_t distance; };
struct TPixelCoordf { float x, y; TPixelCoordf(float x_ = 0, float y_ = 0) : x(x_), y(y_) {} };
// the slice of srba::RbaEngine the extracted functions read (srba-stereo-slam.h:30-54)
struct mySRBA {
    struct rba_problem_state_t {
        struct TSpanningTree { typedef std::map<TKeyFrameID, std::map<TKeyFrameID, TSpanTreeEntry> > next_edge_maps_t; struct { next_edge_maps_t next_edge; } sym; } spanning_tree;
    };
    struct pose_flag_t { mrpt::poses::CPose3D pose; };
    typedef std::map<TKeyFrameID, pose_flag_t> frameid2pose_map_t;
    struct { struct { size_t submap_size; } ecp; struct { size_t max_tree_depth; } srba; } parameters;
    struct TNewKeyFrameInfo { TKeyFrameID kf_id; };
    struct new_kf_observation_t {
        bool is_fixed, is_unknown_with_init_val; double rel_pos[3];
        struct { TLandmarkID feat_id; struct { TPixelCoordf left_px, right_px; } obs_data; } obs;
        void setRelPos(const mrpt::math::TPoint3D& p) { rel_pos[0] = p.x; rel_pos[1] = p.y; rel_pos[2] = p.z; }
    };
    typedef std::deque<new_kf_observation_t> new_kf_observations_t;
    rba_problem_state_t state; std::map<TKeyFrameID, frameid2pose_map_t> trees;
    rba_problem_state_t& get_rba_state() { return state; }
    struct profiler_t { void enter(const char*) {} void leave(const char*) {} } prof; profiler_t& get_time_profiler() { return prof; }
    void create_complete_spanning_tree(TKeyFrameID root, frameid2pose_map_t& out, size_t) { out = trees[root]; }
};

@@PROJECT@@

@@THRESHOLDS@@

namespace refda {
class CSRBAStereoSLAMEstimator {
public:
    typedef ::TVectorDAMatchInfo TVectorDAMatchInfo;
    TSRBAStereoSLAMOptions srba_options; mySRBA rba;
    t_vector_kf m_keyframes; size_t m_last_match_ID = 0;
    rso::CStereoOdometryEstimator m_voEngine;
    rso::CStereoOdometryEstimator::TStereoOdometryRequest m_odom_request; rso::CStereoOdometryEstimator::TStereoOdometryResult m_odom_result;
    CBoWManager m_bow_manager; size_t m_last_kf_ID = 0;
    double m_max_rotation, m_max_translation, m_max_rotation_limit, m_max_translation_limit;
    CPose3DRotVec m_incr_pose_from_last_kf, m_current_pose, m_incr_pose_from_last_check;
    void initialize(const mrpt::utils::CConfigFile& config);
    void ref_first_frame(mrpt::obs::CObservationStereoImagesPtr stImgs);
    bool ref_next_frame(mrpt::obs::CObservationStereoImagesPtr stImgs, size_t& count);
    void ref_candidate(QueryResults& qResults_o);
    bool m_get_similar_kfs(const TKeyFrameID& newKfId, const QueryResults& dbQueryResults, TLoopClosureInfo& out);
    void m_data_association(const CStereoSLAMKF& kf, const TLoopClosureInfo& lc_info, TVectorKfsDaInfo& out_da);
    void m_internal_data_association(const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, const Mat& curLDesc, const Mat& curRDesc,
                                     t_kf_da_info& out_da, const CPose3DRotVec& kf_ini_rel_pose = CPose3DRotVec());
    void m_detect_outliers_with_change_in_pose(TVectorDAMatchInfo& this_matches, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, const CPose3DRotVec& kf_ini_rel_pose);
    void m_detect_outliers_with_F(TVectorDAMatchInfo& this_matches, const size_t& num_tracked, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf);
    void m_detect_outliers_with_direction(TVectorDAMatchInfo& this_matches, const size_t& offset, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf);
    void m_detect_outliers_with_orb_distance(TVectorDAMatchInfo& this_matches, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf);
    // the two inline blocks of performStereoSLAM, wrapped with the locals they use
    void ref_assign_ids(CStereoSLAMKF& new_kf, TVectorKfsDaInfo& daInfo, vector<size_t>& sortedIndices, size_t& numberNewFeats_o, size_t& numberTrackedFeats_o, vector<int>& perKf_o);
    void ref_observations(CStereoSLAMKF& new_kf, mySRBA::new_kf_observations_t& listObs_o);
};

@@FUNCTIONS@@

void CSRBAStereoSLAMEstimator::initialize( const mrpt::utils::CConfigFile & config )
{
    // ---- CSRBAStereoSLAMEstimator.cpp:1102-1144, 1163-1167 ----
@@INIT_A@@
@@INIT_B@@
}
void CSRBAStereoSLAMEstimator::ref_first_frame(mrpt::obs::CObservationStereoImagesPtr stImgs)
{
    m_keyframes.push_back( CStereoSLAMKF() ); CStereoSLAMKF & new_kf = m_keyframes.back(); new_kf.setKFID( m_last_kf_ID++ );      // :84-86
    // ---- CSRBAStereoSLAMEstimator.cpp:111-115, 118 ----
@@FIRST_A@@
@@FIRST_B@@
}
bool CSRBAStereoSLAMEstimator::ref_next_frame(mrpt::obs::CObservationStereoImagesPtr stImgs, size_t& count)
{
    bool skipped = true;
    do {
    // ---- CSRBAStereoSLAMEstimator.cpp:258-330 (the `continue` of :322 leaves this do/while) ----
@@NEXT@@
    skipped = false;
    } while (0);
    return !skipped;
}
void CSRBAStereoSLAMEstimator::ref_candidate(QueryResults& qResults_o)
{
    // ---- CSRBAStereoSLAMEstimator.cpp:413-427 ----
@@CANDIDATE@@
    qResults_o = qResults; (void)nMatches;
}

void CSRBAStereoSLAMEstimator::ref_assign_ids(CStereoSLAMKF& new_kf, TVectorKfsDaInfo& daInfo, vector<size_t>& sortedIndices, size_t& numberNewFeats_o, size_t& numberTrackedFeats_o, vector<int>& perKf_o)
{
    const size_t nMatches = new_kf.m_matches.size(); const size_t numberSimilarKfs = daInfo.size(); FILE* f = NULL;
    { // ---- CSRBAStereoSLAMEstimator.cpp:572-617 ----
@@IDS_BLOCK@@
    numberNewFeats_o = numberNewFeats; numberTrackedFeats_o = numberTrackedFeats; perKf_o = vectorNumberTrackedFeats;
    }
}
void CSRBAStereoSLAMEstimator::ref_observations(CStereoSLAMKF& new_kf, mySRBA::new_kf_observations_t& listObs_o)
{
    const size_t nMatches = new_kf.m_matches.size(); const bool confirmedLoopClosure = false; const size_t count = 0; t_vector_kf& m_keyframes_ = m_keyframes; (void)m_keyframes_; (void)count;
    struct { struct { unsigned Id; } a[1]; auto& operator[](int) { return a[0]; } } qResults; qResults.a[0].Id = 0; (void)confirmedLoopClosure;
    // ---- CSRBAStereoSLAMEstimator.cpp:666-704 ----
@@OBS_BLOCK@@
    }
    listObs_o = listObs;
}
}  // namespace refda

// ---- C shims ---------------------------------------------------------------------------------------------------------------
static int cpu_brute_force(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out)
{
    for (int i = 0; i < nq; ++i) {
        int best = 1 << 30, bj = -1;
        for (int j = 0; j < nt; ++j) {
            int d = 0; for (int k = 0; k < 32; ++k) d += __builtin_popcount(q[32 * i + k] ^ t[32 * j + k]);
            if (d < best) { best = d; bj = j; }
        }
        out[i] = cv::DMatch(i, bj, (float)best); out[i].imgIdx = 0;
    }
    return nq;
}

extern "C" {
struct ref_kf { int n_left, n_right, n_matches; const cv::KeyPoint* kl; const cv::KeyPoint* kr; const uint8_t* dl; const uint8_t* dr; const cv::DMatch* m; const long long* ids; long long kf_id; };
static void fill_kf(CStereoSLAMKF& k, const ref_kf& r)
{
    k.m_keypoints_left.assign(r.kl, r.kl + r.n_left); k.m_keypoints_right.assign(r.kr, r.kr + r.n_right);
    k.m_descriptors_left.create(r.n_left, 32, CV_8U); k.m_descriptors_right.create(r.n_right, 32, CV_8U);
    if (r.n_left) memcpy(k.m_descriptors_left.ptr(0), r.dl, (size_t)r.n_left * 32);
    if (r.n_right) memcpy(k.m_descriptors_right.ptr(0), r.dr, (size_t)r.n_right * 32);
    k.m_matches.assign(r.m, r.m + r.n_matches); k.m_matches_ID.clear();
    for (int i = 0; i < r.n_matches; ++i) k.m_matches_ID.push_back((size_t)r.ids[i]);
    k.setKFID((size_t)r.kf_id);
}
void* ref_create(const char* ini_file)
{
    cv::bruteForceHook() = &cpu_brute_force;
    refda::CSRBAStereoSLAMEstimator* e = new refda::CSRBAStereoSLAMEstimator();
    mrpt::utils::CConfigFile cfg(ini_file);
    e->srba_options.loadFromConfigFile(cfg);
    std::vector<std::string> s = { "RECTIFY", "DETECT", "MATCH", "IF-MATCH", "LEAST_SQUARES", "GUI", "GENERAL" };
    e->m_voEngine.loadParamsFromConfigFile(cfg, s);
    const CPose3D img_to_camera_pose(0, 0, 0, DEG2RAD(-90), 0, DEG2RAD(-90));
    e->srba_options.camera_pose_on_robot_rvt = CPose3DRotVec(img_to_camera_pose);
    e->srba_options.camera_pose_on_robot_rvt_inverse = e->srba_options.camera_pose_on_robot_rvt.getInverse();
    return e;
}
void ref_destroy(void* h) { delete (refda::CSRBAStereoSLAMEstimator*)h; }
void ref_set_keyframes(void* h, const ref_kf* kfs, int n)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    e->m_keyframes.assign(n, CStereoSLAMKF());
    for (int i = 0; i < n; ++i) fill_kf(e->m_keyframes[i], kfs[i]);
}
// m_internal_data_association(this = keyframes[this_idx], other = keyframes[other_idx]); out: first[n] (-1 = INVALID_IDX), dist[n]
int ref_internal_da(void* h, int this_idx, int other_idx, const double* ini_pose6 /* w, t */, long long* first, double* dist)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    const CStereoSLAMKF& kf = e->m_keyframes[this_idx];
    const size_t n = kf.m_matches.size();
    cv::Mat curLDesc((int)n, 32, CV_8U), curRDesc((int)n, 32, CV_8U);
    for (size_t k = 0; k < n; ++k) {                                             // EST.cpp:1354-1359
        kf.m_descriptors_left.row(kf.m_matches[k].queryIdx).copyTo(curLDesc.row(k));
        kf.m_descriptors_right.row(kf.m_matches[k].trainIdx).copyTo(curRDesc.row(k));
    }
    t_kf_da_info da;
    const CPose3DRotVec ini(ini_pose6[0], ini_pose6[1], ini_pose6[2], ini_pose6[3], ini_pose6[4], ini_pose6[5]);
    e->m_internal_data_association(kf, e->m_keyframes[other_idx], curLDesc, curRDesc, da, ini);
    for (size_t k = 0; k < n; ++k) { first[k] = da.tracking_info[k].first == INVALID_IDX ? -1 : (long long)da.tracking_info[k].first; dist[k] = da.tracking_info[k].second; }
    return (int)da.tracked_matches;
}
// the filters one by one on a caller-provided status / match list (status codes = TVectorDAMatchInfo::TDAMatchStatus)
void ref_filter(void* h, int which, int this_idx, int other_idx, const cv::DMatch* bf, int* status, const double* ini_pose6)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    const CStereoSLAMKF& a = e->m_keyframes[this_idx]; const CStereoSLAMKF& b = e->m_keyframes[other_idx];
    const size_t n = a.m_matches.size();
    TVectorDAMatchInfo m(n);
    for (size_t k = 0; k < n; ++k) { m.m_matches[k] = bf[k]; m.m_status[k] = (TVectorDAMatchInfo::TDAMatchStatus)status[k]; }
    size_t tracked = 0; for (size_t k = 0; k < n; ++k) tracked += m.m_status[k] == TVectorDAMatchInfo::sTRACKED;
    if (which == 0) e->m_detect_outliers_with_direction(m, e->srba_options.stereo_camera.leftCamera.nrows, a, b);
    else if (which == 1) e->m_detect_outliers_with_orb_distance(m, a, b);
    else if (which == 2) e->m_detect_outliers_with_F(m, tracked, a, b);
    else e->m_detect_outliers_with_change_in_pose(m, a, b, CPose3DRotVec(ini_pose6[0], ini_pose6[1], ini_pose6[2], ini_pose6[3], ini_pose6[4], ini_pose6[5]));
    for (size_t k = 0; k < n; ++k) status[k] = (int)m.m_status[k];
}
// ids: daInfo given as first[nkf][nmatches] (-1 invalid) + kf_idx[nkf]; new_kf = keyframes.back()
void ref_assign_ids(void* h, int nkf, const long long* first, const int* da_kf_idx, const int* sorted, long long* last_match_id, long long* out_ids, long long* counts2, int* per_kf)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    CStereoSLAMKF& new_kf = e->m_keyframes.back();
    const size_t n = new_kf.m_matches.size();
    TVectorKfsDaInfo da(nkf);
    for (int k = 0; k < nkf; ++k) {
        da[k].kf_idx = da_kf_idx[k]; da[k].tracking_info.resize(n);
        for (size_t m = 0; m < n; ++m) { const long long f = first[(size_t)k * n + m]; da[k].tracking_info[m] = make_pair(f < 0 ? INVALID_IDX : (size_t)f, 0.0); }
    }
    vector<size_t> srt(sorted, sorted + nkf); vector<int> pk; size_t nn = 0, nt = 0;
    e->m_last_match_ID = (size_t)*last_match_id;
    new_kf.m_matches_ID.resize(n);
    e->ref_assign_ids(new_kf, da, srt, nn, nt, pk);
    for (size_t m = 0; m < n; ++m) out_ids[m] = (long long)new_kf.m_matches_ID[m];
    *last_match_id = (long long)e->m_last_match_ID; counts2[0] = (long long)nt; counts2[1] = (long long)nn;
    for (int k = 0; k < nkf; ++k) per_kf[k] = pk[k];
}
// observation list of keyframes[idx]: out[n][8] = id, lx, ly, rx, ry, X, Y, Z
void ref_observations(void* h, int idx, double* out)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    mySRBA::new_kf_observations_t l;
    e->ref_observations(e->m_keyframes[idx], l);
    for (size_t i = 0; i < l.size(); ++i) {
        double* o = out + 8 * i;
        o[0] = (double)l[i].obs.feat_id; o[1] = l[i].obs.obs_data.left_px.x; o[2] = l[i].obs.obs_data.left_px.y; o[3] = l[i].obs.obs_data.right_px.x; o[4] = l[i].obs.obs_data.right_px.y;
        o[5] = l[i].rel_pos[0]; o[6] = l[i].rel_pos[1]; o[7] = l[i].rel_pos[2];
    }
}
// m_get_similar_kfs over a stubbed SRBA state: tree[from] = {(to, distance)}, poses[root] = {(id, x y z yaw pitch roll)}
void ref_clear_graph(void* h) { refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h; e->rba.state.spanning_tree.sym.next_edge.clear(); e->rba.trees.clear(); }
void ref_graph_params(void* h, int submap, int tree_depth) { refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h; e->rba.parameters.ecp.submap_size = submap; e->rba.parameters.srba.max_tree_depth = tree_depth; }
void ref_add_tree_entry(void* h, long long from, long long to, long long dist)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    auto& mp = e->rba.state.spanning_tree.sym.next_edge[(TKeyFrameID)from];
    if (to >= 0) { TSpanTreeEntry t; t.next = 0; t.distance = (topo_dist_t)dist; mp[(TKeyFrameID)to] = t; }
}
void ref_add_tree_pose(void* h, long long root, long long id, const double* rv6 /* w, t */)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    mySRBA::pose_flag_t p; p.pose = CPose3D(CPose3DRotVec(rv6[0], rv6[1], rv6[2], rv6[3], rv6[4], rv6[5]));
    e->rba.trees[(TKeyFrameID)root][(TKeyFrameID)id] = p;
}
int ref_similar_kfs(void* h, long long new_kf_id, int nq, const int* q_ids, const double* q_scores, const double* incr_pose6, long long* sim_out, double* poses_out /* [n][6] w,t */,
                    long long* lc_id, int* n_poses)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    QueryResults q; for (int i = 0; i < nq; ++i) q.push_back(DBoW2::Result((DBoW2::EntryId)q_ids[i], q_scores[i]));
    e->m_incr_pose_from_last_kf = CPose3DRotVec(incr_pose6[0], incr_pose6[1], incr_pose6[2], incr_pose6[3], incr_pose6[4], incr_pose6[5]);
    TLoopClosureInfo out;
    const bool lc = e->m_get_similar_kfs((TKeyFrameID)new_kf_id, q, out);
    for (size_t i = 0; i < out.similar_kfs.size(); ++i) sim_out[i] = (long long)out.similar_kfs[i];
    *n_poses = (int)out.similar_kfs_poses.size();
    for (size_t i = 0; i < out.similar_kfs_poses.size(); ++i) { const CPose3DRotVec r(out.similar_kfs_poses[i]); for (int k = 0; k < 3; ++k) { poses_out[6 * i + k] = r.m_rotvec[k]; poses_out[6 * i + 3 + k] = r.m_coords[k]; } }
    *lc_id = out.lc_id == (TKeyFrameID)INVALID_KF_ID ? -1 : (long long)out.lc_id;
    return (int)out.similar_kfs.size() | (lc ? 0x10000 : 0);
}
// ---- the reference's own front-end loop on the GPU: initialize, then one call per stereo frame ----
void* ref_vo_open(const char* ini_file)
{
    refda::CSRBAStereoSLAMEstimator* e = new refda::CSRBAStereoSLAMEstimator();
    mrpt::utils::CConfigFile cfg(ini_file);
    general_options.loadFromConfigFile(cfg);
    try { e->initialize(cfg); } catch (std::exception& ex) { fprintf(stderr, "ref_vo_open: %s\n", ex.what()); delete e; return nullptr; }
    return e;
}
// out_i: [valid, n_left, n_right, n_matches, tracked_last_frame, tracked_last_kf, fast_th, orb_th, n_query]; out_pose: m_current_pose (x y z wx wy wz);
// ids[n_matches]; q: up to 4 (id, score) of queryDB on the candidate keyframe (frames >= 1; the candidate is dropped again, the first
// frame and every `insert_every`-th frame are inserted into the database)
int ref_vo_frame(void* h, const uint8_t* left, const uint8_t* right, int W, int H, int insert, int* out_i, double* out_pose, long long* ids, int* q_ids, double* q_scores)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    mrpt::obs::CObservationStereoImagesPtr stImgs(new mrpt::obs::CObservationStereoImages());
    stImgs->imageLeft.loadFromMemoryBuffer(W, H, false, const_cast<uint8_t*>(left)); stImgs->imageRight.loadFromMemoryBuffer(W, H, false, const_cast<uint8_t*>(right));
    size_t count = 0; bool ok = true; int nq = 0;
    try {
        if (e->m_keyframes.empty()) e->ref_first_frame(stImgs);
        else {
            ok = e->ref_next_frame(stImgs, count);
            if (ok) {
                QueryResults q; e->ref_candidate(q);
                nq = (int)q.size(); for (int i = 0; i < nq; ++i) { q_ids[i] = (int)q[i].Id; q_scores[i] = q[i].Score; }
                if (insert) { e->m_bow_manager.insertIntoDB(e->m_keyframes.back()); e->m_last_kf_ID++; } else e->m_keyframes.pop_back();
            }
        }
    } catch (std::exception& ex) { fprintf(stderr, "ref_vo_frame: %s\n", ex.what()); return -1; }
    CStereoSLAMKF tmp; tmp.getDataFromVOEngine(e->m_voEngine);
    const auto& r = e->m_odom_result;
    out_i[0] = ok && (e->m_keyframes.size() == 1 || r.valid); out_i[1] = (int)tmp.m_keypoints_left.size(); out_i[2] = (int)tmp.m_keypoints_right.size(); out_i[3] = (int)tmp.m_matches.size();
    out_i[4] = (int)r.tracked_feats_from_last_frame; out_i[5] = (int)r.tracked_feats_from_last_KF; out_i[6] = e->m_voEngine.getFASTThreshold(); out_i[7] = e->m_voEngine.getORBThreshold(); out_i[8] = nq;
    for (int i = 0; i < 6; ++i) out_pose[i] = e->m_current_pose[i];
    for (size_t i = 0; i < tmp.m_matches_ID.size(); ++i) ids[i] = (long long)tmp.m_matches_ID[i];
    return (int)tmp.m_matches.size();
}
void ref_vo_close(void* h) { delete (refda::CSRBAStereoSLAMEstimator*)h; }
double ref_update_translation_threshold(double x, double th) { return ref_updateTranslationThreshold(x, th); }
double ref_update_rotation_threshold(double x, double th) { return ref_updateRotationThreshold(x, th); }
void ref_project(double ul, double vl, double ur, void* h, double* xyz)
{
    refda::CSRBAStereoSLAMEstimator* e = (refda::CSRBAStereoSLAMEstimator*)h;
    const mrpt::math::TPoint3D p = ref_projectMatchTo3D(ul, vl, ur, e->srba_options.stereo_camera);
    xyz[0] = p.x; xyz[1] = p.y; xyz[2] = p.z;
}
}
'''


def extract(name):
    f, a, b, must = RANGES[name]
    lines = open(os.path.join(REF, f), encoding="latin-1").read().split("\n")
    if must not in lines[a - 1]:
        raise SystemExit("reference line %s:%d does not contain %r -- line ranges need updating" % (f, a, must))
    return "\n".join(lines[a - 1:b]) + "\n"


def generate():
    funcs = "\n".join("// ---- /root/reference/src/%s:%d-%d ----\n%s" % (RANGES[n][0], RANGES[n][1], RANGES[n][2], extract(n))
                      for n in ("data_association", "internal_da", "similar_kfs", "direction", "orb_distance", "fund", "pose"))
    # the free functions collide with the product's own (same names by design): the reference copies get a ref_ prefix
    project = extract("project").replace("projectMatchTo3D(", "ref_projectMatchTo3D(")
    thresholds = extract("thresholds").replace("updateTranslationThreshold(", "ref_updateTranslationThreshold(").replace("updateRotationThreshold(", "ref_updateRotationThreshold(")
    src = PRELUDE.replace("@@PROJECT@@", project).replace("@@THRESHOLDS@@", thresholds).replace("@@FUNCTIONS@@", funcs)
    src = src.replace("@@IDS_BLOCK@@", extract("ids_block")).replace("@@OBS_BLOCK@@", extract("obs_block"))
    for k in ("init_a", "init_b", "first_a", "first_b", "next", "candidate"):
        src = src.replace("@@%s@@" % k.upper(), extract(k))
    return src


def build(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print("oracle/ref_build.py: /root/reference is not present; using the prebuilt oracle/_ref/libref_da.so if it travelled with the snapshot")
        return os.path.exists(os.path.join(OUT, "libref_da.so"))
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "ref_da_generated.cpp")
    open(cpp, "w", encoding="latin-1").write(generate())
    lib = os.path.join(ROOT, "srba-stereo-slam_b200")
    cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I", HOST, "-include", "cstring", cpp, "-o", os.path.join(OUT, "libref_da.so"),
           "-L", HOST, "-lsrba_b200_host", "-L", lib, "-lsrba_b200", "-Wl,-rpath,$ORIGIN/../../srba-stereo-slam_b200/host", "-Wl,-rpath,$ORIGIN/../../srba-stereo-slam_b200"]
    subprocess.check_call(cmd)
    if not os.environ.get("REF_KEEP"): os.remove(cpp)                      # the generated unit holds reference source: it does not stay on disk, only the .so does
    if verbose:
        print("built", os.path.join(OUT, "libref_da.so"))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
