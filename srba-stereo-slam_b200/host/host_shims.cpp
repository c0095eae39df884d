// host_shims.cpp -- C entry points over the host-side data-association code (da_filters.h) for the ctypes test-suite.
// Same signatures as the shims the test-suite's reference build puts over the reference's own functions, so one Python driver feeds both
// with the same keyframes and diffs the results (tests/test_ref_da.py).  No GPU work here: the brute-force result is an input.
#include <cstring>
#include <map>
#include "da_filters.h"

namespace {
struct HostState {
    TSRBAStereoSLAMOptions srba_options; t_vector_kf m_keyframes; rso::CStereoOdometryEstimator m_voEngine;
    std::map<TKeyFrameID, std::map<TKeyFrameID, size_t>> tree; std::map<TKeyFrameID, std::map<TKeyFrameID, mrpt::poses::CPose3D>> poses;
    size_t submap = 15;
};
}
extern "C" {
struct srba_kf_view { int n_left, n_right, n_matches; const cv::KeyPoint* kl; const cv::KeyPoint* kr; const uint8_t* dl; const uint8_t* dr; const cv::DMatch* m; const long long* ids; long long kf_id; };

void* srba_host_create(const char* ini_file)
{
    HostState* e = new HostState();
    mrpt::utils::CConfigFile cfg(ini_file);
    e->srba_options.loadFromConfigFile(cfg);
    std::vector<std::string> s = { "RECTIFY", "DETECT", "MATCH", "IF-MATCH", "LEAST_SQUARES", "GUI", "GENERAL" };
    e->m_voEngine.loadParamsFromConfigFile(cfg, s);
    const mrpt::poses::CPose3D img_to_camera_pose(0, 0, 0, DEG2RAD(-90), 0, DEG2RAD(-90));
    e->srba_options.camera_pose_on_robot_rvt = mrpt::poses::CPose3DRotVec(img_to_camera_pose);
    e->srba_options.camera_pose_on_robot_rvt_inverse = e->srba_options.camera_pose_on_robot_rvt.getInverse();
    return e;
}
void srba_host_destroy(void* h) { delete (HostState*)h; }
void srba_host_set_keyframes(void* h, const srba_kf_view* kfs, int n)
{
    HostState* e = (HostState*)h;
    e->m_keyframes.assign(n, CStereoSLAMKF());
    for (int i = 0; i < n; ++i) {
        CStereoSLAMKF& k = e->m_keyframes[i]; const srba_kf_view& r = kfs[i];
        k.m_keypoints_left.assign(r.kl, r.kl + r.n_left); k.m_keypoints_right.assign(r.kr, r.kr + r.n_right);
        k.m_descriptors_left.create(r.n_left, 32, CV_8U); k.m_descriptors_right.create(r.n_right, 32, CV_8U);
        if (r.n_left) std::memcpy(k.m_descriptors_left.ptr(0), r.dl, (size_t)r.n_left * 32);
        if (r.n_right) std::memcpy(k.m_descriptors_right.ptr(0), r.dr, (size_t)r.n_right * 32);
        k.m_matches.assign(r.m, r.m + r.n_matches); k.m_matches_ID.clear();
        for (int j = 0; j < r.n_matches; ++j) k.m_matches_ID.push_back((size_t)r.ids[j]);
        k.setKFID((size_t)r.kf_id);
    }
}
int srba_host_internal_da(void* h, int this_idx, int other_idx, const cv::DMatch* bf, int nbf, const double* ini_pose6, long long* first, double* dist, int* status)
{
    HostState* e = (HostState*)h;
    const CStereoSLAMKF& kf = e->m_keyframes[this_idx];
    std::vector<cv::DMatch> m(bf, bf + nbf); t_kf_da_info da; std::vector<int> st;
    const mrpt::poses::CPose3DRotVec ini(ini_pose6[0], ini_pose6[1], ini_pose6[2], ini_pose6[3], ini_pose6[4], ini_pose6[5]);
    internalDataAssociationFromMatches(e->m_voEngine, e->srba_options, kf, e->m_keyframes[other_idx], m, da, ini, &st);
    for (size_t k = 0; k < kf.m_matches.size(); ++k) {
        first[k] = da.tracking_info[k].first == INVALID_IDX ? -1 : (long long)da.tracking_info[k].first; dist[k] = da.tracking_info[k].second;
        if (status) status[k] = st[k];
    }
    return (int)da.tracked_matches;
}
void srba_host_filter(void* h, int which, int this_idx, int other_idx, const cv::DMatch* bf, int* status, const double* ini_pose6)
{
    HostState* e = (HostState*)h;
    const CStereoSLAMKF& a = e->m_keyframes[this_idx]; const CStereoSLAMKF& b = e->m_keyframes[other_idx];
    const size_t n = a.m_matches.size();
    TVectorDAMatchInfo m(n);
    for (size_t k = 0; k < n; ++k) { m.m_matches[k] = bf[k]; m.m_status[k] = (TVectorDAMatchInfo::TDAMatchStatus)status[k]; }
    size_t tracked = 0; for (size_t k = 0; k < n; ++k) tracked += m.m_status[k] == TVectorDAMatchInfo::sTRACKED;
    const TSRBAStereoSLAMOptions& o = e->srba_options;
    if (which == 0) detectOutliersWithDirection(m, o.stereo_camera.leftCamera.nrows, a, b);
    else if (which == 1) detectOutliersWithOrbDistance(m, o.max_orb_distance_da, a, b);
    else if (which == 2) detectOutliersWithF(m, tracked, a, b, o.max_y_diff_epipolar, o.ransac_fit_prob);
    else detectOutliersWithChangeInPose(e->m_voEngine, m, a, b, mrpt::poses::CPose3DRotVec(ini_pose6[0], ini_pose6[1], ini_pose6[2], ini_pose6[3], ini_pose6[4], ini_pose6[5]),
                                        o.stereo_camera, o.residual_th);
    for (size_t k = 0; k < n; ++k) status[k] = (int)m.m_status[k];
}
void srba_host_assign_ids(void* h, int nkf, const long long* first, const int* da_kf_idx, const int* sorted, long long* last_match_id, long long* out_ids, long long* counts2, int* per_kf)
{
    HostState* e = (HostState*)h;
    CStereoSLAMKF& new_kf = e->m_keyframes.back();
    const size_t n = new_kf.m_matches.size();
    TVectorKfsDaInfo da(nkf);
    for (int k = 0; k < nkf; ++k) {
        da[k].kf_idx = da_kf_idx[k]; da[k].tracking_info.resize(n);
        for (size_t m = 0; m < n; ++m) { const long long f = first[(size_t)k * n + m]; da[k].tracking_info[m] = std::make_pair(f < 0 ? INVALID_IDX : (size_t)f, 0.0); }
    }
    std::vector<size_t> srt(sorted, sorted + nkf); std::vector<int> pk; size_t last = (size_t)*last_match_id;
    const std::pair<size_t, size_t> r = assignMatchIds(new_kf, da, srt, e->m_keyframes, last, &pk, nullptr);
    for (size_t m = 0; m < n; ++m) out_ids[m] = (long long)new_kf.m_matches_ID[m];
    *last_match_id = (long long)last; counts2[0] = (long long)r.first; counts2[1] = (long long)r.second;
    for (int k = 0; k < nkf; ++k) per_kf[k] = pk[k];
}
void srba_host_observations(void* h, int idx, double* out)
{
    HostState* e = (HostState*)h;
    std::vector<TObservation> l; assembleObservations(e->m_keyframes[idx], e->srba_options, l);
    for (size_t i = 0; i < l.size(); ++i) { double* o = out + 8 * i; o[0] = (double)l[i].feat_id; o[1] = l[i].lx; o[2] = l[i].ly; o[3] = l[i].rx; o[4] = l[i].ry; o[5] = l[i].X; o[6] = l[i].Y; o[7] = l[i].Z; }
}
void srba_host_clear_graph(void* h) { HostState* e = (HostState*)h; e->tree.clear(); e->poses.clear(); }
void srba_host_graph_params(void* h, int submap, int) { ((HostState*)h)->submap = submap; }
void srba_host_add_tree_entry(void* h, long long from, long long to, long long dist) { HostState* e = (HostState*)h; auto& m = e->tree[(TKeyFrameID)from]; if (to >= 0) m[(TKeyFrameID)to] = (size_t)dist; }
void srba_host_add_tree_pose(void* h, long long root, long long id, const double* rv6)
{
    ((HostState*)h)->poses[(TKeyFrameID)root][(TKeyFrameID)id] = mrpt::poses::CPose3D(mrpt::poses::CPose3DRotVec(rv6[0], rv6[1], rv6[2], rv6[3], rv6[4], rv6[5]));
}
int srba_host_similar_kfs(void* h, long long new_kf_id, int nq, const int* q_ids, const double* q_scores, const double* incr_pose6, long long* sim_out, double* poses_out,
                          long long* lc_id, int* n_poses)
{
    HostState* e = (HostState*)h;
    QueryResults q; for (int i = 0; i < nq; ++i) q.push_back(DBoW2::Result((DBoW2::EntryId)q_ids[i], q_scores[i]));
    const mrpt::poses::CPose3DRotVec incr(incr_pose6[0], incr_pose6[1], incr_pose6[2], incr_pose6[3], incr_pose6[4], incr_pose6[5]);
    TLoopClosureInfo out;
    const bool lc = getSimilarKfs((TKeyFrameID)new_kf_id, q, e->srba_options, e->submap, incr,
        [e](TKeyFrameID a, TKeyFrameID b) -> size_t {
            auto f = e->tree.find(a); if (f == e->tree.end()) throw std::logic_error("[ERROR :: Check Loop Closure] 'itFrom' is not into the spanning_tree!");
            auto t = f->second.find(b); return t == f->second.end() ? std::numeric_limits<size_t>::max() : t->second; },
        [e](TKeyFrameID r, TKeyFrameID i, mrpt::poses::CPose3D& o) -> bool {
            auto f = e->poses.find(r); if (f == e->poses.end()) return false; auto t = f->second.find(i); if (t == f->second.end()) return false; o = t->second; return true; },
        out);
    for (size_t i = 0; i < out.similar_kfs.size(); ++i) sim_out[i] = (long long)out.similar_kfs[i];
    *n_poses = (int)out.similar_kfs_poses.size();
    for (size_t i = 0; i < out.similar_kfs_poses.size(); ++i) { const mrpt::poses::CPose3DRotVec r(out.similar_kfs_poses[i]); for (int k = 0; k < 3; ++k) { poses_out[6 * i + k] = r.m_rotvec[k]; poses_out[6 * i + 3 + k] = r.m_coords[k]; } }
    *lc_id = out.lc_id == (TKeyFrameID)INVALID_KF_ID ? -1 : (long long)out.lc_id;
    return (int)out.similar_kfs.size() | (lc ? 0x10000 : 0);
}
double srba_host_update_translation_threshold(double x, double th) { return updateTranslationThreshold(x, th); }
double srba_host_update_rotation_threshold(double x, double th) { return updateRotationThreshold(x, th); }
}
