#include "da_filters.h"
#include "CBoWManager.h"
#include <cmath>
#include <algorithm>
#include <deque>
#include <stdexcept>

// ---- cores on plain arrays (shared by the class-level functions and the C test shims) ----------------------
static void direction_core(const float* txy, const float* oxy, int n, float offset, int* status)
{
    const int N_BINS = 36, BIN_WIDTH = 10;
    std::vector<size_t> hist(N_BINS, 0);
    std::vector<std::deque<size_t>> hist_m(N_BINS);
    for (int k = 0; k < n; ++k) {
        if (status[k] != TVectorDAMatchInfo::sTRACKED) continue;
        const float tx = txy[2 * k], ty = txy[2 * k + 1];
        const float ox = oxy[2 * k], oy = oxy[2 * k + 1] + offset;     // 'offset' reduces the influence of location errors
        double ang = 180 / M_PI * atanf((oy - ty) / (ox - tx)) + 180.;  // as written in the reference: atanf, not atan2
        if (!(ang >= 0.0 && ang < 360.0)) continue;                    // NaN (coincident points): UB in the reference, skipped here
        const size_t bin = (size_t)floor(ang / BIN_WIDTH);
        hist[bin]++; hist_m[bin].push_back(k);
    }
    int mode = 0, best_so_far = 0;
    for (size_t k = 0; k < hist.size(); ++k) if ((int)hist[k] > best_so_far) { best_so_far = (int)hist[k]; mode = (int)k; }
    const size_t ibin_0 = mode == 0 ? N_BINS - 1 : mode - 1, ibin_1 = mode, ibin_2 = mode == N_BINS - 1 ? 0 : mode + 1;
    for (size_t k = 0; k < (size_t)N_BINS; ++k) {
        if (k == ibin_0 || k == ibin_1 || k == ibin_2) continue;
        for (size_t k1 = 0; k1 < hist_m[k].size(); ++k1) status[hist_m[k][k1]] = TVectorDAMatchInfo::sREJ_SLOPE;
    }
}

static void orb_distance_core(const int* query_idx, const int* train_idx, const float* distance, int n, int n_other, double max_dist, int* status)
{
    std::vector<std::pair<int, float>> other_matched(n_other, std::make_pair(-1, std::numeric_limits<float>::max()));   // t_vector_pair_idx_distance, UTL.h:509
    size_t stage1_counter = 0;
    for (int i = 0; i < n; ++i) {
        const int q = query_idx[i], t = train_idx[i];
        if (status[q] != TVectorDAMatchInfo::sTRACKED) continue;                       // rejected by a previous stage
        if (distance[i] > max_dist) { status[q] = TVectorDAMatchInfo::sREJ_ORB; continue; }
        if (distance[i] > other_matched[t].second) { status[q] = TVectorDAMatchInfo::sNON_TRACKED; continue; }
        if (other_matched[t].first != -1) { status[other_matched[t].first] = TVectorDAMatchInfo::sNON_TRACKED; --stage1_counter; }
        other_matched[t].first = q;
        other_matched[t].second = (float)(++stage1_counter);    // the reference's bug (EST.cpp:2003-2007): the counter, not the distance
    }
}

// ---- reference-shaped entry points ------------------------------------------------------------------------------
void detectOutliersWithDirection(TVectorDAMatchInfo& m, const size_t& offset, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf)
{
    const int n = (int)m.size();
    std::vector<float> txy(2 * n), oxy(2 * n); std::vector<int> st(n);
    for (int k = 0; k < n; ++k) {
        st[k] = m.m_status[k];
        if (m.m_status[k] != TVectorDAMatchInfo::sTRACKED) continue;
        const cv::KeyPoint& a = this_kf.m_keypoints_left[this_kf.m_matches[k].queryIdx];
        const cv::KeyPoint& b = other_kf.m_keypoints_left[other_kf.m_matches[m.m_matches[k].trainIdx].queryIdx];
        txy[2 * k] = a.pt.x; txy[2 * k + 1] = a.pt.y; oxy[2 * k] = b.pt.x; oxy[2 * k + 1] = b.pt.y;
    }
    direction_core(txy.data(), oxy.data(), n, (float)offset, st.data());
    for (int k = 0; k < n; ++k) m.m_status[k] = (TVectorDAMatchInfo::TDAMatchStatus)st[k];
}

void detectOutliersWithOrbDistance(TVectorDAMatchInfo& m, double max_orb_distance_da, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf)
{
    const int n = (int)m.m_matches.size();
    std::vector<int> q(n), t(n), st(m.size()); std::vector<float> d(n);
    for (int i = 0; i < n; ++i) { q[i] = m.m_matches[i].queryIdx; t[i] = m.m_matches[i].trainIdx; d[i] = m.m_matches[i].distance; }
    for (size_t k = 0; k < m.size(); ++k) st[k] = m.m_status[k];
    (void)this_kf;
    orb_distance_core(q.data(), t.data(), d.data(), n, (int)other_kf.m_matches.size(), max_orb_distance_da, st.data());
    for (size_t k = 0; k < m.size(); ++k) m.m_status[k] = (TVectorDAMatchInfo::TDAMatchStatus)st[k];
}

void fillTrackingInfo(const TVectorDAMatchInfo& m, const CStereoSLAMKF& other_kf, t_kf_da_info& out_da)
{
    out_da.kf_idx = other_kf.m_kf_ID; out_da.tracked_matches = 0;
    out_da.tracking_info.assign(m.size(), std::make_pair(INVALID_IDX, 0.0));
    for (size_t k = 0; k < m.size(); ++k)
        if (m.m_status[k] == TVectorDAMatchInfo::sTRACKED) {
            out_da.tracking_info[k] = std::make_pair((size_t)m.m_matches[k].trainIdx, (double)m.m_matches[k].distance);
            out_da.tracked_matches++;
        }
}

void internalDataAssociation(sfe_context* ctx, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, bool by_direction, bool by_orb,
                             double max_orb_distance_da, size_t image_rows, t_kf_da_info& out_da)
{
    TVectorDAMatchInfo this_matches(this_kf.m_matches.size());
    matchKeyframes(ctx, this_kf, other_kf, this_matches.m_matches);                     // EST.cpp:1432-1452 on the GPU
    if (this_matches.m_matches.size() != this_matches.m_status.size()) {               // empty train set: nothing tracked
        out_da.kf_idx = other_kf.m_kf_ID; out_da.tracked_matches = 0;
        out_da.tracking_info.assign(this_kf.m_matches.size(), std::make_pair(INVALID_IDX, 0.0));
        return;
    }
    if (by_direction) detectOutliersWithDirection(this_matches, image_rows, this_kf, other_kf);
    if (by_orb) detectOutliersWithOrbDistance(this_matches, max_orb_distance_da, this_kf, other_kf);
    fillTrackingInfo(this_matches, other_kf, out_da);
}

// ---- geometric filters -----------------------------------------------------------------------------------------------
void detectOutliersWithF(TVectorDAMatchInfo& m, const size_t& num_tracked, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf,
                         double max_y_diff_epipolar, double ransac_fit_prob)
{
    std::vector<float> p_this(2 * num_tracked), p_other(2 * num_tracked);
    size_t k0 = 0;
    for (size_t k = 0; k < m.size() && k0 < num_tracked; ++k) {
        if (m.m_status[k] != TVectorDAMatchInfo::sTRACKED) continue;
        const cv::KeyPoint& a = this_kf.m_keypoints_left[this_kf.m_matches[k].queryIdx];
        const cv::KeyPoint& b = other_kf.m_keypoints_left[other_kf.m_matches[m.m_matches[k].trainIdx].queryIdx];
        p_this[2 * k0] = a.pt.x; p_this[2 * k0 + 1] = a.pt.y; p_other[2 * k0] = b.pt.x; p_other[2 * k0 + 1] = b.pt.y; ++k0;
    }
    std::vector<uint8_t> inl(num_tracked, 0);
    findFundamentalMatRansac(p_other.data(), p_this.data(), (int)k0, max_y_diff_epipolar, ransac_fit_prob, inl.data(), nullptr, nullptr);   // EST.cpp:2043
    k0 = 0;
    for (size_t k = 0; k < m.size(); ++k)
        if (m.m_status[k] == TVectorDAMatchInfo::sTRACKED && k0 < num_tracked && inl[k0++] == 0) m.m_status[k] = TVectorDAMatchInfo::sREJ_FUND_MATRIX;
}

void detectOutliersWithChangeInPose(rso::CStereoOdometryEstimator& voEngine, TVectorDAMatchInfo& m, const CStereoSLAMKF& this_kf,
                                    const CStereoSLAMKF& other_kf, const mrpt::poses::CPose3DRotVec& kf_ini_rel_pose,
                                    const mrpt::utils::TStereoCamera& stereo_camera, double residual_th)
{
    rso::vector_index_pairs_t tracked_pairs; tracked_pairs.reserve(this_kf.m_matches.size());
    for (size_t k = 0; k < m.size(); ++k)
        if (m.m_status[k] == TVectorDAMatchInfo::sTRACKED) tracked_pairs.push_back(std::make_pair((size_t)m.m_matches[k].trainIdx /*other*/, k /*this*/));
    std::vector<double> initialPoseVector(6);                           // [w1,w2,w3,tx,ty,tz], EST.cpp:2131-2135
    for (int i = 0; i < 3; ++i) { initialPoseVector[i] = kf_ini_rel_pose.m_rotvec[i]; initialPoseVector[3 + i] = kf_ini_rel_pose.m_coords[i]; }
    rso::CStereoOdometryEstimator::TStereoOdometryResult odometry_result;
    voEngine.params_least_squares.use_custom_initial_pose = true;
    voEngine.getChangeInPose(tracked_pairs, other_kf.m_matches, this_kf.m_matches, other_kf.m_keypoints_left, other_kf.m_keypoints_right,
                             this_kf.m_keypoints_left, this_kf.m_keypoints_right, stereo_camera, odometry_result, initialPoseVector);
    voEngine.params_least_squares.use_custom_initial_pose = false;
    VERBOSE_LEVEL(2) << "Iterations: " << odometry_result.num_it << " (initial) and " << odometry_result.num_it_final << " (final) " << std::endl
                     << "Change in pose: " << odometry_result.outPose << std::endl;
    size_t k0 = 0;
    for (size_t k = 0; k < m.size(); ++k) {
        if (m.m_status[k] != TVectorDAMatchInfo::sTRACKED) continue;
        if (!odometry_result.valid || odometry_result.out_residual[k0] > residual_th) m.m_status[k] = TVectorDAMatchInfo::sREJ_CHANGE_POSE;
        ++k0;
    }
}

static size_t count_tracked(const TVectorDAMatchInfo& m)
{
    size_t n = 0; for (size_t k = 0; k < m.size(); ++k) n += m.m_status[k] == TVectorDAMatchInfo::sTRACKED; return n;
}

void internalDataAssociation(rso::CStereoOdometryEstimator& voEngine, const TSRBAStereoSLAMOptions& o, const CStereoSLAMKF& this_kf,
                             const CStereoSLAMKF& other_kf, t_kf_da_info& out_da, const mrpt::poses::CPose3DRotVec& kf_ini_rel_pose)
{
    std::vector<cv::DMatch> bf;
    matchKeyframes(voEngine.context(), this_kf, other_kf, bf);                          // EST.cpp:1432-1452 on the GPU
    internalDataAssociationFromMatches(voEngine, o, this_kf, other_kf, bf, out_da, kf_ini_rel_pose, nullptr);
}

void internalDataAssociationFromMatches(rso::CStereoOdometryEstimator& voEngine, const TSRBAStereoSLAMOptions& o, const CStereoSLAMKF& this_kf,
                                        const CStereoSLAMKF& other_kf, const std::vector<cv::DMatch>& bf_matches, t_kf_da_info& out_da,
                                        const mrpt::poses::CPose3DRotVec& kf_ini_rel_pose, std::vector<int>* status_out)
{
    const size_t this_num_matches = this_kf.m_matches.size();
    out_da.kf_idx = other_kf.m_kf_ID; out_da.tracked_matches = 0;
    out_da.tracking_info.assign(this_num_matches, std::make_pair(INVALID_IDX, 0.0));
    TVectorDAMatchInfo this_matches(this_num_matches);
    this_matches.m_matches = bf_matches;
    if (status_out) status_out->assign(this_num_matches, (int)TVectorDAMatchInfo::sTRACKED);
    if (this_matches.m_matches.size() != this_num_matches) return;                      // empty train set: BFMatcher returns no matches
    const std::string sep = general_options.out_dir + SRBA_PATH_SEP;
    if (general_options.debug) {                                                        // if_raw_match_kfXXXX_with_kfYYYY.txt, EST.cpp:1454-1473
        srba_prepare_out_dir();
        if (FILE* f = fopen((sep + mrpt::format("if_raw_match_kf%04d_with_kf%04d.txt", (int)this_kf.m_kf_ID, (int)other_kf.m_kf_ID)).c_str(), "wt")) {
            fprintf(f, "%% OTHER_LX OTHER_LY THIS_LX THIS_LY DISTANCE\n");
            for (const cv::DMatch& d : this_matches.m_matches) {
                const cv::KeyPoint& t = this_kf.m_keypoints_left[this_kf.m_matches[d.queryIdx].queryIdx];
                const cv::KeyPoint& ot = other_kf.m_keypoints_left[other_kf.m_matches[d.trainIdx].queryIdx];
                fprintf(f, "%.2f %.2f %.2f %.2f %.2f\n", ot.pt.x, ot.pt.y, t.pt.x, t.pt.y, d.distance);
            }
            fclose(f);
        }
    }
    bool invalid = false;
    size_t inliers = this_num_matches;
    if (o.da_filter_by_direction) { detectOutliersWithDirection(this_matches, o.stereo_camera.leftCamera.nrows, this_kf, other_kf); inliers = count_tracked(this_matches); }
    if (o.da_filter_by_orb_distance) { detectOutliersWithOrbDistance(this_matches, o.max_orb_distance_da, this_kf, other_kf); inliers = count_tracked(this_matches); }
    if (o.da_filter_by_fund_matrix) {
        if (inliers < 15) invalid = true;                                               // cv::findFundamentalMat(RANSAC) needs 15 (EST.cpp:1591)
        else { detectOutliersWithF(this_matches, inliers, this_kf, other_kf, o.max_y_diff_epipolar, o.ransac_fit_prob); inliers = count_tracked(this_matches); }
    }
    if (o.da_filter_by_pose_change) {
        if (inliers < 15) invalid = true;                                               // EST.cpp:1619 (the count is the F stage's even when that stage was skipped)
        else { detectOutliersWithChangeInPose(voEngine, this_matches, this_kf, other_kf, kf_ini_rel_pose, o.stereo_camera, o.residual_th); inliers = count_tracked(this_matches); }
    }
    VERBOSE_LEVEL(2) << "[iDA " << this_kf.m_kf_ID << "->" << other_kf.m_kf_ID << "]: Final number of inliers: " << inliers << "/" << this_num_matches << std::endl;
    if (status_out) for (size_t k = 0; k < this_num_matches; ++k) (*status_out)[k] = (int)this_matches.m_status[k];
    if (invalid) return;
    FILE* f2 = nullptr;
    if (general_options.debug) {                                                        // if_match_after_kfXXXX_with_kfYYYY.txt, EST.cpp:1650-1689
        f2 = fopen((sep + mrpt::format("if_match_after_kf%04d_with_kf%04d.txt", (int)this_kf.m_kf_ID, (int)other_kf.m_kf_ID)).c_str(), "wt");
        if (f2) fprintf(f2, "%%STATUS THIS_LU THIS_LV OTHER_LU OTHER_LV DISTANCE\n");
    }
    for (size_t k = 0; k < this_matches.size(); ++k) {
        const cv::DMatch& d = this_matches.m_matches[k];
        if (this_matches.m_status[k] == TVectorDAMatchInfo::sTRACKED) { out_da.tracking_info[k] = std::make_pair((size_t)d.trainIdx, (double)d.distance); out_da.tracked_matches++; }
        if (f2) {
            const cv::KeyPoint& t = this_kf.m_keypoints_left[this_kf.m_matches[d.queryIdx].queryIdx];
            const cv::KeyPoint& ot = other_kf.m_keypoints_left[other_kf.m_matches[d.trainIdx].queryIdx];
            fprintf(f2, "%d %.2f %.2f %.2f %.2f %.2f\n", (int)this_matches.m_status[k], (double)t.pt.x, (double)t.pt.y, (double)ot.pt.x, (double)ot.pt.y, d.distance);
        }
    }
    if (f2) fclose(f2);
}

void dataAssociation(rso::CStereoOdometryEstimator& voEngine, const TSRBAStereoSLAMOptions& o, const t_vector_kf& keyframes, const CStereoSLAMKF& kf,
                     const TLoopClosureInfo& lc_info, TVectorKfsDaInfo& out_da)
{
    out_da.clear(); out_da.reserve(lc_info.similar_kfs.size());
    const bool with_pose = o.da_stage2_method == TSRBAStereoSLAMOptions::ST2M_CHANGEPOSE || o.da_stage2_method == TSRBAStereoSLAMOptions::ST2M_BOTH;
    for (size_t k = 0; k < lc_info.similar_kfs.size(); ++k) {
        const size_t idx = lc_info.similar_kfs[k];
        VERBOSE_LEVEL(1) << "	DA with KF #" << keyframes[idx].m_kf_ID << std::endl;
        mrpt::poses::CPose3DRotVec initialPose;
        if (with_pose && k < lc_info.similar_kfs_poses.size()) initialPose = mrpt::poses::CPose3DRotVec(lc_info.similar_kfs_poses[k]);
        out_da.push_back(t_kf_da_info());
        internalDataAssociation(voEngine, o, kf, keyframes[idx], out_da.back(), initialPose);
    }
    if (lc_info.similar_kfs.empty() && keyframes.size() >= 2) {          // "this shouldn't happen, but just in case" (EST.cpp:1399-1410)
        out_da.push_back(t_kf_da_info());
        internalDataAssociation(voEngine, o, kf, *(keyframes.rbegin() + 1), out_da.back());
    }
}

bool getSimilarKfs(const TKeyFrameID& newKfId, const QueryResults& q, const TSRBAStereoSLAMOptions& o, size_t submap_size,
                   const mrpt::poses::CPose3DRotVec& incr_pose_from_last_kf, const TTopoDistanceFn& topo_distance, const TTreePoseFn& tree_pose,
                   TLoopClosureInfo& out)
{
    if (q.empty()) throw std::logic_error("Parameter 'dbQueryResults' is empty. This method should not be called here.");
    VERBOSE_LEVEL(2) << "dbQueryResults: " << q << std::endl;
    if (q[0].Score < o.query_score_th) std::cerr << "[WARNING] Best result in 'dbQueryResults' is below a minimum threshold. Lost camera?" << std::endl;
    const bool with_pose = o.da_stage2_method == TSRBAStereoSLAMOptions::ST2M_CHANGEPOSE || o.da_stage2_method == TSRBAStereoSLAMOptions::ST2M_BOTH;
    const mrpt::poses::CPose3DRotVec to_prev = incr_pose_from_last_kf.getInverse();      // previous KF wrt this one, camera to camera
    out.similar_kfs.clear(); out.similar_kfs_poses.clear();
    out.similar_kfs.push_back(newKfId - 1);                                              // the previous keyframe is always a candidate
    if (q.size() == 1) {
        if (with_pose) out.similar_kfs_poses.push_back(mrpt::poses::CPose3D(to_prev));
        return false;
    }
    out.lc_id = (TKeyFrameID)INVALID_KF_ID;
    bool potential_lc = false;
    const TKeyFrameID fromIdBase = submap_size * ((newKfId - 1) / submap_size);          // current local-map centre
    const double add_similar_kf_th = 0.8 * q[0].Score;
    for (size_t i = 0; i < q.size(); ++i) {
        const TKeyFrameID toId = q[i].Id;
        if (toId == newKfId - 1) continue;
        if (q[i].Score > add_similar_kf_th) { out.similar_kfs.push_back(toId); continue; }
        const size_t topo = fromIdBase == toId ? 0 : topo_distance(fromIdBase, toId);
        if (out.lc_id == (TKeyFrameID)INVALID_KF_ID && topo > o.lc_distance && q[i].Score > 0.05) {   // first far-away hit only
            out.lc_id = toId; potential_lc = true; out.similar_kfs.push_back(toId);
            VERBOSE_LEVEL(1) << "		FOUND POTENTIAL LOOP CLOSURE " << std::endl;
        }
        VERBOSE_LEVEL(2) << "		Distance from " << fromIdBase << " to " << toId << " : " << topo << std::endl;
    }
    out.similar_kfs_poses.assign(out.similar_kfs.size(), mrpt::poses::CPose3D());
    if (with_pose)
        for (size_t k = 0; k < out.similar_kfs.size(); ++k) {
            if (out.similar_kfs[k] == newKfId - 1) { out.similar_kfs_poses[k] = mrpt::poses::CPose3D(to_prev); continue; }
            mrpt::poses::CPose3D tp;
            if (tree_pose(newKfId - 1, out.similar_kfs[k], tp)) {                         // K^-1 * pose_other_wrt_prev * K, K = camera pose on the robot
                const mrpt::poses::CPose3DRotVec aux = o.camera_pose_on_robot_rvt_inverse + mrpt::poses::CPose3DRotVec(tp) + o.camera_pose_on_robot_rvt;
                out.similar_kfs_poses[k] = mrpt::poses::CPose3D(to_prev + aux);
            }
        }
    return potential_lc;
}

std::pair<size_t, size_t> assignMatchIds(CStereoSLAMKF& new_kf, const TVectorKfsDaInfo& daInfo, const std::vector<size_t>& sorted,
                                         const t_vector_kf& keyframes, size_t& last_match_id, std::vector<int>* per_kf_tracked, std::vector<double>* da_dist)
{
    const size_t nMatches = new_kf.m_matches.size(), numberSimilarKfs = sorted.size();
    new_kf.m_matches_ID.resize(nMatches);            // the candidate carries the engine's ids (getDataFromVOEngine); every entry is overwritten
    if (per_kf_tracked) per_kf_tracked->assign(daInfo.size(), 0);
    if (da_dist) da_dist->clear();
    size_t numberNewFeats = 0, numberTrackedFeats = 0;
    std::set<size_t> foundIds;
    for (size_t m = 0; m < nMatches; ++m) {
        bool tracked = false;
        for (size_t k = 0; !tracked && k < numberSimilarKfs; ++k) {
            const t_kf_da_info& da_info = daInfo[sorted[k]];
            if (da_info.tracking_info[m].first != INVALID_IDX) {
                const size_t other_match_idx = da_info.tracking_info[m].first;
                const size_t other_match_id = keyframes[da_info.kf_idx].m_matches_ID[other_match_idx];
                if (foundIds.find(other_match_id) != foundIds.end()) break;            // "keep the first one" (EST.cpp:587-592)
                foundIds.insert(other_match_id);
                new_kf.m_matches_ID[m] = other_match_id;
                if (per_kf_tracked) (*per_kf_tracked)[sorted[k]]++;
                if (da_dist) da_dist->push_back(da_info.tracking_info[m].second);
                numberTrackedFeats++; tracked = true;
            }
        }
        if (!tracked) { new_kf.m_matches_ID[m] = ++last_match_id; numberNewFeats++; if (da_dist) da_dist->push_back(-1.0); }
    }
    return std::make_pair(numberTrackedFeats, numberNewFeats);
}

void projectMatchTo3D(double ul, double vl, double ur, const TStereoCameraParams& c, double& X, double& Y, double& Z)
{
    const double b_d = c.baseline / (c.fl * (c.cur - ur) + c.fr * (ul - c.cul));
    X = b_d * c.fr * (ul - c.cul); Y = b_d * c.fr * (vl - c.cvl); Z = b_d * c.fl * c.fr;
}

mrpt::math::TPoint3D projectMatchTo3D(const double& ul, const double& vl, const double& ur, const mrpt::utils::TStereoCamera& c)
{
    const TStereoCameraParams p = { c.leftCamera.cx(), c.leftCamera.cy(), c.leftCamera.fx(), c.rightCamera.cx(), c.rightCamera.fx(), c.rightCameraPose[0] };
    mrpt::math::TPoint3D r; projectMatchTo3D(ul, vl, ur, p, r.x, r.y, r.z);
    return r;
}

void assembleObservations(const CStereoSLAMKF& kf, const TSRBAStereoSLAMOptions& o, std::vector<TObservation>& out)
{
    out.resize(kf.m_matches.size());
    for (size_t i = 0; i < kf.m_matches.size(); ++i) {
        const cv::KeyPoint& l = kf.m_keypoints_left[kf.m_matches[i].queryIdx]; const cv::KeyPoint& r = kf.m_keypoints_right[kf.m_matches[i].trainIdx];
        TObservation& ob = out[i];
        ob.feat_id = kf.m_matches_ID[i];
        ob.lx = l.pt.x; ob.ly = l.pt.y; ob.rx = r.pt.x; ob.ry = r.pt.y;
        const mrpt::math::TPoint3D p = o.camera_pose_on_robot_rvt + projectMatchTo3D(l.pt.x, l.pt.y, r.pt.x, o.stereo_camera);   // EST.cpp:158, 698
        ob.X = p.x; ob.Y = p.y; ob.Z = p.z;
    }
}

void assembleObservations(const CStereoSLAMKF& kf, const TStereoCameraParams& cam, std::vector<TObservation>& out)
{
    out.resize(kf.m_matches.size());
    for (size_t i = 0; i < kf.m_matches.size(); ++i) {
        const cv::KeyPoint& l = kf.m_keypoints_left[kf.m_matches[i].queryIdx]; const cv::KeyPoint& r = kf.m_keypoints_right[kf.m_matches[i].trainIdx];
        TObservation& o = out[i];
        o.feat_id = i < kf.m_matches_ID.size() ? kf.m_matches_ID[i] : i;
        o.lx = l.pt.x; o.ly = l.pt.y; o.rx = r.pt.x; o.ry = r.pt.y;
        projectMatchTo3D(l.pt.x, l.pt.y, r.pt.x, cam, o.X, o.Y, o.Z);
    }
}

int processWithAdaptiveThresholds(rso::CStereoOdometryEstimator& vo, rso::CStereoOdometryEstimator::TStereoOdometryRequest& req,
                                  rso::CStereoOdometryEstimator::TStereoOdometryResult& res, const TAdaptiveThOptions& opt)
{
    const size_t octave = 0;
    int calls = 0;
    req.repeat = false;
    do {
        vo.processNewImagePair(req, res); ++calls;
        if (res.stereo_matches.empty()) throw std::runtime_error("processNewImagePair failed: " + vo.lastError());
        if (opt.orb_adaptive_fast_th) {
            if (int(res.stereo_matches[octave]) < opt.adaptive_th_min_matches) {
                if (!vo.isFASTThMin()) { vo.setFASTThreshold(std::max(0, vo.getFASTThreshold() - 10)); req.repeat = true; }
                else if (!vo.isORBThMax()) { vo.setORBThreshold(vo.getORBThreshold() + 10); req.repeat = true; }
                else if (res.stereo_matches[octave] >= 8) { req.repeat = false; /* limits reached but usable: go on */ }
                else throw std::runtime_error("The number of found matches is less than the minimum. Aborting");   // EST.cpp:294
            } else if (res.stereo_matches[octave] < opt.adaptive_th_min_matches * 1.2) {
                if (!vo.isFASTThMin()) vo.setFASTThreshold(std::max(0, vo.getFASTThreshold() - 5));
                else if (!vo.isORBThMax()) vo.setORBThreshold(vo.getORBThreshold() + 5);
                req.repeat = false;
            } else {
                vo.setFASTThreshold(std::min(opt.detect_fast_th, vo.getFASTThreshold() + 5));
                vo.resetORBThreshold();
                req.repeat = false;
            }
        }
    } while (req.repeat);
    return calls;
}

// ---- C shims for the CPU test-suite (ctypes) -----------------------------------------------------------------------
extern "C" {
void srba_da_filter_direction(const float* this_xy, const float* other_xy, int n, float offset, int* status)
{ direction_core(this_xy, other_xy, n, offset, status); }
void srba_da_filter_orb_distance(const int* query_idx, const int* train_idx, const float* distance, int n, int n_other, double max_dist, int* status)
{ orb_distance_core(query_idx, train_idx, distance, n, n_other, max_dist, status); }
void srba_project_match_to_3d(double ul, double vl, double ur, const double* cam6, double* xyz)
{ TStereoCameraParams c = { cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5] }; projectMatchTo3D(ul, vl, ur, c, xyz[0], xyz[1], xyz[2]); }
// ids: tracking_first [nkf][nmatches] (int64, -1 = INVALID), kf_ids [nkf], per-kf id arrays flattened with offsets
void srba_da_assign_ids(int nmatches, int nkf, const long long* tracking_first, const int* da_kf_idx, const int* sorted,
                        const long long* kf_ids_flat, const int* kf_ids_ofs, int n_keyframes, long long* last_match_id, long long* out_ids,
                        long long* n_tracked, long long* n_new)
{
    t_vector_kf kfs(n_keyframes);
    for (int k = 0; k < n_keyframes; ++k) for (int i = kf_ids_ofs[k]; i < kf_ids_ofs[k + 1]; ++i) kfs[k].m_matches_ID.push_back((size_t)kf_ids_flat[i]);
    TVectorKfsDaInfo da(nkf);
    for (int k = 0; k < nkf; ++k) {
        da[k].kf_idx = (size_t)da_kf_idx[k]; da[k].tracking_info.resize(nmatches);
        for (int m = 0; m < nmatches; ++m) {
            const long long f = tracking_first[(size_t)k * nmatches + m];
            da[k].tracking_info[m] = std::make_pair(f < 0 ? INVALID_IDX : (size_t)f, 0.0);
        }
    }
    CStereoSLAMKF nk; nk.m_matches.resize(nmatches);
    std::vector<size_t> srt(sorted, sorted + nkf);
    size_t last = (size_t)*last_match_id;
    const std::pair<size_t, size_t> r = assignMatchIds(nk, da, srt, kfs, last);
    for (int m = 0; m < nmatches; ++m) out_ids[m] = (long long)nk.m_matches_ID[m];
    *last_match_id = (long long)last; *n_tracked = (long long)r.first; *n_new = (long long)r.second;
}
}
