#include "stereo_slam_frontend.h"
#include <algorithm>
#include <deque>
#include <chrono>
#include <fstream>

using mrpt::poses::CPose3D;
using mrpt::poses::CPose3DRotVec;

__attribute__((weak)) TGeneralOptions general_options;          // the application defines the strong one (MAIN.cpp:26)

// =====================================================================================================================
// CSubmapGraphBackend
// =====================================================================================================================
void CSubmapGraphBackend::configure(size_t submap_size, size_t max_tree_depth, size_t max_optimize_depth, const CPose3D& sensor_pose_on_robot)
{
    m_submap = std::max<size_t>(submap_size, 1); m_tree_depth = max_tree_depth; m_opt_depth = max_optimize_depth;
    m_sensor = CPose3DRotVec(sensor_pose_on_robot); m_sensor_inv = m_sensor.getInverse();
}

size_t CSubmapGraphBackend::bfs(TKeyFrameID from, TKeyFrameID to, size_t max_depth) const
{
    const size_t INF = std::numeric_limits<size_t>::max();
    if (from >= m_adj.size() || to >= m_adj.size()) return INF;
    if (from == to) return 0;
    std::vector<size_t> dist(m_adj.size(), INF); std::deque<TKeyFrameID> q;
    dist[from] = 0; q.push_back(from);
    while (!q.empty()) {
        const TKeyFrameID u = q.front(); q.pop_front();
        if (dist[u] >= max_depth) continue;
        for (TKeyFrameID v : m_adj[u]) if (dist[v] == INF) { dist[v] = dist[u] + 1; if (v == to) return dist[v]; q.push_back(v); }
    }
    return INF;
}

TKeyFrameID CSubmapGraphBackend::define_new_keyframe(const std::vector<TObservation>& listObs, bool /*optimize*/, const CPose3D& initial_kf_pose,
                                                     std::vector<std::pair<TKeyFrameID, TKeyFrameID>>& created_edges)
{
    created_edges.clear();
    const TKeyFrameID id = m_pose.size();
    m_adj.push_back(std::vector<TKeyFrameID>());
    auto link = [&](TKeyFrameID a, TKeyFrameID b) { m_adj[a].push_back(b); m_adj[b].push_back(a); m_edges.push_back(std::make_pair(a, b)); created_edges.push_back(std::make_pair(a, b)); };
    auto centre = [&](TKeyFrameID k) { return m_submap * (k / m_submap); };
    if (id == 0) m_pose.push_back(CPose3D());
    else {
        // dead reckoning: robot pose = previous robot pose (+) K (+) camera increment (+) K^-1
        const CPose3DRotVec step = m_sensor + CPose3DRotVec(initial_kf_pose) + m_sensor_inv;
        m_pose.push_back(CPose3D(CPose3DRotVec(m_pose[id - 1]) + step));
        if (id % m_submap != 0) link(centre(id), id);                    // ordinary keyframe: one edge to its local-map centre
        else {                                                           // new centre: vote the areas its landmarks come from
            std::map<TKeyFrameID, size_t> votes;
            for (const TObservation& o : listObs) { auto it = m_base_kf.find(o.feat_id); if (it != m_base_kf.end()) votes[centre(it->second)]++; }
            std::vector<std::pair<size_t, TKeyFrameID>> sorted;
            for (const auto& v : votes) sorted.push_back(std::make_pair(v.second, v.first));
            std::stable_sort(sorted.begin(), sorted.end(), [](const std::pair<size_t, TKeyFrameID>& a, const std::pair<size_t, TKeyFrameID>& b) { return a.first > b.first; });
            for (const auto& s : sorted)
                if (bfs(id, s.second, m_tree_depth) >= m_opt_depth && s.first >= min_obs_to_loop_closure) link(s.second, id);
            if (created_edges.empty()) link(centre(id - 1), id);         // the reference asserts here; the stand-in keeps the chain connected
        }
    }
    for (const TObservation& o : listObs) m_base_kf.insert(std::make_pair(o.feat_id, id));
    return id;
}

size_t CSubmapGraphBackend::topological_distance(TKeyFrameID from, TKeyFrameID to) const { return bfs(from, to, m_tree_depth); }

bool CSubmapGraphBackend::spanning_tree_pose(TKeyFrameID root, TKeyFrameID id, CPose3D& out) const
{
    if (bfs(root, id, m_tree_depth) == std::numeric_limits<size_t>::max()) return false;
    out = CPose3D(CPose3DRotVec(m_pose[root]).getInverse() + CPose3DRotVec(m_pose[id]));
    return true;
}
bool CSubmapGraphBackend::keyframe_global_pose(TKeyFrameID id, CPose3D& out) const
{
    if (id >= m_pose.size()) return false;
    out = m_pose[id];
    return true;
}

// =====================================================================================================================
// CSRBAStereoSLAMEstimator
// =====================================================================================================================
CSRBAStereoSLAMEstimator::CSRBAStereoSLAMEstimator()
    : m_last_kf_ID(0), m_last_match_ID(0), m_last_num_tracked_feats((size_t)UNINITIALIZED_TRACKED_NUMBER), m_max_rotation(0), m_max_translation(0),
      m_max_rotation_limit(0), m_max_translation_limit(0), m_source(nullptr), m_rba(&m_default_backend), m_count(0), m_end_app(false), m_f_da(nullptr), m_f_ls(nullptr)
{
}
CSRBAStereoSLAMEstimator::~CSRBAStereoSLAMEstimator() { if (m_f_da) fclose(m_f_da); if (m_f_ls) fclose(m_f_ls); }

void CSRBAStereoSLAMEstimator::initialize(const mrpt::utils::CConfigFile& config)
{
    m_bow_manager.loadVocabularyFromConfigFile(config, "SRBA_GENERAL", "voc_filename");       // throws if the file is missing

    // camera on the robot: image axes (Z forwards, Y downwards, X to the right) wrt the robot frame
    std::vector<double> p(6);
    config.read_vector("GENERAL", "camera_pose_on_robot", std::vector<double>(6, 0), p, false);
    if (p.size() != 6) p.assign(6, 0.0);
    const CPose3D camera_pose_on_robot(p[0], p[1], p[2], DEG2RAD(p[3]), DEG2RAD(p[4]), DEG2RAD(p[5]));
    const CPose3D img_to_camera_pose(0, 0, 0, DEG2RAD(-90), 0, DEG2RAD(-90));
    CPose3D image_pose_on_robot; image_pose_on_robot.composeFrom(camera_pose_on_robot, img_to_camera_pose);
    const CPose3DRotVec camera_pose_on_robot_rvt(image_pose_on_robot);
    m_current_pose = camera_pose_on_robot_rvt;

    srba_options.loadFromConfigFile(config);
    srba_options.camera_pose_on_robot_rvt = camera_pose_on_robot_rvt;
    srba_options.camera_pose_on_robot_rvt_inverse = camera_pose_on_robot_rvt.getInverse();

    std::vector<std::string> paramSections;
    for (const char* s : { "RECTIFY", "DETECT", "MATCH", "IF-MATCH", "LEAST_SQUARES", "GUI", "GENERAL" }) paramSections.push_back(s);
    m_voEngine.loadParamsFromConfigFile(config, paramSections);
    m_voEngine.setVerbosityLevel(general_options.verbose_level);
    m_voEngine.params_detect.detect_method = rso::CStereoOdometryEstimator::TDetectParams::dmORB;
    m_voEngine.params_lr_match.match_method = rso::CStereoOdometryEstimator::TLeftRightMatchParams::smDescRbR;
    m_voEngine.params_if_match.ifm_method = rso::CStereoOdometryEstimator::TInterFrameMatchingParams::ifmDescBF;
    m_voEngine.params_detect.initial_FAST_threshold = srba_options.detect_fast_th;
    m_voEngine.params_detect.orb_nlevels = (int)srba_options.n_levels;
    m_voEngine.params_detect.orb_nfeats = (int)srba_options.n_feats;
    if (general_options.verbose_level >= 1) m_voEngine.dumpToConsole();

    m_rba->configure(srba_options.srba_submap_size, srba_options.srba_max_tree_depth, srba_options.srba_max_optimize_depth, image_pose_on_robot);

    m_max_rotation_limit = srba_options.max_rotation; m_max_translation_limit = srba_options.max_translation;
    m_max_rotation = 2 * m_max_rotation_limit; m_max_translation = 2 * m_max_translation_limit;
    m_odom_request.stereo_cam = srba_options.stereo_camera;
}

void CSRBAStereoSLAMEstimator::performStereoSLAM()
{
    if (!m_source) throw std::logic_error("CSRBAStereoSLAMEstimator: no image source set");
    if (general_options.debug && !m_f_da) { srba_prepare_out_dir(); m_f_da = fopen("da.txt", "wt"); m_f_ls = fopen("ls.txt", "wt"); }
    mrpt::obs::CObservationStereoImagesPtr stImgs;
    while ((stImgs = m_source->getNextFrame()).present() && !m_end_app) {
        VERBOSE_LEVEL(1) << std::endl << " >> Frame # " << m_count << std::endl;
        if (m_keyframes.empty()) processFirstFrame(stImgs);
        else if (!processNextFrame(stImgs)) { m_count++; continue; }     // invalid odometry: the frame is skipped before the stop checks (EST.cpp:318-323)
        m_frame_pose.push_back(m_current_pose);
        m_count++;
        if ((general_options.max_num_kfs > 0 && m_keyframes.size() == (size_t)general_options.max_num_kfs) ||
            (general_options.to_step != 0 && m_count >= (size_t)general_options.to_step))
            m_end_app = true;
    }
    writeShutdownFiles();
}

// time_new_kf.txt: "<ms in define_new_keyframe> <observations>" per keyframe (EST.cpp:939-942); out_kf_poses.txt: "kf x y z yaw pitch roll",
// %.3f (EST.cpp:977-986) -- the reference writes the poses of its final global pose-graph optimisation there, here they are the
// backend's keyframe poses (no optimiser behind the stand-in backend: dead reckoning).
void CSRBAStereoSLAMEstimator::writeShutdownFiles()
{
    if (!general_options.debug) return;
    if (FILE* f = fopen("time_new_kf.txt", "wt")) {
        for (const auto& st : m_stats) fprintf(f, "%.3f %d\n", 1000.0 * st.first, (int)st.second);
        fclose(f);
    }
    if (FILE* f = fopen("out_kf_poses.txt", "wt")) {
        for (size_t kf = 0; kf < m_keyframes.size(); ++kf) {
            CPose3D p;
            if (!m_rba->keyframe_global_pose(kf, p)) continue;
            double yaw, pitch, roll; p.getYawPitchRoll(yaw, pitch, roll);
            fprintf(f, "%d %.3f %.3f %.3f %.3f %.3f %.3f\n", (int)kf, p.x(), p.y(), p.z(), yaw, pitch, roll);
        }
        fclose(f);
    }
}

void CSRBAStereoSLAMEstimator::processFirstFrame(const mrpt::obs::CObservationStereoImagesPtr& stImgs)
{
    m_keyframes.push_back(CStereoSLAMKF());
    CStereoSLAMKF& new_kf = m_keyframes.back();
    new_kf.setKFID(m_last_kf_ID++);

    m_odom_request.stereo_imgs = stImgs;
    m_voEngine.processNewImagePair(m_odom_request, m_odom_result);
    m_voEngine.setThisFrameAsKF();                      // virtual keyframe in the odometer: id tracking starts here
    new_kf.getDataFromVOEngine(m_voEngine);
    if (new_kf.m_matches_ID.empty()) throw std::runtime_error("first frame has no stereo matches: " + m_voEngine.lastError());
    m_last_match_ID = new_kf.m_matches_ID.back();
    m_bow_manager.insertIntoDB(new_kf);
    VERBOSE_LEVEL(1) << "	# feats (L/R) = " << new_kf.m_keypoints_left.size() << "/" << new_kf.m_keypoints_right.size() << " -- # matches = " << new_kf.m_matches.size() << std::endl;
    if (general_options.verbose_level >= 2) new_kf.dumpToConsole();
    if (general_options.debug) new_kf.saveInfoToFiles();

    std::vector<TObservation> listObs;
    assembleObservations(new_kf, srba_options, listObs);
    std::vector<std::pair<TKeyFrameID, TKeyFrameID>> edges;
    const auto t0 = std::chrono::steady_clock::now();
    m_rba->define_new_keyframe(listObs, false, CPose3D(), edges);          // not optimised the first time
    m_stats.push_back(std::make_pair(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), listObs.size()));
    m_obs_log.push_back(listObs); m_kf_frame.push_back(m_count);
    m_last_kf_pose = new_kf.m_camera_pose;
}

bool CSRBAStereoSLAMEstimator::processNextFrame(const mrpt::obs::CObservationStereoImagesPtr& stImgs)
{
    VERBOSE_LEVEL(1) << " --> Looking for KF#" << m_last_kf_ID << std::endl;
    m_odom_request.stereo_imgs = stImgs; m_odom_request.use_precomputed_data = false; m_odom_request.repeat = false;
    TAdaptiveThOptions aopt;
    aopt.orb_adaptive_fast_th = srba_options.orb_adaptive_fast_th; aopt.adaptive_th_min_matches = srba_options.adaptive_th_min_matches;
    aopt.detect_fast_th = srba_options.detect_fast_th;
    processWithAdaptiveThresholds(m_voEngine, m_odom_request, m_odom_result, aopt);                    // EST.cpp:263-315
    if (!m_odom_result.valid) { VERBOSE_LEVEL(1) << "	[Warning - VO Engine] -- Not a valid result! Skipping this frame." << std::endl; return false; }

    const CPose3DRotVec incr_pose(m_odom_result.outPose);              // current stereo frame wrt the previous one
    m_current_pose += incr_pose; m_incr_pose_from_last_kf += incr_pose; m_incr_pose_from_last_check += incr_pose;
    VERBOSE_LEVEL(1) << "	[VO] # tracked features from last frame: " << m_odom_result.tracked_feats_from_last_frame << std::endl
                     << "	[VO] # tracked features from last KF: " << m_odom_result.tracked_feats_from_last_KF << std::endl;

    // far enough from the last keyframe to force a new one?
    const bool voForceNewKf = m_incr_pose_from_last_kf.m_coords.norm() > m_max_translation_limit ||
                              m_incr_pose_from_last_kf.m_rotvec.norm() > DEG2RAD(m_max_rotation_limit);
    // too few features tracked since the last keyframe -> check
    const bool voForceCheckTracking = srba_options.vo_id_tracking_th == 0 ? false : int(m_odom_result.tracked_feats_from_last_KF) < srba_options.vo_id_tracking_th;
    // far enough from the last check (adaptive limits) -> check
    const bool voForceCheckDistance = m_incr_pose_from_last_check.m_coords.norm() > m_max_translation ||
                                      m_incr_pose_from_last_check.m_rotvec.norm() > DEG2RAD(m_max_rotation);
    if (voForceNewKf || voForceCheckTracking || voForceCheckDistance) checkForNewKeyframe(voForceNewKf, voForceCheckTracking, voForceCheckDistance);
    return true;
}

bool CSRBAStereoSLAMEstimator::m_get_similar_kfs(const TKeyFrameID& newKfId, const QueryResults& dbQueryResults, TLoopClosureInfo& out)
{
    return getSimilarKfs(newKfId, dbQueryResults, srba_options, srba_options.srba_submap_size, m_incr_pose_from_last_kf,
                         [this](TKeyFrameID a, TKeyFrameID b) { return m_rba->topological_distance(a, b); },
                         [this](TKeyFrameID r, TKeyFrameID i, CPose3D& o) { return m_rba->spanning_tree_pose(r, i, o); }, out);
}
void CSRBAStereoSLAMEstimator::m_data_association(const CStereoSLAMKF& kf, const TLoopClosureInfo& lc_info, TVectorKfsDaInfo& out_da)
{
    dataAssociation(m_voEngine, srba_options, m_keyframes, kf, lc_info, out_da);
}

void CSRBAStereoSLAMEstimator::checkForNewKeyframe(bool voForceNewKf, bool voForceCheckTracking, bool voForceCheckDistance)
{
    VERBOSE_LEVEL(1) << "	[VO Check] -- Visual odometry asked for CHECKING for a new keyframe." << std::endl;
    TKfCheckRecord rec; rec.frame = m_count; rec.vo_force_new_kf = voForceNewKf; rec.vo_force_check_tracking = voForceCheckTracking; rec.vo_force_check_distance = voForceCheckDistance;
    if (voForceCheckTracking) m_voEngine.resetIds();     // the previous frame becomes the odometer's virtual keyframe: no check storm (EST.cpp:398-404)
    m_incr_pose_from_last_check = CPose3DRotVec();

    m_keyframes.push_back(CStereoSLAMKF());             // candidate, dropped again if not needed
    CStereoSLAMKF& new_kf = m_keyframes.back();
    new_kf.getDataFromVOEngine(m_voEngine);
    new_kf.setKFID(m_last_kf_ID);
    rec.candidate_id = new_kf.m_kf_ID;

    QueryResults qResults;
    m_bow_manager.queryDB(new_kf, qResults, 4);
    ASSERT_(qResults.size() > 0)
    for (const DBoW2::Result& r : qResults) rec.query.push_back(std::make_pair((unsigned)r.Id, r.Score));

    TLoopClosureInfo similar_kfs_info;
    const bool potentialLoopClosure = m_get_similar_kfs(new_kf.m_kf_ID, qResults, similar_kfs_info);
    rec.similar_kfs = similar_kfs_info.similar_kfs; rec.lc_id = similar_kfs_info.lc_id; rec.potential_lc = potentialLoopClosure;
    bool insertNewKf = qResults[0].Score < 0.05;        // nothing similar in the database: a new keyframe is needed

    TVectorKfsDaInfo daInfo;
    m_data_association(new_kf, similar_kfs_info, daInfo);
    const size_t numberSimilarKfs = daInfo.size();
    ASSERT_(numberSimilarKfs > 0)
    for (const t_kf_da_info& d : daInfo) rec.tracked.push_back(std::make_pair(d.kf_idx, d.tracked_matches));

    // keyframes by decreasing number of commonly observed features (DATrackedSorter, CStereoSLAMKF.h:119-126)
    std::vector<size_t> sortedIndices(numberSimilarKfs);
    for (size_t i = 0; i < numberSimilarKfs; i++) sortedIndices[i] = i;
    std::stable_sort(sortedIndices.begin(), sortedIndices.end(), [&](size_t a, size_t b) { return daInfo[a].tracked_matches > daInfo[b].tracked_matches; });
    const size_t highestNumberTrackedFeats = daInfo[sortedIndices[0]].tracked_matches;
    m_last_num_tracked_feats = highestNumberTrackedFeats;
    if (voForceNewKf) insertNewKf = true;

    bool confirmedLoopClosure = false;
    if (potentialLoopClosure) {
        // confirmed when the old keyframe tracks more than half of the best count
        size_t loopClosureIdx = 0;
        for (size_t i = 0; !confirmedLoopClosure && i < daInfo.size(); i++) {
            confirmedLoopClosure = daInfo[i].kf_idx == similar_kfs_info.lc_id && daInfo[i].tracked_matches > 0.5 * highestNumberTrackedFeats;
            if (confirmedLoopClosure) loopClosureIdx = i;
        }
        if (confirmedLoopClosure) {
            m_rba->loopClosureDetected(); m_rba->setLoopClosureOldID(similar_kfs_info.lc_id);
            insertNewKf = true;
            sortedIndices.erase(std::find(sortedIndices.begin(), sortedIndices.end(), loopClosureIdx));   // the old keyframe gets priority in the id assignment
            sortedIndices.insert(sortedIndices.begin(), loopClosureIdx);
        } else m_rba->loopClosureDetected(false);
    } else {
        if (highestNumberTrackedFeats < srba_options.updated_matches_th) insertNewKf = true;
        else {
            const size_t olimit = srba_options.updated_matches_th + srba_options.up_matches_th_plus;
            if (highestNumberTrackedFeats <= olimit) {      // still plenty tracked: tighten the motion limits for the next check instead
                m_max_translation = updateTranslationThreshold(highestNumberTrackedFeats - srba_options.updated_matches_th, srba_options.up_matches_th_plus);
                m_max_rotation = updateRotationThreshold(highestNumberTrackedFeats, olimit);
            }
        }
        m_rba->loopClosureDetected(false);
    }
    rec.confirmed_lc = confirmedLoopClosure; rec.inserted = insertNewKf;
    if (!insertNewKf) m_keyframes.resize(m_keyframes.size() - 1);
    else insertKeyframe(new_kf, daInfo, sortedIndices, highestNumberTrackedFeats, confirmedLoopClosure, qResults, rec);
    m_checks.push_back(rec);
}

void CSRBAStereoSLAMEstimator::insertKeyframe(CStereoSLAMKF& new_kf, const TVectorKfsDaInfo& daInfo, const std::vector<size_t>& sortedIndices,
                                              size_t highestNumberTrackedFeats, bool confirmedLoopClosure, const QueryResults& qResults, TKfCheckRecord& rec)
{
    const std::string sep = general_options.out_dir + SRBA_PATH_SEP;
    std::vector<int> vectorNumberTrackedFeats; std::vector<double> da_dist;
    const std::pair<size_t, size_t> n = assignMatchIds(new_kf, daInfo, sortedIndices, m_keyframes, m_last_match_ID, &vectorNumberTrackedFeats, &da_dist);   // EST.cpp:572-617
    rec.n_tracked_ids = n.first; rec.n_new_ids = n.second;
    m_voEngine.resetIds();
    size_t realNumberSimilarKfs = 0;
    for (size_t m = 0; m < sortedIndices.size(); ++m) { if (vectorNumberTrackedFeats[sortedIndices[m]] >= 0.5 * highestNumberTrackedFeats) realNumberSimilarKfs++; else break; }
    VERBOSE_LEVEL(1) << "	Real number of similar kfs: " << realNumberSimilarKfs << "/" << sortedIndices.size() << " tracked " << n.first << " new " << n.second << std::endl;
    if (general_options.debug) {
        srba_prepare_out_dir();
        if (FILE* f = fopen((sep + mrpt::format("da_dist_kf%04d.txt", (int)new_kf.m_kf_ID)).c_str(), "wt")) {      // "%2.f" for tracked, "0.00" for new (EST.cpp:602, 614)
            for (double d : da_dist) { if (d >= 0.0) fprintf(f, "%2.f\n", d); else fprintf(f, "0.00\n"); }
            fclose(f);
        }
        if (m_f_da) for (size_t k = 0; k < sortedIndices.size(); ++k)
            fprintf(m_f_da, "%d %d %d\n", (int)new_kf.m_kf_ID, (int)daInfo[sortedIndices[k]].kf_idx, vectorNumberTrackedFeats[sortedIndices[k]]);
    }
    m_last_num_tracked_feats = (size_t)UNINITIALIZED_TRACKED_NUMBER;
    m_last_kf_ID++;
    m_max_translation = srba_options.max_translation; m_max_rotation = srba_options.max_rotation;       // back to the configured limits (EST.cpp:655-656)

    std::vector<TObservation> listObs;
    assembleObservations(new_kf, srba_options, listObs);                                                 // EST.cpp:666-699
    if (general_options.debug) {
        new_kf.saveInfoToFiles();
        std::ofstream fstr((sep + mrpt::format("da_info_%04d.txt", (int)new_kf.m_kf_ID)).c_str());         // EST.cpp:749-764
        fstr.setf(std::ios::fixed, std::ios::floatfield); fstr.precision(2);
        for (const TObservation& o : listObs) fstr << o.feat_id << " " << o.lx << " " << o.ly << " " << o.rx << " " << o.ry << std::endl;
        if (confirmedLoopClosure && !qResults.empty() && qResults[0].Id < m_keyframes.size()) {           // loop_closure_info_<frame>.txt, EST.cpp:676-737
            const CStereoSLAMKF& okf = m_keyframes[qResults[0].Id];
            if (FILE* flc = fopen((sep + mrpt::format("loop_closure_info_%d.txt", (int)m_count)).c_str(), "wt")) {
                for (const TObservation& o : listObs)
                    for (size_t mm = okf.m_matches_ID.size(); mm-- > 0;)                                   // the reference keeps the LAST equal id
                        if (okf.m_matches_ID[mm] == o.feat_id) {
                            const cv::KeyPoint& ol = okf.m_keypoints_left[okf.m_matches[mm].queryIdx]; const cv::KeyPoint& orr = okf.m_keypoints_right[okf.m_matches[mm].trainIdx];
                            fprintf(flc, "%d %.2f %.2f %.2f %.2f %.2f %.2f %.2f %.2f\n", (int)o.feat_id, o.lx, o.ly, o.rx, o.ry, ol.pt.x, ol.pt.y, orr.pt.x, orr.pt.y);
                            break;
                        }
                fclose(flc);
            }
        }
    }
    if (general_options.verbose_level >= 2) new_kf.dumpToConsole();

    const CPose3D inputPose = srba_options.use_initial_pose ? CPose3D(m_incr_pose_from_last_kf) : CPose3D();
    std::vector<std::pair<TKeyFrameID, TKeyFrameID>> edges;
    VERBOSE_LEVEL(1) << "	Inserting " << listObs.size() << " observations in srba-slam engine" << std::endl;
    const auto t0 = std::chrono::steady_clock::now();
    const TKeyFrameID new_id = m_rba->define_new_keyframe(listObs, true, inputPose, edges);             // EST.cpp:773-782
    m_stats.push_back(std::make_pair(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), listObs.size()));
    if (m_f_ls) fprintf(m_f_ls, "%d %.4f\n", (int)new_id, m_rba->last_obs_rmse());                      // ls.txt, EST.cpp:789-790
    m_obs_log.push_back(listObs); m_kf_frame.push_back(m_count);

    CPose3D kf_pose;                                                                                     // EST.cpp:881-916
    if (m_rba->keyframe_global_pose(new_kf.m_kf_ID, kf_pose)) new_kf.m_camera_pose = m_last_kf_pose = CPose3DRotVec(kf_pose);
    m_bow_manager.insertIntoDB(new_kf);
    m_current_pose = m_last_kf_pose + srba_options.camera_pose_on_robot_rvt;
    m_incr_pose_from_last_kf = CPose3DRotVec();
}
