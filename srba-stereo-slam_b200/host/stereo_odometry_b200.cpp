#include "stereo_odometry_b200.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <set>

namespace rso {

static std::mutex g_ctx_mu;
static std::vector<CStereoOdometryEstimator*> g_engines;       // construction order; currentContext() = newest engine with a context

CStereoOdometryEstimator::CStereoOdometryEstimator() { std::lock_guard<std::mutex> lk(g_ctx_mu); g_engines.push_back(this); }
CStereoOdometryEstimator::~CStereoOdometryEstimator()
{
    { std::lock_guard<std::mutex> lk(g_ctx_mu); g_engines.erase(std::remove(g_engines.begin(), g_engines.end(), this), g_engines.end()); }
    if (m_ctx) sfe_destroy(m_ctx);
}
sfe_context* CStereoOdometryEstimator::currentContext()
{
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    for (auto it = g_engines.rbegin(); it != g_engines.rend(); ++it) if ((*it)->m_ctx) return (*it)->m_ctx;
    return nullptr;
}

// cv::BFMatcher(NORM_HAMMING).match of cv_compat.h -> sfe_match_bf on the newest engine's context
static int gpu_brute_force_hook(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out)
{
    sfe_context* ctx = CStereoOdometryEstimator::currentContext();
    if (!ctx) return -1;
    int32_t n = 0;
    static_assert(sizeof(cv::DMatch) == sizeof(sfe_dmatch), "DMatch layout");
    return sfe_match_bf(ctx, q, nq, t, nt, reinterpret_cast<sfe_dmatch*>(out), &n) == SFE_OK ? (int)n : -1;
}

// Keys of SURVEY.md Appendix B that parameterise the hot path; a section is only read if the caller lists it (EST.cpp:1122-1131).
void CStereoOdometryEstimator::loadParamsFromConfigFile(const mrpt::utils::CConfigFile& cfg, const std::vector<std::string>& sections)
{
    auto wanted = [&](const char* s) { return std::find(sections.begin(), sections.end(), std::string(s)) != sections.end(); };
    if (wanted("DETECT")) {
        TDetectParams& d = params_detect;
        d.initial_FAST_threshold = cfg.read_int("DETECT", "initial_FAST_threshold", d.initial_FAST_threshold);
        d.fast_min_th = cfg.read_int("DETECT", "fast_min_th", d.fast_min_th);
        d.fast_max_th = cfg.read_int("DETECT", "fast_max_th", d.fast_max_th);
        d.orb_nfeats = cfg.read_int("DETECT", "orb_nfeats", d.orb_nfeats);
        d.orb_nlevels = cfg.read_int("DETECT", "orb_nlevels", d.orb_nlevels);
        d.min_distance = cfg.read_int("DETECT", "min_distance", d.min_distance);
        d.non_maximal_suppression = cfg.read_bool("DETECT", "non_maximal_suppression", d.non_maximal_suppression);
        d.non_max_supp_method = cfg.read_int("DETECT", "non_max_supp_method", d.non_max_supp_method);
    }
    if (wanted("MATCH")) {
        TLeftRightMatchParams& m = params_lr_match;
        m.max_y_diff = cfg.read_double("MATCH", "max_y_diff", m.max_y_diff);
        m.enable_robust_1to1_match = cfg.read_bool("MATCH", "enable_robust_1to1_match", m.enable_robust_1to1_match);
        m.rectified_images = cfg.read_bool("MATCH", "rectified_images", m.rectified_images);
        m.orb_max_distance = cfg.read_double("MATCH", "orb_max_distance", m.orb_max_distance);
        m.orb_min_th = cfg.read_int("MATCH", "orb_min_th", m.orb_min_th);
        m.orb_max_th = cfg.read_int("MATCH", "orb_max_th", m.orb_max_th);
        params_if_match.orb_max_distance = m.orb_max_distance;       // "the same values as for stereo-matching are employed"
    }
    // example.ini:64 spells the section [LEAST-SQUARES]; the reference asks for LEAST_SQUARES (EST.cpp:1127)
    if (wanted("LEAST_SQUARES")) {
        TLeastSquaresParams& l = params_least_squares;
        l.initial_max_iters = cfg.read_int("LEAST_SQUARES", "initial_max_iters", l.initial_max_iters);
        l.max_iters = cfg.read_int("LEAST_SQUARES", "max_iters", l.max_iters);
        l.max_incr_cost = cfg.read_int("LEAST_SQUARES", "max_incr_cost", l.max_incr_cost);
        l.residual_threshold = cfg.read_double("LEAST_SQUARES", "residual_threshold", l.residual_threshold);
        l.min_mod_out_vector = cfg.read_double("LEAST_SQUARES", "min_mod_out_vector", l.min_mod_out_vector);
        l.bad_tracking_th = cfg.read_int("LEAST_SQUARES", "bad_tracking_th", l.bad_tracking_th);
        l.use_robust_kernel = cfg.read_bool("LEAST_SQUARES", "use_robust_kernel", l.use_robust_kernel);
        l.kernel_param = cfg.read_double("LEAST_SQUARES", "kernel_param", l.kernel_param);
        l.use_previous_pose_as_initial = cfg.read_bool("LEAST_SQUARES", "use_previous_pose_as_initial", l.use_previous_pose_as_initial);
    }
    m_fast_th = params_detect.initial_FAST_threshold;
    m_orb_th = (int)params_lr_match.orb_max_distance;
}
void CStereoOdometryEstimator::loadParamsFromConfigFile(const std::string& ini_file, const std::vector<std::string>& sections)
{
    loadParamsFromConfigFile(mrpt::utils::CConfigFile(ini_file), sections);
}

void CStereoOdometryEstimator::dumpToConsole() const
{
    std::cout << "[b200 stereo front-end] ORB nfeats=" << params_detect.orb_nfeats << " nlevels=" << params_detect.orb_nlevels
              << " FAST th=" << m_fast_th << " | L/R row-band max_y_diff=" << params_lr_match.max_y_diff << " ORB th=" << m_orb_th
              << " 1to1=" << params_lr_match.enable_robust_1to1_match
              << " | [DETECT] min_distance=" << params_detect.min_distance << " non_maximal_suppression=" << params_detect.non_maximal_suppression
              << " (method " << params_detect.non_max_supp_method << "): parsed, not applied under dmORB"
              << " | " << sfe_version() << std::endl;
}

bool CStereoOdometryEstimator::ensureContext(int width, int height)
{
    // EST.cpp:1140-1142 overwrite params_detect AFTER loadParamsFromConfigFile: the context is built from the values in
    // force when the first frame arrives
    if (m_ctx && width == m_w && height == m_h) return true;
    if (m_ctx) { sfe_destroy(m_ctx); m_ctx = nullptr; }
    if (params_detect.detect_method != TDetectParams::dmORB || params_lr_match.match_method != TLeftRightMatchParams::smDescRbR ||
        params_if_match.ifm_method != TInterFrameMatchingParams::ifmDescBF) {
        m_err = "only the modes the reference forces are implemented: dmORB / smDescRbR / ifmDescBF (EST.cpp:1135-1137)";
        return false;
    }
    if (!m_thresholds_touched) { m_fast_th = params_detect.initial_FAST_threshold; }
    sfe_config cfg; sfe_default_config(&cfg);
    cfg.width = width; cfg.height = height;
    cfg.nlevels = params_detect.orb_nlevels; cfg.nfeatures = params_detect.orb_nfeats;
    cfg.fast_threshold = m_fast_th; cfg.orb_max_distance = m_orb_th;
    cfg.max_y_diff = (float)params_lr_match.max_y_diff; cfg.robust_1to1 = params_lr_match.enable_robust_1to1_match ? 1 : 0;
    cfg.min_disparity = (float)params_lr_match.min_disparity; cfg.max_disparity = (float)params_lr_match.max_disparity;
    cfg.max_pairs = 1; cfg.device = device;
    if (sfe_create(&cfg, &m_ctx) != SFE_OK) { m_err = sfe_last_error(nullptr); m_ctx = nullptr; return false; }
    m_w = width; m_h = height;
#ifndef SRBA_B200_WITH_OPENCV
    cv::bruteForceHook() = &gpu_brute_force_hook;
#endif
    return true;
}

void CStereoOdometryEstimator::setFASTThreshold(int th)
{
    m_fast_th = std::max(params_detect.fast_min_th, std::min(th, params_detect.fast_max_th)); m_thresholds_touched = true;
    if (m_ctx) sfe_set_fast_threshold(m_ctx, m_fast_th);
}
void CStereoOdometryEstimator::setORBThreshold(int th)
{
    m_orb_th = std::max(params_lr_match.orb_min_th, std::min(th, params_lr_match.orb_max_th));
    if (m_ctx) sfe_set_orb_threshold(m_ctx, m_orb_th);
}

static TStereoPoseCamera pose_camera(const mrpt::utils::TStereoCamera& c)
{
    // the six numbers projectMatchTo3D reads (srba-stereo-slam_utils.h:564-569)
    TStereoPoseCamera p = { c.leftCamera.cx(), c.leftCamera.cy(), c.leftCamera.fx(), c.rightCamera.cx(), c.rightCamera.fx(), c.rightCameraPose[0] };
    return p;
}

bool CStereoOdometryEstimator::processNewImagePair(TStereoOdometryRequest& req, TStereoOdometryResult& res)
{
    res = TStereoOdometryResult();
    if (!req.repeat) {
        if (!req.stereo_imgs || !req.stereo_imgs->imageLeft.get_unsafe(0, 0) || !req.stereo_imgs->imageRight.get_unsafe(0, 0)) { m_err = "no images in request"; return false; }
        const mrpt::utils::CImage& L = req.stereo_imgs->imageLeft; const mrpt::utils::CImage& R = req.stereo_imgs->imageRight;
        if (L.getWidth() != R.getWidth() || L.getHeight() != R.getHeight() || L.getRowStride() != R.getRowStride()) { m_err = "left/right images differ in size"; return false; }
        if (!ensureContext((int)L.getWidth(), (int)L.getHeight())) return false;
    } else if (!m_ctx) { m_err = "repeat without a previous frame"; return false; }
    sfe_set_fast_threshold(m_ctx, m_fast_th);
    sfe_set_orb_threshold(m_ctx, m_orb_th);
    const int rc = req.repeat ? sfe_process_stereo(m_ctx, nullptr, nullptr, 0, 1)
                              : sfe_process_stereo(m_ctx, req.stereo_imgs->imageLeft.get_unsafe(0, 0), req.stereo_imgs->imageRight.get_unsafe(0, 0),
                                                   req.stereo_imgs->imageLeft.getRowStride(), 0);
    if (rc != SFE_OK) { m_err = sfe_last_error(m_ctx); return false; }

    // A new frame (repeat == false) first retires the last call's result to "previous frame"; request.repeat re-runs
    // detection on the same images (EST.cpp:278,285,315), so it must be matched against that SAME previous frame
    // again and may not keep the ids its discarded first attempt consumed.
    if (!req.repeat) {
        if (m_have_frame) {
            m_prev_matched_desc = m_cur_matched_desc; m_prev_ids = m_ids;
            m_prev_kpL = m_kpL; m_prev_kpR = m_kpR; m_prev_matches = m_matches;
        }
        m_next_id_at_frame_start = m_next_id;
        std::memcpy(m_pose_guess_frame_start, m_last_pose_guess, sizeof(m_last_pose_guess));
    } else {
        m_next_id = m_next_id_at_frame_start;
    }

    // pull the frame to the host in the reference's containers
    const int cap = sfe_keypoint_capacity(m_ctx);
    std::vector<sfe_keypoint> kl(cap), kr(cap); std::vector<uint8_t> dl((size_t)cap * 32), dr((size_t)cap * 32);
    std::vector<sfe_dmatch> mm(cap); sfe_pair_counts cnt;
    if (sfe_get_values(m_ctx, 0, kl.data(), kr.data(), dl.data(), dr.data(), mm.data(), &cnt) != SFE_OK) { m_err = sfe_last_error(m_ctx); return false; }
    m_kpL.resize(cnt.n_left); m_kpR.resize(cnt.n_right); m_matches.resize(cnt.n_matches);
    std::memcpy((void*)m_kpL.data(), kl.data(), sizeof(sfe_keypoint) * cnt.n_left);       // identical 28-byte layout
    std::memcpy((void*)m_kpR.data(), kr.data(), sizeof(sfe_keypoint) * cnt.n_right);
    std::memcpy((void*)m_matches.data(), mm.data(), sizeof(sfe_dmatch) * cnt.n_matches);  // identical 16-byte layout
    m_descL.create(cnt.n_left, 32, CV_8U); m_descR.create(cnt.n_right, 32, CV_8U);
    if (cnt.n_left) std::memcpy(m_descL.ptr(0), dl.data(), (size_t)cnt.n_left * 32);
    if (cnt.n_right) std::memcpy(m_descR.ptr(0), dr.data(), (size_t)cnt.n_right * 32);

    // ---- inter-frame matching (ifmDescBF) on the stereo-matched left descriptors: id propagation ----------------
    const int nm = cnt.n_matches;
    cv::Mat cur(nm, 32, CV_8U);
    for (int k = 0; k < nm; ++k) std::memcpy(cur.ptr(k), m_descL.ptr(m_matches[k].queryIdx), 32);
    std::vector<size_t> ids(nm);
    size_t tracked = 0, tracked_kf = 0;
    std::vector<sfe_dmatch> ifm(std::max(nm, 1)); int32_t nifm = 0;
    const cv::Mat& pre = m_prev_matched_desc;
    if (nm > 0 && pre.rows > 0) {
        if (sfe_match_bf(m_ctx, cur.ptr(0), nm, pre.ptr(0), pre.rows, ifm.data(), &nifm) != SFE_OK) { m_err = sfe_last_error(m_ctx); return false; }
    }
    std::set<size_t> kfset(m_kf_ids.begin(), m_kf_ids.end());
    std::vector<char> claimed(m_prev_ids.size(), 0);
    vector_index_pairs_t tracked_pairs;
    for (int k = 0; k < nm; ++k) {
        bool ok = k < nifm && ifm[k].distance <= (float)params_if_match.orb_max_distance && !claimed[ifm[k].trainIdx];
        if (ok) {
            ids[k] = m_prev_ids[ifm[k].trainIdx]; claimed[ifm[k].trainIdx] = 1; ++tracked; if (kfset.count(ids[k])) ++tracked_kf;
            tracked_pairs.push_back(std::make_pair((size_t)ifm[k].trainIdx, (size_t)k));
        } else ids[k] = m_next_id++;
    }
    m_ids = ids; m_cur_matched_desc = cur;

    res.detected_feats.assign(1, std::make_pair((size_t)cnt.n_left, (size_t)cnt.n_right));
    res.stereo_matches.assign(1, (size_t)cnt.n_matches);
    res.tracked_feats_from_last_frame = tracked; res.tracked_feats_from_last_KF = tracked_kf;
    res.valid = cnt.n_matches >= 8;          // the reference needs >= 8 stereo matches to go on (EST.cpp:288-295)

    // ---- change in pose from the tracked pairs (host least squares, as in stereo-vo) ------------------------------------------
    if (m_have_frame && res.valid) {
        TStereoOdometryResult pr;
        std::vector<double> init;
        if (params_least_squares.use_previous_pose_as_initial) init.assign(m_pose_guess_frame_start, m_pose_guess_frame_start + 6);
        const bool custom = params_least_squares.use_custom_initial_pose;
        params_least_squares.use_custom_initial_pose = !init.empty();
        const bool ok = getChangeInPose(tracked_pairs, m_prev_matches, m_matches, m_prev_kpL, m_prev_kpR, m_kpL, m_kpR, req.stereo_cam, pr, init);
        params_least_squares.use_custom_initial_pose = custom;
        res.valid = ok; res.outPose = pr.outPose; res.num_it = pr.num_it; res.num_it_final = pr.num_it_final; res.out_residual = pr.out_residual;
        if (ok) std::memcpy(m_last_pose_guess, m_last_solver_pose, sizeof(m_last_pose_guess));    // previous wrt current: the next frame's initial guess
    }
    m_have_frame = true;
    if (m_verbose > 1) std::cout << "[b200] L " << cnt.n_left << " R " << cnt.n_right << " matches " << cnt.n_matches << " tracked " << tracked
                                 << " valid " << res.valid << " outPose " << res.outPose << std::endl;
    return res.valid;
}

void CStereoOdometryEstimator::setThisFrameAsKF() { m_kf_ids = m_ids; }
void CStereoOdometryEstimator::resetIds() { m_kf_ids = m_ids; }

void CStereoOdometryEstimator::getValues(std::vector<cv::KeyPoint>& kpL, std::vector<cv::KeyPoint>& kpR, cv::Mat& descL, cv::Mat& descR,
                                         std::vector<cv::DMatch>& matches, std::vector<size_t>& matches_ID)
{
    kpL = m_kpL; kpR = m_kpR; descL = m_descL; descR = m_descR; matches = m_matches; matches_ID = m_ids;
}

bool CStereoOdometryEstimator::getChangeInPose(const vector_index_pairs_t& tracked_pairs, const std::vector<cv::DMatch>& pre_matches,
                                               const std::vector<cv::DMatch>& cur_matches, const std::vector<cv::KeyPoint>& pre_left,
                                               const std::vector<cv::KeyPoint>& pre_right, const std::vector<cv::KeyPoint>& cur_left,
                                               const std::vector<cv::KeyPoint>& cur_right, const mrpt::utils::TStereoCamera& cam,
                                               TStereoOdometryResult& result, const std::vector<double>& initial)
{
    const int n = (int)tracked_pairs.size();
    std::vector<double> pre(3 * (size_t)n), cur(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const cv::DMatch& pm = pre_matches[tracked_pairs[i].first]; const cv::DMatch& cm = cur_matches[tracked_pairs[i].second];
        pre[3 * i] = pre_left[pm.queryIdx].pt.x; pre[3 * i + 1] = pre_left[pm.queryIdx].pt.y; pre[3 * i + 2] = pre_right[pm.trainIdx].pt.x;
        cur[3 * i] = cur_left[cm.queryIdx].pt.x; cur[3 * i + 1] = cur_left[cm.queryIdx].pt.y; cur[3 * i + 2] = cur_right[cm.trainIdx].pt.x;
    }
    TStereoPoseResult r;
    const bool use_init = params_least_squares.use_custom_initial_pose && initial.size() == 6;
    solveStereoChangeInPose(pre.data(), cur.data(), n, pose_camera(cam), use_init ? initial.data() : nullptr, params_least_squares, r);
    result.valid = r.valid; result.num_it = r.num_it; result.num_it_final = r.num_it_final; result.out_residual = r.residual;
    std::memcpy(m_last_solver_pose, r.pose, sizeof(m_last_solver_pose));
    const mrpt::poses::CPose3DRotVec prev_wrt_cur(r.pose[0], r.pose[1], r.pose[2], r.pose[3], r.pose[4], r.pose[5]);
    result.outPose = prev_wrt_cur.getInverse();       // pose of the current frame wrt the previous one (EST.cpp:325-327)
    if (!r.valid) m_err = "getChangeInPose: not enough consistent tracked pairs";
    return r.valid;
}

}  // namespace rso
