#include "stereo_odometry_b200.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <set>
#include <sstream>

namespace rso {

CStereoOdometryEstimator::CStereoOdometryEstimator() {}
CStereoOdometryEstimator::~CStereoOdometryEstimator() { if (m_ctx) sfe_destroy(m_ctx); }

static std::string trim(const std::string& s)
{
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

// Plain MRPT-style .ini: [SECTION], key = value // comment.  Keys of Appendix B that parameterise the hot path.
void CStereoOdometryEstimator::loadParamsFromConfigFile(const std::string& ini_file, const std::vector<std::string>& sections)
{
    std::ifstream f(ini_file.c_str());
    std::string line, sec;
    auto wanted = [&](const std::string& s) { return std::find(sections.begin(), sections.end(), s) != sections.end(); };
    while (std::getline(f, line)) {
        line = trim(line.substr(0, line.find("//")));
        if (line.empty() || line[0] == ';' || line[0] == '#') continue;
        if (line[0] == '[') { sec = line.substr(1, line.find(']') - 1); continue; }
        const size_t eq = line.find('=');
        if (eq == std::string::npos || !wanted(sec)) continue;
        const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        const double num = atof(val.c_str()); const bool flag = val == "true" || val == "1";
        if (sec == "DETECT") {
            if (key == "initial_FAST_threshold") params_detect.initial_FAST_threshold = (int)num;
            else if (key == "fast_min_th") params_detect.fast_min_th = (int)num;
            else if (key == "fast_max_th") params_detect.fast_max_th = (int)num;
            else if (key == "orb_nfeats") params_detect.orb_nfeats = (int)num;
            else if (key == "orb_nlevels") params_detect.orb_nlevels = (int)num;
        } else if (sec == "MATCH") {
            if (key == "max_y_diff") params_lr_match.max_y_diff = num;
            else if (key == "enable_robust_1to1_match") params_lr_match.enable_robust_1to1_match = flag;
            else if (key == "rectified_images") params_lr_match.rectified_images = flag;
            else if (key == "orb_max_distance") { params_lr_match.orb_max_distance = num; params_if_match.orb_max_distance = num; }
            else if (key == "orb_min_th") params_lr_match.orb_min_th = (int)num;
            else if (key == "orb_max_th") params_lr_match.orb_max_th = (int)num;
        }
    }
    m_fast_th = params_detect.initial_FAST_threshold;
    m_orb_th = (int)params_lr_match.orb_max_distance;
}

void CStereoOdometryEstimator::dumpToConsole() const
{
    std::cout << "[b200 stereo front-end] ORB nfeats=" << params_detect.orb_nfeats << " nlevels=" << params_detect.orb_nlevels
              << " FAST th=" << m_fast_th << " | L/R row-band max_y_diff=" << params_lr_match.max_y_diff << " ORB th=" << m_orb_th
              << " 1to1=" << params_lr_match.enable_robust_1to1_match << " | " << sfe_version() << std::endl;
}

bool CStereoOdometryEstimator::ensureContext(int width, int height)
{
    if (m_ctx && width == m_w && height == m_h) return true;
    if (m_ctx) { sfe_destroy(m_ctx); m_ctx = nullptr; }
    if (params_detect.detect_method != TDetectParams::dmORB || params_lr_match.match_method != TLeftRightMatchParams::smDescRbR ||
        params_if_match.ifm_method != TInterFrameMatchingParams::ifmDescBF) {
        m_err = "only the modes the reference forces are implemented: dmORB / smDescRbR / ifmDescBF (EST.cpp:1135-1137)";
        return false;
    }
    sfe_config cfg; sfe_default_config(&cfg);
    cfg.width = width; cfg.height = height;
    cfg.nlevels = params_detect.orb_nlevels; cfg.nfeatures = params_detect.orb_nfeats;
    cfg.fast_threshold = m_fast_th; cfg.orb_max_distance = m_orb_th;
    cfg.max_y_diff = (float)params_lr_match.max_y_diff; cfg.robust_1to1 = params_lr_match.enable_robust_1to1_match ? 1 : 0;
    cfg.min_disparity = (float)params_lr_match.min_disparity; cfg.max_disparity = (float)params_lr_match.max_disparity;
    cfg.max_pairs = 1;
    if (sfe_create(&cfg, &m_ctx) != SFE_OK) { m_err = sfe_last_error(nullptr); m_ctx = nullptr; return false; }
    m_w = width; m_h = height;
    return true;
}

void CStereoOdometryEstimator::setFASTThreshold(int th)
{
    m_fast_th = std::max(params_detect.fast_min_th, std::min(th, params_detect.fast_max_th));
    if (m_ctx) sfe_set_fast_threshold(m_ctx, m_fast_th);
}
void CStereoOdometryEstimator::setORBThreshold(int th)
{
    m_orb_th = std::max(params_lr_match.orb_min_th, std::min(th, params_lr_match.orb_max_th));
    if (m_ctx) sfe_set_orb_threshold(m_ctx, m_orb_th);
}

bool CStereoOdometryEstimator::processNewImagePair(TStereoOdometryRequest& req, TStereoOdometryResult& res)
{
    res = TStereoOdometryResult();
    if (!req.repeat) {
        if (!req.stereo_imgs || !req.stereo_imgs->imageLeft || !req.stereo_imgs->imageRight) { m_err = "no images in request"; return false; }
        if (!ensureContext(req.stereo_imgs->width, req.stereo_imgs->height)) return false;
    } else if (!m_ctx) { m_err = "repeat without a previous frame"; return false; }
    sfe_set_fast_threshold(m_ctx, m_fast_th);
    sfe_set_orb_threshold(m_ctx, m_orb_th);
    const int rc = req.repeat ? sfe_process_stereo(m_ctx, nullptr, nullptr, 0, 1)
                              : sfe_process_stereo(m_ctx, req.stereo_imgs->imageLeft, req.stereo_imgs->imageRight, req.stereo_imgs->stride, 0);
    if (rc != SFE_OK) { m_err = sfe_last_error(m_ctx); return false; }

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
    // A new frame (repeat == false) first retires the last call's result to "previous frame"; request.repeat re-runs
    // detection on the same images (EST.cpp:278,285,315), so it must be matched against that SAME previous frame
    // again and may not keep the ids its discarded first attempt consumed.
    if (!req.repeat) {
        if (m_have_frame) { m_prev_matched_desc = m_cur_matched_desc; m_prev_ids = m_ids; }
        m_next_id_at_frame_start = m_next_id;
    } else {
        m_next_id = m_next_id_at_frame_start;
    }
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
    for (int k = 0; k < nm; ++k) {
        bool ok = k < nifm && ifm[k].distance <= (float)params_if_match.orb_max_distance && !claimed[ifm[k].trainIdx];
        if (ok) { ids[k] = m_prev_ids[ifm[k].trainIdx]; claimed[ifm[k].trainIdx] = 1; ++tracked; if (kfset.count(ids[k])) ++tracked_kf; }
        else ids[k] = m_next_id++;
    }
    m_ids = ids; m_cur_matched_desc = cur; m_have_frame = true;

    res.detected_feats.assign(1, std::make_pair((size_t)cnt.n_left, (size_t)cnt.n_right));
    res.stereo_matches.assign(1, (size_t)cnt.n_matches);
    res.tracked_feats_from_last_frame = tracked; res.tracked_feats_from_last_KF = tracked_kf;
    res.valid = cnt.n_matches >= 8;          // the reference needs >= 8 stereo matches to go on (EST.cpp:288-295)
    if (m_verbose > 1) std::cout << "[b200] L " << cnt.n_left << " R " << cnt.n_right << " matches " << cnt.n_matches << " tracked " << tracked << std::endl;
    return res.valid;
}

void CStereoOdometryEstimator::setThisFrameAsKF() { m_kf_ids = m_ids; }
void CStereoOdometryEstimator::resetIds() { m_kf_ids = m_ids; }

void CStereoOdometryEstimator::getValues(std::vector<cv::KeyPoint>& kpL, std::vector<cv::KeyPoint>& kpR, cv::Mat& descL, cv::Mat& descR,
                                         std::vector<cv::DMatch>& matches, std::vector<size_t>& matches_ID)
{
    kpL = m_kpL; kpR = m_kpR; descL = m_descL; descR = m_descR; matches = m_matches; matches_ID = m_ids;
}

bool CStereoOdometryEstimator::getChangeInPose(const std::vector<std::pair<size_t, size_t>>&, const std::vector<cv::DMatch>&,
                                               const std::vector<cv::DMatch>&, const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&,
                                               const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&, const TStereoCameraLite&,
                                               TStereoOdometryResult& result, const std::vector<double>&)
{
    result.valid = false;
    m_err = "getChangeInPose: the least-squares pose solver stays with the host pipeline (out of scope, BASELINE.json)";
    return false;
}

}  // namespace rso
