// srba_options.h -- the two option blocks of the reference application that parameterise the hot path and its host-side
// callers, with the reference's member names, defaults and .ini keys:
//   TGeneralOptions          [APP_OPTIONS] / [IMG_SOURCE]                      srba-stereo-slam_utils.h:87-216
//   TSRBAStereoSLAMOptions   [SRBA_GENERAL] [SRBA_DETECT] [SRBA_DATA_ASSOCIATION] [SRBA_KF_CREATION] [CAMERA_*]
//                                                                               srba-stereo-slam_utils.h:221-487
// plus the binary option stream of the (disabled) state files, CSRBAStereoSLAMEstimator.cpp:2411-2438, 2562-2589.
#pragma once
#include <iostream>
#include <string>
#include "mrpt_lite.h"

#define INVALID_KF_ID -1
#define UNINITIALIZED_TRACKED_NUMBER -1
typedef size_t TKeyFrameID;

struct TGeneralOptions {
    enum captureSource { csRawlog, csImgDir };
    captureSource cap_src = csImgDir;
    int from_step = 0, to_step = 0, save_at_iteration = 0, max_num_kfs = 0, start_index = 0, end_index = 0, verbose_level = 0;
    bool debug = false, show3D = false, enable_logger = false, load_state_from_file = false, save_state_to_file = false,
         pause_after_show_op = false, pause_at_each_iteration = false;
    std::string out_dir, rawlog_file, state_file, image_dir_url, left_format, right_format;

    void loadFromConfigFile(const mrpt::utils::CConfigFile& config)
    {
        const char* A = "APP_OPTIONS"; const char* I = "IMG_SOURCE";
        pause_after_show_op = config.read_bool(A, "pause_after_show_op", pause_after_show_op);
        out_dir = config.read_string(A, "out_dir", out_dir);
        debug = config.read_bool(A, "debug", debug); show3D = config.read_bool(A, "show3D", show3D);
        enable_logger = config.read_bool(A, "enable_logger", enable_logger);
        verbose_level = config.read_int(A, "verbose_level", verbose_level);
        pause_at_each_iteration = config.read_bool(A, "pause_at_each_iteration", pause_at_each_iteration);
        from_step = config.read_int(A, "from_step", from_step); to_step = config.read_int(A, "to_step", to_step);
        max_num_kfs = config.read_int(A, "max_num_kfs", max_num_kfs);
        save_state_to_file = config.read_bool(A, "save_state_to_file", save_state_to_file);
        save_at_iteration = config.read_int(A, "save_at_iteration", save_at_iteration);
        state_file = config.read_string(A, "state_file", state_file);
        load_state_from_file = save_state_to_file ? false : config.read_bool(A, "load_state_from_file", load_state_from_file);   // saving wins (UTL.h:160-165)
        cap_src = config.read_int(A, "capture_source", cap_src) == 0 ? csRawlog : csImgDir;
        rawlog_file = config.read_string(I, "rawlog_file", rawlog_file);
        image_dir_url = config.read_string(I, "image_dir_url", image_dir_url);
        left_format = config.read_string(I, "left_format", left_format); right_format = config.read_string(I, "right_format", right_format);
        start_index = config.read_int(I, "start_index", start_index); end_index = config.read_int(I, "end_index", end_index);
    }
    void dumpToConsole() const
    {
        std::cout << " Application options: source " << (cap_src == csRawlog ? rawlog_file : image_dir_url) << " [" << start_index << "," << end_index
                  << "] steps " << from_step << ".." << to_step << " max KFs " << max_num_kfs << " debug " << debug << " show3D " << show3D
                  << " out_dir '" << out_dir << "' verbose " << verbose_level << std::endl;
    }
};
extern TGeneralOptions general_options;        // the reference's one global (EST.h:31, MAIN.cpp:26); a weak definition lives in the host library
#define VERBOSE_LEVEL(_LEV) if (general_options.verbose_level >= _LEV) std::cout

// The five mrpt::vision::TMatchingOptions members the option stream carries (EST.cpp:2424-2431).  Member types as in
// MRPT 1.3's mrpt/vision/types.h as far as recalled (MRPT is not vendored: the stream's byte widths for these five are
// unverified here).
struct TMatchingOptionsLite { float epipolar_TH = 1.5f; double maxORB_dist = 60; float max_disp = 1e4f, min_disp = 1.0f; bool enable_robust_1to1_match = false; };

struct TSRBAStereoSLAMOptions {
    enum TDetectMethod { DM_ORB_ONLY = 0, DM_FAST_ORB };
    enum TNonMaxSuppMethod { NMSM_STANDARD = 0, NMSM_ADAPTIVE };
    enum TDAStage2Method { ST2M_NONE = 0, ST2M_FUNDMATRIX, ST2M_CHANGEPOSE, ST2M_BOTH };

    mrpt::utils::TStereoCamera stereo_camera;
    mrpt::poses::CPose3DRotVec camera_pose_on_robot_rvt, camera_pose_on_robot_rvt_inverse;
    TDetectMethod detect_method = DM_ORB_ONLY;
    size_t n_levels = 1, n_feats = 500;                       // n_levels is NOT loaded from the file (UTL.h:414 is commented out)
    int min_ORB_distance = 0, detect_fast_th = 5, adaptive_th_min_matches = 100;
    bool orb_adaptive_fast_th = false;
    TNonMaxSuppMethod non_max_supp_method = NMSM_STANDARD;
    TMatchingOptionsLite matching_options;
    double ransac_fit_prob = 0.95, max_y_diff_epipolar = 1.5, max_orb_distance_da = 60;
    TDAStage2Method da_stage2_method = ST2M_CHANGEPOSE;
    double query_score_th = 0.04;
    bool use_initial_pose = true;
    int vo_id_tracking_th = 40;
    bool da_filter_by_direction = false, da_filter_by_orb_distance = true, da_filter_by_fund_matrix = true, da_filter_by_pose_change = true;
    double residual_th = 50, max_rotation = 15., max_translation = 0.30, srba_kernel_param = 3.0;
    bool non_maximal_suppression = false, pause_after_show_op = false, srba_use_robust_kernel = true, srba_use_robust_kernel_stage1 = true;
    size_t updated_matches_th = 50, up_matches_th_plus = 25, lc_distance = 2, srba_submap_size = 15, srba_max_tree_depth = 3, srba_max_optimize_depth = 3;

    void loadFromConfigFile(const mrpt::utils::CConfigFile& config)
    {
        stereo_camera.loadFromConfigFile("CAMERA", config);
        const char* G = "SRBA_GENERAL"; const char* D = "SRBA_DETECT"; const char* A = "SRBA_DATA_ASSOCIATION"; const char* K = "SRBA_KF_CREATION";
        pause_after_show_op = config.read_bool(G, "pause_after_show_op", pause_after_show_op);
        srba_max_tree_depth = config.read_int(G, "srba_max_tree_depth", (int)srba_max_tree_depth);
        srba_max_optimize_depth = config.read_int(G, "srba_max_optimize_depth", (int)srba_max_optimize_depth);
        srba_submap_size = config.read_int(G, "srba_submap_size", (int)srba_submap_size);
        srba_use_robust_kernel = config.read_bool(G, "srba_use_robust_kernel", srba_use_robust_kernel);
        srba_use_robust_kernel_stage1 = config.read_bool(G, "srba_use_robust_kernel_stage1", srba_use_robust_kernel_stage1);
        srba_kernel_param = config.read_double(G, "srba_kernel_param", srba_kernel_param);
        n_feats = config.read_int(D, "n_feats", (int)n_feats);
        orb_adaptive_fast_th = config.read_bool(D, "orb_adaptive_fast_th", orb_adaptive_fast_th);
        detect_fast_th = config.read_int(D, "detect_fast_th", detect_fast_th);
        adaptive_th_min_matches = config.read_int(D, "adaptive_th_min_matches", adaptive_th_min_matches);
        const int st2 = config.read_int(A, "da_stage2_method", da_stage2_method);
        da_stage2_method = st2 == 1 ? ST2M_FUNDMATRIX : st2 == 2 ? ST2M_CHANGEPOSE : st2 == 3 ? ST2M_BOTH : ST2M_NONE;
        residual_th = config.read_double(A, "residual_th", residual_th);
        max_y_diff_epipolar = config.read_double(A, "max_y_diff_epipolar", max_y_diff_epipolar);
        ransac_fit_prob = config.read_double(A, "ransac_fit_prob", ransac_fit_prob);
        max_orb_distance_da = config.read_double(A, "max_orb_distance_da", max_orb_distance_da);
        query_score_th = config.read_double(A, "query_score_th", query_score_th);
        da_filter_by_direction = config.read_bool(A, "da_filter_by_direction", da_filter_by_direction);
        da_filter_by_orb_distance = config.read_bool(A, "da_filter_by_orb_distance", da_filter_by_orb_distance);
        da_filter_by_fund_matrix = config.read_bool(A, "da_filter_by_fund_matrix", da_filter_by_fund_matrix);
        da_filter_by_pose_change = config.read_bool(A, "da_filter_by_pose_change", da_filter_by_pose_change);
        max_rotation = config.read_double(K, "max_rotation", max_rotation);
        max_translation = config.read_double(K, "max_translation", max_translation);
        updated_matches_th = config.read_int(K, "updated_matches_th", (int)updated_matches_th);
        up_matches_th_plus = config.read_int(K, "up_matches_th_plus", (int)up_matches_th_plus);
        lc_distance = config.read_int(K, "lc_distance", (int)lc_distance);
        vo_id_tracking_th = config.read_int(K, "vo_id_tracking_th", vo_id_tracking_th);
        use_initial_pose = config.read_bool(K, "use_initial_pose", use_initial_pose);
    }
    void dumpToConsole() const
    {
        std::cout << " Stereo SLAM options: n_feats " << n_feats << " n_levels " << n_levels << " FAST " << detect_fast_th << " adaptive " << orb_adaptive_fast_th
                  << " (min matches " << adaptive_th_min_matches << ") | DA filters dir/orb/F/pose " << da_filter_by_direction << da_filter_by_orb_distance
                  << da_filter_by_fund_matrix << da_filter_by_pose_change << " stage2 " << da_stage2_method << " residual_th " << residual_th << " ep " << max_y_diff_epipolar
                  << " orb " << max_orb_distance_da << " ransac p " << ransac_fit_prob << " query_score_th " << query_score_th << " | KF: " << max_translation << " m "
                  << max_rotation << " deg, updated_matches_th " << updated_matches_th << "+" << up_matches_th_plus << " lc_distance " << lc_distance
                  << " vo_id_tracking_th " << vo_id_tracking_th << " submap " << srba_submap_size << std::endl;
    }
};

// KF-check limit schedules (srba-stereo-slam_utils.cpp:153-171)
inline double updateTranslationThreshold(const double x, const double th) { const double t = 0.02 + (0.25 / th) * x; return t < 0.02 ? 0.02 : (t > 0.3 ? 0.3 : t); }
inline double updateRotationThreshold(const double x, const double th) { const double t = 15 + 13 / th * (x - th); return t < 2 ? 2 : (t > 15 ? 15 : t); }
