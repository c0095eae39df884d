// stereo_odometry_b200.h -- drop-in for the members of rso::CStereoOdometryEstimator (famoreno/stereo-vo,
// `libstereo-odometry.h`, not vendored by the reference) that famoreno/srba-stereo-slam touches.  Every member is
// listed with the reference call site that relies on it; the detect / describe / match work is
// done by the CUDA library behind include/srba_b200.h, there is no CPU path.  The reference's own call lines compile
// unchanged against this header (tests/conformance/ref_calls.cpp).
//
// The least-squares change in pose (outPose, getChangeInPose -- EST.cpp:327, 2139-2146) is host code as in stereo-vo:
// stereo_pose_solver.h (this repo's recipe; stereo-vo's is not available, parity unpinned).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "cv_compat.h"
#include "mrpt_lite.h"
#include "stereo_pose_solver.h"
#include "../../include/srba_b200.h"

namespace rso {

using mrpt::obs::CObservationStereoImages;
using mrpt::obs::CObservationStereoImagesPtr;
typedef std::vector<std::pair<size_t, size_t>> vector_index_pairs_t;          // EST.cpp:2120

class CStereoOdometryEstimator {
public:
    // ---- parameter blocks (enumerations as documented in the reference's config comments, example.ini:9,33,51) ----
    struct TDetectParams {
        enum TDetectMethod { dmORB = 0, dmFAST_ORB, dmFASTER, dmKLT };
        TDetectMethod detect_method = dmORB;          // forced at EST.cpp:1135
        int initial_FAST_threshold = 20;              // EST.cpp:1140
        int fast_min_th = 5, fast_max_th = 30;        // example.ini:15-16
        int orb_nlevels = 8;                          // EST.cpp:1141 (reference default: 1, UTL.h:288)
        int orb_nfeats = 500;                         // EST.cpp:1142
        // [DETECT] post-filters (kitti.ini:9-13).  Parsed and reported; NOT applied: under dmORB the keypoints are
        // cv::ORB's own (already 3x3-NMS'd FAST corners re-ranked by Harris), the 5x5 post-NMS / minimum-distance stage
        // belongs to stereo-vo's FAST paths as far as its config comments say -- a documented parity risk.
        int min_distance = 3; bool non_maximal_suppression = true; int non_max_supp_method = 0;
    };
    struct TLeftRightMatchParams {
        enum TLRMatchMethod { smDescBF = 0, smDescRbR, smSAD };
        TLRMatchMethod match_method = smDescRbR;      // forced at EST.cpp:1136
        double max_y_diff = 2.0;                      // kitti.ini:17
        bool enable_robust_1to1_match = false;        // kitti.ini:18
        bool rectified_images = true;                 // kitti.ini:19
        double orb_max_distance = 60;                 // kitti.ini:24
        int orb_min_th = 30, orb_max_th = 100;        // example.ini:45-46
        double min_disparity = 0.0, max_disparity = 1e9;
    };
    struct TInterFrameMatchingParams {
        enum TIFMMethod { ifmDescBF = 0, ifmDescWin, ifmSAD, ifmOpticalFlow };
        TIFMMethod ifm_method = ifmDescBF;            // forced at EST.cpp:1137
        double orb_max_distance = 60;                 // "the same values as for stereo-matching are employed", kitti.ini:38
    };
    struct TLeastSquaresParams : public TStereoPoseSolverParams {             // [LEAST_SQUARES], kitti.ini:40-55
        bool use_custom_initial_pose = false;         // toggled at EST.cpp:2138,2147
        bool use_previous_pose_as_initial = true;     // kitti.ini:55
    };

    struct TStereoOdometryRequest {
        CObservationStereoImagesPtr stereo_imgs;      // EST.cpp:111,258
        bool use_precomputed_data = false;            // EST.cpp:259
        bool repeat = false;                          // EST.cpp:260,278,285,315
        mrpt::utils::TStereoCamera stereo_cam;        // EST.cpp:1167
    };
    struct TStereoOdometryResult {
        mrpt::poses::CPose3DRotVec outPose;           // pose of the current frame wrt the previous one, EST.cpp:327
        bool valid = false;                           // EST.cpp:318
        std::vector<std::pair<size_t, size_t>> detected_feats;   // [octave] (left, right), EST.cpp:268
        std::vector<size_t> stereo_matches;           // [octave], EST.cpp:269,273,288,297
        size_t tracked_feats_from_last_frame = 0;     // EST.cpp:359
        size_t tracked_feats_from_last_KF = 0;        // EST.cpp:360,378
        int num_it = 0, num_it_final = 0;             // EST.cpp:2151
        std::vector<double> out_residual;             // EST.cpp:2172 (one per tracked pair, px^2)
    };

    TDetectParams params_detect;
    TLeftRightMatchParams params_lr_match;
    TInterFrameMatchingParams params_if_match;
    TLeastSquaresParams params_least_squares;

    CStereoOdometryEstimator();
    ~CStereoOdometryEstimator();
    CStereoOdometryEstimator(const CStereoOdometryEstimator&) = delete;
    CStereoOdometryEstimator& operator=(const CStereoOdometryEstimator&) = delete;

    // EST.cpp:1122-1131: sections RECTIFY, DETECT, MATCH, IF-MATCH, LEAST_SQUARES, GUI, GENERAL.
    void loadParamsFromConfigFile(const mrpt::utils::CConfigFile& config, const std::vector<std::string>& sections);
    void loadParamsFromConfigFile(const std::string& ini_file, const std::vector<std::string>& sections);   // convenience
    void setVerbosityLevel(int v) { m_verbose = v; }                       // EST.cpp:1132
    void dumpToConsole() const;                                           // EST.cpp:1144

    // EST.cpp:112,267.  Creates the GPU context on first use from params_* and the image size.  Returns result.valid.
    bool processNewImagePair(TStereoOdometryRequest& request, TStereoOdometryResult& result);

    void setThisFrameAsKF();                                              // EST.cpp:113
    void resetIds();                                                      // EST.cpp:403,620

    int  getFASTThreshold() const { return m_fast_th; }                   // EST.cpp:275
    void setFASTThreshold(int th);                                        // EST.cpp:277,302,310
    bool isFASTThMin() const { return m_fast_th <= params_detect.fast_min_th; }   // EST.cpp:275,301
    int  getORBThreshold() const { return m_orb_th; }                     // EST.cpp:281
    void setORBThreshold(int th);                                         // EST.cpp:284,304,311
    bool isORBThMax() const { return m_orb_th >= params_lr_match.orb_max_th; }    // EST.cpp:281
    void resetORBThreshold() { setORBThreshold((int)params_lr_match.orb_max_distance); }   // EST.cpp:303

    // CStereoSLAMKF.h:68-70: argument order and types of the reference.
    void getValues(std::vector<cv::KeyPoint>& kpL, std::vector<cv::KeyPoint>& kpR, cv::Mat& descL, cv::Mat& descR,
                   std::vector<cv::DMatch>& matches, std::vector<size_t>& matches_ID);

    // EST.cpp:2139-2146.  tracked_pairs[i] = (index into pre_matches, index into cur_matches) (EST.cpp:2127); `initial`
    // = [w1,w2,w3,tx,ty,tz], pose of the previous frame wrt the current one, used when
    // params_least_squares.use_custom_initial_pose is set.  Fills result.{valid, outPose, num_it, num_it_final,
    // out_residual (one entry per tracked pair, same order)}.  Returns result.valid.
    bool getChangeInPose(const vector_index_pairs_t& tracked_pairs, const std::vector<cv::DMatch>& pre_matches,
                         const std::vector<cv::DMatch>& cur_matches, const std::vector<cv::KeyPoint>& pre_left,
                         const std::vector<cv::KeyPoint>& pre_right, const std::vector<cv::KeyPoint>& cur_left,
                         const std::vector<cv::KeyPoint>& cur_right, const mrpt::utils::TStereoCamera& cam, TStereoOdometryResult& result,
                         const std::vector<double>& initial = std::vector<double>());

    sfe_context* context() { return m_ctx; }          // for CBoWManager / keyframe matching on the same GPU
    // The GPU context of the most recently constructed engine that has one (CBoWManager picks it up lazily: the reference
    // loads the vocabulary before the engine has seen an image, EST.cpp:1102 vs :112).
    static sfe_context* currentContext();
    // Creates the GPU context ahead of the first frame (image size from the stereo camera block).
    bool prepare(int width, int height) { return ensureContext(width, height); }
    const std::string& lastError() const { return m_err; }
    int device = 0;                                   // CUDA device ordinal of the context this engine creates

private:
    bool ensureContext(int width, int height);
    sfe_context* m_ctx = nullptr;
    int m_w = 0, m_h = 0, m_fast_th = 20, m_orb_th = 60, m_verbose = 0;
    std::string m_err;
    // current frame (host copies of the GPU results)
    std::vector<cv::KeyPoint> m_kpL, m_kpR; cv::Mat m_descL, m_descR; std::vector<cv::DMatch> m_matches; std::vector<size_t> m_ids;
    // previous frame (the last NON-repeat call's predecessor), for ifmDescBF id propagation and the pose solver; the
    // current frame is only retired to "previous" when the caller moves on to a new frame
    cv::Mat m_prev_matched_desc, m_cur_matched_desc; std::vector<size_t> m_prev_ids;
    std::vector<cv::KeyPoint> m_prev_kpL, m_prev_kpR; std::vector<cv::DMatch> m_prev_matches;
    size_t m_next_id_at_frame_start = 0;
    std::vector<size_t> m_kf_ids;                     // ids alive at the last keyframe / resetIds()
    size_t m_next_id = 0; bool m_have_frame = false, m_thresholds_touched = false;
    double m_last_pose_guess[6] = { 0, 0, 0, 0, 0, 0 };   // last solution (previous wrt current), use_previous_pose_as_initial
    double m_last_solver_pose[6] = { 0, 0, 0, 0, 0, 0 };         // raw output of the last getChangeInPose call
    double m_pose_guess_frame_start[6] = { 0, 0, 0, 0, 0, 0 };   // ... as it stood when the current frame first arrived (repeat re-uses it)
};

}  // namespace rso
