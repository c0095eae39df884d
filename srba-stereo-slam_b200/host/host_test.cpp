// host_test.cpp -- drives the C++ adapter classes the way CSRBAStereoSLAMEstimator::performStereoSLAM does
// (EST.cpp:82-216, 220-926) on raw frames written by the test-suite, and dumps what the reference would see.
//   host_test <frames.bin> <W> <H> <nframes> <nlevels> <nfeats> <voc.yml[.gz]|-> <out.bin> [adaptive_min_matches]
// With the optional last argument the frames go through the reference's adaptive threshold loop (EST.cpp:263-315)
// and every frame's record is prefixed by {calls, fast_th, orb_th}.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>
#include "CBoWManager.h"
#include "da_filters.h"

template <typename T> static void put(std::ofstream& o, const T& v) { o.write(reinterpret_cast<const char*>(&v), sizeof(T)); }

int main(int argc, char** argv)
{
    if (argc < 9) { fprintf(stderr, "usage\n"); return 2; }
    const int W = atoi(argv[2]), H = atoi(argv[3]), NF = atoi(argv[4]);
    std::vector<uint8_t> frames((size_t)NF * 2 * W * H);
    { std::ifstream f(argv[1], std::ios::binary); f.read(reinterpret_cast<char*>(frames.data()), frames.size()); if (!f) { fprintf(stderr, "short read\n"); return 2; } }
    rso::CStereoOdometryEstimator vo;
    vo.params_detect.detect_method = rso::CStereoOdometryEstimator::TDetectParams::dmORB;                    // EST.cpp:1135
    vo.params_lr_match.match_method = rso::CStereoOdometryEstimator::TLeftRightMatchParams::smDescRbR;       // EST.cpp:1136
    vo.params_if_match.ifm_method = rso::CStereoOdometryEstimator::TInterFrameMatchingParams::ifmDescBF;     // EST.cpp:1137
    vo.params_detect.initial_FAST_threshold = 20; vo.params_detect.orb_nlevels = atoi(argv[5]); vo.params_detect.orb_nfeats = atoi(argv[6]);
    vo.setFASTThreshold(20); vo.setORBThreshold(60);
    CBoWManager bow; bool have_voc = std::string(argv[7]) != "-";
    const int adaptive = argc > 9 ? atoi(argv[9]) : 0;
    TAdaptiveThOptions aopt; aopt.adaptive_th_min_matches = adaptive; aopt.detect_fast_th = 20;
    t_vector_kf keyframes;
    std::ofstream out(argv[8], std::ios::binary);
    rso::CStereoOdometryEstimator::TStereoOdometryRequest req; rso::CStereoOdometryEstimator::TStereoOdometryResult res;
    req.stereo_cam.leftCamera.ncols = req.stereo_cam.rightCamera.ncols = W; req.stereo_cam.leftCamera.nrows = req.stereo_cam.rightCamera.nrows = H;
    req.stereo_cam.leftCamera.setIntrinsicParamsFromValues(707.0912 * W / 1241.0, 707.0912 * W / 1241.0, 601.8873 * W / 1241.0, 183.1104 * H / 376.0);      // synth.CorridorScene's pinhole
    req.stereo_cam.rightCamera = req.stereo_cam.leftCamera; req.stereo_cam.rightCameraPose[0] = 0.54;
    for (int k = 0; k < NF; ++k) {
        req.stereo_imgs = rso::CObservationStereoImagesPtr(new rso::CObservationStereoImages());       // what m_myCam.getNextFrame() yields (EST.cpp:44,79)
        req.stereo_imgs->imageLeft.loadFromMemoryBuffer(W, H, false, &frames[(size_t)(2 * k) * W * H]);
        req.stereo_imgs->imageRight.loadFromMemoryBuffer(W, H, false, &frames[(size_t)(2 * k + 1) * W * H]);
        req.repeat = false;
        if (adaptive) {
            const int calls = processWithAdaptiveThresholds(vo, req, res, aopt);
            put(out, (int32_t)calls); put(out, (int32_t)vo.getFASTThreshold()); put(out, (int32_t)vo.getORBThreshold());
        } else if (!vo.processNewImagePair(req, res)) { fprintf(stderr, "frame %d invalid: %s\n", k, vo.lastError().c_str()); }
        if (k == 0) { vo.setThisFrameAsKF(); if (have_voc) bow.loadVocabulary(vo.context(), argv[7]); }
        keyframes.push_back(CStereoSLAMKF());
        CStereoSLAMKF& kf = keyframes.back();
        kf.setKFID(k); kf.getDataFromVOEngine(vo);
        QueryResults qr;
        if (have_voc) { if (k > 0) bow.queryDB(kf, qr, 4); bow.insertIntoDB(kf); }
        std::vector<cv::DMatch> da;
        if (k > 0) matchKeyframes(vo.context(), kf, keyframes[k - 1], da);
        // dump
        put(out, (int32_t)kf.m_keypoints_left.size()); put(out, (int32_t)kf.m_keypoints_right.size()); put(out, (int32_t)kf.m_matches.size());
        put(out, (int32_t)res.tracked_feats_from_last_frame); put(out, (int32_t)qr.size()); put(out, (int32_t)da.size());
        out.write(reinterpret_cast<const char*>(kf.m_keypoints_left.data()), sizeof(cv::KeyPoint) * kf.m_keypoints_left.size());
        out.write(reinterpret_cast<const char*>(kf.m_keypoints_right.data()), sizeof(cv::KeyPoint) * kf.m_keypoints_right.size());
        out.write(reinterpret_cast<const char*>(kf.m_descriptors_left.ptr(0)), (size_t)kf.m_descriptors_left.rows * 32);
        out.write(reinterpret_cast<const char*>(kf.m_descriptors_right.ptr(0)), (size_t)kf.m_descriptors_right.rows * 32);
        out.write(reinterpret_cast<const char*>(kf.m_matches.data()), sizeof(cv::DMatch) * kf.m_matches.size());
        for (size_t id : kf.m_matches_ID) put(out, (int64_t)id);
        for (const DBoW2::Result& r : qr) { put(out, (int32_t)r.Id); put(out, r.Score); }
        out.write(reinterpret_cast<const char*>(da.data()), sizeof(cv::DMatch) * da.size());
        vo.resetIds();
    }
    std::cout << "host_test ok: " << NF << " frames" << std::endl;
    return 0;
}
