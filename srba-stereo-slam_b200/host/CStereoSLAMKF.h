// CStereoSLAMKF.h -- the reference's keyframe container (src/CStereoSLAMKF.h:56-106) over the B200 front-end.
// Same class name, members and method signatures; the only change is the engine type behind getDataFromVOEngine,
// which is the CUDA-backed rso::CStereoOdometryEstimator of stereo_odometry_b200.h.
#pragma once
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>
#include "stereo_odometry_b200.h"
#include "srba_options.h"

typedef std::vector<cv::KeyPoint> TKeyPointList;
typedef std::vector<cv::DMatch> TDMatchList;

class CStereoSLAMKF;
typedef std::vector<CStereoSLAMKF> t_vector_kf;

// The reference builds its debug file names with a literal backslash ("%s\\info_kf%04d.txt", srba-stereo-slam_common.h:
// 81-84, CStereoSLAMKF.cpp:69); on Linux that yields files called `out\info_kf0000.txt` in the working directory.  The
// default here is '/', define SRBA_B200_REFERENCE_PATH_SEPARATOR to get the reference's names byte for byte.
#ifdef SRBA_B200_REFERENCE_PATH_SEPARATOR
#define SRBA_PATH_SEP "\\"
#else
#define SRBA_PATH_SEP "/"
#endif
inline void srba_prepare_out_dir()
{
    if (!general_options.out_dir.empty()) mkdir(general_options.out_dir.c_str(), 0777);        // CStereoSLAMKF.cpp:63-64
}

class CStereoSLAMKF {
public:
    CStereoSLAMKF() : m_kf_ID(0) {}

    inline void setKFID(const size_t ID) { m_kf_ID = ID; }                                     // KF.h:63-65

    // KF.h:68-70: pulls the last keypoints, descriptors, L/R matches and ids out of the engine.
    inline void getDataFromVOEngine(rso::CStereoOdometryEstimator& voEngine)
    {
        voEngine.getValues(m_keypoints_left, m_keypoints_right, m_descriptors_left, m_descriptors_right, m_matches, m_matches_ID);
    }

    inline void generateMatchesIDs(const size_t _starting_idx)                                  // KF.h:75-80
    {
        size_t starting_idx = _starting_idx;
        for (size_t m = 0; m < m_matches.size(); ++m) m_matches_ID.push_back(starting_idx++);
    }

    void dumpToConsole() const                                                                  // KF.cpp:30-55
    {
        std::cout << "KEYFRAME [" << m_kf_ID << "]\n   :: Camera pose = " << m_camera_pose << "\n   :: Matches [" << m_matches.size() << " out of "
                  << m_keypoints_left.size() << "/" << m_keypoints_right.size() << "]: ID: left_kp_x,left_kp_y --> right_kp_x,right_kp_y" << std::endl;
        for (size_t k = 0; k < m_matches.size(); ++k) {
            const cv::KeyPoint& l = m_keypoints_left[m_matches[k].queryIdx]; const cv::KeyPoint& r = m_keypoints_right[m_matches[k].trainIdx];
            std::cout << "   " << (k < m_matches_ID.size() ? m_matches_ID[k] : 0) << ": " << l.pt.x << "," << l.pt.y << " --> " << r.pt.x << "," << r.pt.y << std::endl;
        }
    }

    // KF.cpp:60-110.  <out_dir>/[<modif>_]info_kf%04d.txt: header + one `kf_id match_id lx ly rx ry distance` line per match
    // (match id 0 when the ids are not set); <out_dir>/[<modif>_]info_feats_kf%04d.txt: `nL nR`, then the left and the
    // right keypoint coordinates, all %.2f.
    void saveInfoToFiles(const std::string& str_modif = std::string()) const
    {
        srba_prepare_out_dir();
        const std::string pre = general_options.out_dir + SRBA_PATH_SEP + (str_modif.empty() ? std::string() : str_modif + "_");
        std::string name = pre + mrpt::format("info_kf%04d.txt", (int)m_kf_ID);
        FILE* f = fopen(name.c_str(), "wt");
        if (!f) throw std::runtime_error("Output file " + name + " could not be opened");
        fprintf(f, "%% [KF_ID] [MATCH_ID] [LEFT_PT{x y}] [RIGHT_PT{x y}] [MATCH_DISTANCE]\n");
        for (size_t m = 0; m < m_matches.size(); ++m) {
            const cv::KeyPoint& l = m_keypoints_left[m_matches[m].queryIdx]; const cv::KeyPoint& r = m_keypoints_right[m_matches[m].trainIdx];
            fprintf(f, "%d %d %.2f %.2f %.2f %.2f %.2f\n", (int)m_kf_ID, m_matches_ID.size() > 0 ? (int)m_matches_ID[m] : 0,
                    l.pt.x, l.pt.y, r.pt.x, r.pt.y, m_matches[m].distance);
        }
        fclose(f);
        name = pre + mrpt::format("info_feats_kf%04d.txt", (int)m_kf_ID);
        f = fopen(name.c_str(), "wt");
        if (!f) throw std::runtime_error("Output file " + name + " could not be opened");
        fprintf(f, "%d %d\n", (int)m_keypoints_left.size(), (int)m_keypoints_right.size());
        for (const cv::KeyPoint& k : m_keypoints_left) fprintf(f, "%.2f %.2f\n", k.pt.x, k.pt.y);
        for (const cv::KeyPoint& k : m_keypoints_right) fprintf(f, "%.2f %.2f\n", k.pt.x, k.pt.y);
        fclose(f);
    }

    // ---- data members (KF.h:99-104) ----
    TKeyPointList m_keypoints_left, m_keypoints_right;
    cv::Mat m_descriptors_left, m_descriptors_right;
    TDMatchList m_matches;
    std::vector<size_t> m_matches_ID;
    mrpt::poses::CPose3DRotVec m_camera_pose;
    size_t m_kf_ID;
};
