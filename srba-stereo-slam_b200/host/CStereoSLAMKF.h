// CStereoSLAMKF.h -- the reference's keyframe container (src/CStereoSLAMKF.h:56-106) over the B200 front-end.
// Same class name, members and method signatures; the only change is the engine type behind getDataFromVOEngine,
// which is the CUDA-backed rso::CStereoOdometryEstimator of stereo_odometry_b200.h.  Pose members that need MRPT
// (CPose3DRotVec m_camera_pose) are carried as a plain 6-vector when MRPT is absent.
#pragma once
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>
#include "stereo_odometry_b200.h"

typedef std::vector<cv::KeyPoint> TKeyPointList;
typedef std::vector<cv::DMatch> TDMatchList;

class CStereoSLAMKF;
typedef std::vector<CStereoSLAMKF> t_vector_kf;

class CStereoSLAMKF {
public:
    CStereoSLAMKF() : m_kf_ID(0) { for (double& v : m_camera_pose) v = 0; }

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
        std::cout << "KF #" << m_kf_ID << ": " << m_keypoints_left.size() << " left / " << m_keypoints_right.size()
                  << " right keypoints, " << m_matches.size() << " matches" << std::endl;
    }

    // KF.cpp:60-110: info_kf%04d.txt = per match `kf_id match_id lx ly rx ry distance`; info_feats_kf%04d.txt = nL nR + coordinates.
    void saveInfoToFiles(const std::string& str_modif = std::string()) const
    {
        char name[256];
        snprintf(name, sizeof(name), "info_kf%04d%s.txt", (int)m_kf_ID, str_modif.c_str());
        if (FILE* f = fopen(name, "wt")) {
            fprintf(f, "%% KF_ID MATCH_ID LX LY RX RY DISTANCE\n");
            for (size_t m = 0; m < m_matches.size(); ++m) {
                const cv::KeyPoint& l = m_keypoints_left[m_matches[m].queryIdx]; const cv::KeyPoint& r = m_keypoints_right[m_matches[m].trainIdx];
                fprintf(f, "%d %d %.2f %.2f %.2f %.2f %.2f\n", (int)m_kf_ID, m < m_matches_ID.size() ? (int)m_matches_ID[m] : -1,
                        l.pt.x, l.pt.y, r.pt.x, r.pt.y, m_matches[m].distance);
            }
            fclose(f);
        }
        snprintf(name, sizeof(name), "info_feats_kf%04d%s.txt", (int)m_kf_ID, str_modif.c_str());
        if (FILE* f = fopen(name, "wt")) {
            fprintf(f, "%d %d\n", (int)m_keypoints_left.size(), (int)m_keypoints_right.size());
            for (const cv::KeyPoint& k : m_keypoints_left) fprintf(f, "%.2f %.2f\n", k.pt.x, k.pt.y);
            for (const cv::KeyPoint& k : m_keypoints_right) fprintf(f, "%.2f %.2f\n", k.pt.x, k.pt.y);
            fclose(f);
        }
    }

    // ---- data members (KF.h:99-104) ----
    TKeyPointList m_keypoints_left, m_keypoints_right;
    cv::Mat m_descriptors_left, m_descriptors_right;
    TDMatchList m_matches;
    std::vector<size_t> m_matches_ID;
    double m_camera_pose[6];          // CPose3DRotVec when built with MRPT
    size_t m_kf_ID;
};
