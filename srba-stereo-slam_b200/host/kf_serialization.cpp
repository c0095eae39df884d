#include "kf_serialization.h"
#include <cstring>

template <typename T> static void wr(std::ostream& s, const T& v) { s.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T> static void rd(std::istream& s, T& v) { s.read(reinterpret_cast<char*>(&v), sizeof(T)); }

bool dumpKeypointsToStream(std::ostream& stream, const TKeyPointList& keypoints, const cv::Mat& descriptors)
{
    if (!stream.good()) return false;
    const size_t num_kp = keypoints.size();
    wr(stream, num_kp);
    for (size_t f = 0; f < num_kp; ++f) {
        wr(stream, keypoints[f].pt.x); wr(stream, keypoints[f].pt.y); wr(stream, keypoints[f].response);
        wr(stream, keypoints[f].size); wr(stream, keypoints[f].angle); wr(stream, keypoints[f].octave); wr(stream, keypoints[f].class_id);
    }
    const int drows = descriptors.rows, dcols = descriptors.cols, dtype = descriptors.type();
    wr(stream, drows); wr(stream, dcols); wr(stream, dtype);
    // the reference walks the elements with MatIterator_<uchar> (EST.cpp:2474-2480): one byte per element, row by row,
    // so a non-continuous Mat (row range / ROI view under real OpenCV) must not be written from ptr(0) in one piece
    if (dtype != CV_8U) return false;
    for (int r = 0; r < drows && dcols > 0; ++r) stream.write(reinterpret_cast<const char*>(descriptors.ptr(r)), (std::streamsize)dcols);
    return stream.good();
}

bool loadKeypointsFromStream(std::istream& stream, TKeyPointList& keypoints, cv::Mat& descriptors)
{
    if (!stream.good()) return false;
    size_t num_kp = 0;
    rd(stream, num_kp);
    if (!stream.good() || num_kp > (1u << 24)) return false;
    keypoints.resize(num_kp);
    for (size_t f = 0; f < num_kp; ++f) {
        rd(stream, keypoints[f].pt.x); rd(stream, keypoints[f].pt.y); rd(stream, keypoints[f].response);
        rd(stream, keypoints[f].size); rd(stream, keypoints[f].angle); rd(stream, keypoints[f].octave); rd(stream, keypoints[f].class_id);
    }
    int drows = 0, dcols = 0, dtype = 0;
    rd(stream, drows); rd(stream, dcols); rd(stream, dtype);
    if (!stream.good() || drows < 0 || dcols < 0 || (size_t)drows * dcols > (1u << 30)) return false;
    if (dtype != CV_8U) return false;
    descriptors.create(drows, dcols, dtype);
    for (int r = 0; r < drows && dcols > 0; ++r) stream.read(reinterpret_cast<char*>(descriptors.ptr(r)), (std::streamsize)dcols);
    return stream.good();
}

bool dumpMatchesToStream(std::ostream& stream, const TDMatchList& matches, const std::vector<size_t>& ids)
{
    if (!stream.good() || ids.size() < matches.size()) return false;
    const size_t num_m = matches.size();
    wr(stream, num_m);
    for (size_t m = 0; m < num_m; ++m) {
        wr(stream, ids[m]); wr(stream, matches[m].queryIdx); wr(stream, matches[m].trainIdx); wr(stream, matches[m].distance); wr(stream, matches[m].imgIdx);
    }
    return stream.good();
}

bool loadMatchesFromStream(std::istream& stream, TDMatchList& matches, std::vector<size_t>& ids)
{
    if (!stream.good()) return false;
    size_t num_m = 0;
    rd(stream, num_m);
    if (!stream.good() || num_m > (1u << 24)) return false;
    matches.resize(num_m); ids.resize(num_m);
    for (size_t m = 0; m < num_m; ++m) {
        rd(stream, ids[m]); rd(stream, matches[m].queryIdx); rd(stream, matches[m].trainIdx); rd(stream, matches[m].distance); rd(stream, matches[m].imgIdx);
    }
    return stream.good();
}

bool dumpKeyframeToStream(std::ostream& s, const CStereoSLAMKF& kf)
{
    return dumpKeypointsToStream(s, kf.m_keypoints_left, kf.m_descriptors_left) && dumpKeypointsToStream(s, kf.m_keypoints_right, kf.m_descriptors_right) &&
           dumpMatchesToStream(s, kf.m_matches, kf.m_matches_ID);
}
bool loadKeyframeFromStream(std::istream& s, CStereoSLAMKF& kf)
{
    return loadKeypointsFromStream(s, kf.m_keypoints_left, kf.m_descriptors_left) && loadKeypointsFromStream(s, kf.m_keypoints_right, kf.m_descriptors_right) &&
           loadMatchesFromStream(s, kf.m_matches, kf.m_matches_ID);
}

// ---- C shims for the CPU test-suite: round trip through a file, from / to plain arrays --------------------------------
extern "C" {
int srba_kf_dump_file(const char* path, const void* kpl, int nl, const void* kpr, int nr, const unsigned char* dl, const unsigned char* dr,
                      const void* matches, int nm, const unsigned long long* ids)
{
    CStereoSLAMKF kf;
    kf.m_keypoints_left.resize(nl); kf.m_keypoints_right.resize(nr); kf.m_matches.resize(nm); kf.m_matches_ID.assign(ids, ids + nm);
    if (nl) std::memcpy((void*)kf.m_keypoints_left.data(), kpl, sizeof(cv::KeyPoint) * nl);
    if (nr) std::memcpy((void*)kf.m_keypoints_right.data(), kpr, sizeof(cv::KeyPoint) * nr);
    if (nm) std::memcpy((void*)kf.m_matches.data(), matches, sizeof(cv::DMatch) * nm);
    kf.m_descriptors_left.create(nl, 32, CV_8U); kf.m_descriptors_right.create(nr, 32, CV_8U);
    if (nl) std::memcpy(kf.m_descriptors_left.ptr(0), dl, (size_t)nl * 32);
    if (nr) std::memcpy(kf.m_descriptors_right.ptr(0), dr, (size_t)nr * 32);
    std::ofstream f(path, std::ios::binary);
    return dumpKeyframeToStream(f, kf) ? 0 : 1;
}
int srba_kf_load_file(const char* path, void* kpl, int* nl, void* kpr, int* nr, unsigned char* dl, unsigned char* dr, void* matches, int* nm,
                      unsigned long long* ids, int cap)
{
    std::ifstream f(path, std::ios::binary);
    CStereoSLAMKF kf;
    if (!loadKeyframeFromStream(f, kf)) return 1;
    *nl = (int)kf.m_keypoints_left.size(); *nr = (int)kf.m_keypoints_right.size(); *nm = (int)kf.m_matches.size();
    if (*nl > cap || *nr > cap || *nm > cap) return 2;
    if (*nl) { std::memcpy(kpl, kf.m_keypoints_left.data(), sizeof(cv::KeyPoint) * *nl); std::memcpy(dl, kf.m_descriptors_left.ptr(0), (size_t)*nl * 32); }
    if (*nr) { std::memcpy(kpr, kf.m_keypoints_right.data(), sizeof(cv::KeyPoint) * *nr); std::memcpy(dr, kf.m_descriptors_right.ptr(0), (size_t)*nr * 32); }
    if (*nm) { std::memcpy(matches, kf.m_matches.data(), sizeof(cv::DMatch) * *nm); for (int i = 0; i < *nm; ++i) ids[i] = kf.m_matches_ID[i]; }
    return 0;
}
}

// ---- options block -----------------------------------------------------------------------------------------------------
#define SRBA_OPT_FIELDS(X, o) \
    X(o.n_levels) X(o.n_feats) X(o.min_ORB_distance) X(o.detect_method) X(o.detect_fast_th) X(o.non_max_supp_method) \
    X(o.matching_options.epipolar_TH) X(o.max_y_diff_epipolar) X(o.matching_options.maxORB_dist) X(o.max_orb_distance_da) \
    X(o.ransac_fit_prob) X(o.matching_options.max_disp) X(o.matching_options.min_disp) X(o.matching_options.enable_robust_1to1_match) \
    X(o.residual_th) X(o.max_translation) X(o.max_rotation) X(o.non_maximal_suppression) X(o.updated_matches_th) X(o.up_matches_th_plus)
bool dumpOptionsToStream(std::ostream& stream, const TSRBAStereoSLAMOptions& options)
{
    if (!stream.good()) return false;
#define X(f) wr(stream, f);
    SRBA_OPT_FIELDS(X, options)
#undef X
    return stream.good();
}
bool loadOptionsFromStream(std::istream& stream, TSRBAStereoSLAMOptions& options)
{
    if (!stream.good()) return false;
#define X(f) rd(stream, f);
    SRBA_OPT_FIELDS(X, options)
#undef X
    return stream.good();
}
