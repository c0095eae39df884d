// CBoWManager.h -- the reference's bag-of-words manager (src/CBoWManager.h:44-111) on the GPU.
// Same class name and method signatures.  DBoW2's BriefVocabulary + BriefDatabase are replaced by the sfe_bow_*
// entry points of include/srba_b200.h; the 256-iteration-per-descriptor bit repack (m_change_structure_binary,
// CBoWManager.h:95-109) disappears because the vocabulary is stored in OpenCV byte order on the device.
#pragma once
#include <iostream>
#include <string>
#include <vector>
#include "CStereoSLAMKF.h"

namespace DBoW2 {
typedef unsigned int EntryId;
struct Result {                                   // DBoW2::Result: {Id, Score}, streamable (EST.cpp:1746)
    EntryId Id; double Score;
    Result() : Id(0), Score(0) {}
    Result(EntryId id, double s) : Id(id), Score(s) {}
};
inline std::ostream& operator<<(std::ostream& os, const Result& r) { return os << "<EntryId: " << r.Id << ", Score: " << r.Score << ">"; }
class QueryResults : public std::vector<Result> {};
inline std::ostream& operator<<(std::ostream& os, const QueryResults& q)
{
    os << q.size() << " results:" << std::endl;
    for (const Result& r : q) os << r << std::endl;
    return os;
}
}  // namespace DBoW2
using DBoW2::QueryResults;

// Flattened vocabulary (nodes indexed by DBoW2 node id, children in file order, OpenCV byte order).
struct TFlatVocabulary {
    int k = 0, L = 0, scoring = 0, weighting = 0, nnodes = 0, nwords = 0;
    std::vector<int32_t> child_ofs, child_idx, node_word;
    std::vector<uint8_t> node_desc;
    std::vector<double> node_weight;
};
// Parses a DBoW2 1.0.2 `voc.yml.gz` (or plain .yml).  Throws std::runtime_error on failure.
void loadFlatVocabulary(const std::string& filename, TFlatVocabulary& out);

class CBoWManager {
public:
    CBoWManager() : m_ctx(nullptr), m_bow(nullptr) {}
    ~CBoWManager();

    // CBoWManager.h:59-66: reads the file name from [section] param of the ini (relative names are also tried next to the
    // ini file) and loads it; throws if the file does not exist (ASSERT_FILE_EXISTS_, :63).  The reference calls this
    // BEFORE the engine has seen an image (EST.cpp:1102 vs :112), i.e. before a GPU context exists: the vocabulary is
    // parsed to host memory here and uploaded on the first insertIntoDB / queryDB, to the context given by setContext()
    // or else to rso::CStereoOdometryEstimator::currentContext().
    void loadVocabularyFromConfigFile(const mrpt::utils::CConfigFile& config, const std::string& section, const std::string& param);
    void loadVocabulary(const std::string& voc_filename);
    void setContext(sfe_context* ctx) { m_ctx = ctx; }
    // round-1 forms (context given up front, uploaded immediately)
    void loadVocabulary(sfe_context* ctx, const std::string& voc_filename);
    void loadVocabularyFromConfigFile(sfe_context* ctx, const std::string& ini_file, const std::string& section, const std::string& param);

    void insertIntoDB(const CStereoSLAMKF& kf);                                                 // CBoWManager.h:71-76
    void queryDB(const CStereoSLAMKF& kf, QueryResults& ret, unsigned int num_results = 1);     // CBoWManager.h:83-88
    int size() const;

private:
    void ensureUploaded();
    sfe_context* m_ctx;
    sfe_bow* m_bow;
    std::shared_ptr<TFlatVocabulary> m_host_voc;      // parsed, waiting for a context
};

// The brute-force part of CSRBAStereoSLAMEstimator::m_internal_data_association (EST.cpp:1432-1452): gathers the
// left descriptors of the stereo-matched features of both keyframes and runs BFMatcher(NORM_HAMMING,false).match
// on the GPU.  out[k].queryIdx == k indexes this_kf.m_matches, trainIdx indexes other_kf.m_matches.
void matchKeyframes(sfe_context* ctx, const CStereoSLAMKF& this_kf, const CStereoSLAMKF& other_kf, std::vector<cv::DMatch>& out);
