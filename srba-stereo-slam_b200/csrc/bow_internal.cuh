// bow_internal.cuh -- vocabulary / database state shared by k_bow.cu (single-keyframe API) and k_kfdb.cu (batched path).
#pragma once
#include "sfe_internal.cuh"

struct sfe_bow {
    sfe_context* ctx = nullptr;
    int nnodes = 0, nwords = 0, depth = 0;
    // BFS-renumbered tree: children of a node are contiguous
    int32_t* d_first = nullptr; int32_t* d_nchild = nullptr; uint4* d_desc = nullptr;
    int32_t* d_word = nullptr; double* d_weight = nullptr;
    // scratch for one keyframe
    int cap_n = 0;
    uint8_t* d_in = nullptr; int32_t* d_w = nullptr; double* d_wt = nullptr;
    int32_t* d_vw = nullptr; double* d_vv = nullptr; int32_t* d_vn = nullptr;
    // database (CSR, append-only)
    std::vector<int32_t> entry_ofs{ 0 };
    int32_t* d_entry_ofs = nullptr; int ofs_cap = 0;
    int32_t* d_db_w = nullptr; double* d_db_v = nullptr; size_t db_cap = 0;
    double* d_qdense = nullptr; double* d_score = nullptr; int32_t* d_common = nullptr; int score_cap = 0;
    int32_t* d_top_id = nullptr; double* d_top_s = nullptr; int32_t* d_top_n = nullptr;
};

