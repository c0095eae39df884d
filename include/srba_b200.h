/*
 * srba_b200.h -- C ABI of the B200-native stereo front-end for SRBA-Stereo-SLAM.
 *
 * This is the drop-in boundary.  The reference (famoreno/srba-stereo-slam) has
 * no FFI of its own: its hot path is reached through three C++ classes and the external
 * `rso::CStereoOdometryEstimator`.  Each entry point below names the reference call site
 * (file:line relative to /root/reference/src) whose work it replaces; the C++ adapter classes
 * in srba-stereo-slam_b200/host/ (same names and signatures as the reference's) call only these.
 *
 * Conventions: plain pointers and sizes, caller-owned output buffers, integer status codes
 * (no exceptions cross the ABI), one context per GPU, one caller thread per context.
 * There is NO CPU fallback: every entry point fails with SFE_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef SRBA_B200_H
#define SRBA_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------ */
enum {
    SFE_OK = 0,
    SFE_ERR_INVALID = 1,   /* bad argument / unsupported configuration                         */
    SFE_ERR_CUDA = 2,      /* CUDA runtime/driver failure or no usable device                  */
    SFE_ERR_CAPACITY = 3,  /* a device-side list overflowed its capacity (result truncated)    */
    SFE_ERR_STATE = 4      /* call order violation (e.g. get_values before process)            */
};

/* ---- plain-old-data mirrors of the OpenCV structs the reference stores --------------------- */
/* cv::KeyPoint: field list srba-stereo-slam_utils.cpp:49-50, serialised order EST.cpp:2540-2546 */
typedef struct sfe_keypoint {
    float x, y;        /* pt, level-0 pixel coordinates                                        */
    float size;        /* 31 * 1.2^octave                                                      */
    float angle;       /* degrees, [0,360)                                                     */
    float response;    /* Harris response                                                      */
    int32_t octave;    /* pyramid level                                                        */
    int32_t class_id;  /* -1                                                                   */
} sfe_keypoint;        /* 28 bytes */

/* cv::DMatch: field list srba-stereo-slam_utils.cpp:75, serialised order EST.cpp:2609-2612 */
typedef struct sfe_dmatch {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} sfe_dmatch;          /* 16 bytes */

/* ---- configuration ------------------------------------------------------------------------- */
enum { SFE_BLUR_CV24_FIXED = 0,   /* OpenCV-2.4 8-bit fixed-point Gaussian (reference parity)     */
       SFE_BLUR_FLOAT = 1 };      /* float sepFilter2D arithmetic (OpenCV >= 3.4; cv2-pinned)     */

enum { SFE_PYRAMID_CV24_LINEAR = 0,  /* OpenCV-2.4 resize(INTER_LINEAR), 11-bit coefficients (reference parity)   */
       SFE_PYRAMID_CV4_EXACT = 1 };  /* INTER_LINEAR_EXACT + cv>=3 level sizes; with SFE_BLUR_FLOAT the whole
                                        extraction is bit-identical to cv2 4.13's ORB (pinned by tests/golden)    */

typedef struct sfe_config {
    int32_t width, height;      /* level-0 image size ([CAMERA_LEFT] ncols/nrows, kitti.ini:139,147) */
    int32_t nlevels;            /* params_detect.orb_nlevels, EST.cpp:1141 (1..8)                  */
    int32_t nfeatures;          /* params_detect.orb_nfeats,  EST.cpp:1142                         */
    float   scale_factor;       /* OpenCV ORB default 1.2                                          */
    int32_t edge_threshold;     /* OpenCV ORB default 31                                           */
    int32_t fast_threshold;     /* params_detect.initial_FAST_threshold, EST.cpp:1140              */
    int32_t blur_mode;          /* SFE_BLUR_*                                                      */
    int32_t max_pairs;          /* stereo pairs per batch the context is sized for                 */
    float   max_y_diff;         /* [MATCH] max_y_diff, kitti.ini:17                                */
    float   min_disparity;      /* row-band match: accept xl-xr in [min,max]                       */
    float   max_disparity;
    int32_t orb_max_distance;   /* [MATCH] orb_max_distance, kitti.ini:24                          */
    int32_t robust_1to1;        /* [MATCH] enable_robust_1to1_match, kitti.ini:18                  */
    int32_t device;             /* CUDA device ordinal                                             */
    int32_t use_graph;          /* 1: replay the batch as a CUDA graph                             */
    int32_t pyramid_mode;       /* SFE_PYRAMID_*                                                   */
} sfe_config;

void sfe_default_config(sfe_config* cfg);   /* kitti.ini values, 1241x376, 8 levels, 1000 feats */

typedef struct sfe_context sfe_context;

/* ---- lifetime --------------------------------------------------------------------------------
 * Replaces: CStereoOdometryEstimator construction + loadParamsFromConfigFile (EST.cpp:1122-1142). */
int  sfe_create(const sfe_config* cfg, sfe_context** out);
void sfe_destroy(sfe_context* ctx);
const char* sfe_last_error(const sfe_context* ctx);   /* ctx may be NULL: last create() error      */
const char* sfe_version(void);
void* sfe_stream(sfe_context* ctx);                   /* the cudaStream_t all work is enqueued on  */

/* ---- run-time thresholds -------------------------------------------------------------------
 * Replaces: get/setFASTThreshold, get/setORBThreshold (EST.cpp:275-277,281-284,301-304,310-311). */
int sfe_set_fast_threshold(sfe_context* ctx, int th);
int sfe_get_fast_threshold(const sfe_context* ctx);
int sfe_set_orb_threshold(sfe_context* ctx, int th);
int sfe_get_orb_threshold(const sfe_context* ctx);

/* ---- geometry queries ------------------------------------------------------------------------ */
int sfe_num_levels(const sfe_context* ctx);
int sfe_level_size(const sfe_context* ctx, int level, int32_t* w, int32_t* h, int32_t* pitch, int32_t* quota);
int sfe_keypoint_capacity(const sfe_context* ctx);    /* per image */

/* ---- the per-frame stage ---------------------------------------------------------------------
 * Replaces: m_voEngine.processNewImagePair(request, result) detect+describe+L/R match
 * (EST.cpp:112, 267; forced modes dmORB / smDescRbR, EST.cpp:1135-1136).
 *
 * images: npairs * 2 host images (pair p: left at index 2p, right at 2p+1), 8-bit grey,
 * `stride` bytes per row, `image_stride` bytes between consecutive images.
 * Host entry: H2D copy + all kernels + (optionally) nothing else; results stay on the device
 * until sfe_get_values / sfe_download_batch.  sfe_process_batch_host synchronises before returning. */
int sfe_process_batch_host(sfe_context* ctx, const uint8_t* images, int npairs,
                           size_t stride, size_t image_stride);
/* Single pair convenience (the reference's call shape).  repeat != 0 re-runs on the images of the
 * previous call with the current thresholds (request.repeat, EST.cpp:278,285,315); left/right may
 * then be NULL. */
int sfe_process_stereo(sfe_context* ctx, const uint8_t* left, const uint8_t* right, size_t stride, int repeat);

/* Pinned host staging owned by the context, sized for max_pairs*2 tightly packed images: fill it and pass
 * it to the functions above to get asynchronous H2D copies. */
uint8_t* sfe_host_staging(sfe_context* ctx, size_t* bytes);

/* Device-resident variants: upload once, process many times (inputs already in HBM). */
int sfe_upload_batch(sfe_context* ctx, const uint8_t* images, int npairs, size_t stride, size_t image_stride);
int sfe_process_batch_device(sfe_context* ctx, int npairs);   /* asynchronous on sfe_stream()    */
int sfe_synchronize(sfe_context* ctx);                        /* also folds device error flags   */

/* ---- results ---------------------------------------------------------------------------------
 * Replaces: voEngine.getValues(kpL, kpR, descL, descR, matches, matches_ID) (CStereoSLAMKF.h:68-70).
 * Any output pointer may be NULL.  Buffers must hold sfe_keypoint_capacity() entries
 * (descriptors: capacity*32 bytes; matches: capacity entries). */
typedef struct sfe_pair_counts { int32_t n_left, n_right, n_matches, status; } sfe_pair_counts;
int sfe_get_counts(sfe_context* ctx, int pair, sfe_pair_counts* out);
int sfe_get_values(sfe_context* ctx, int pair,
                   sfe_keypoint* kp_left, sfe_keypoint* kp_right,
                   uint8_t* desc_left, uint8_t* desc_right,
                   sfe_dmatch* matches, sfe_pair_counts* counts);
/* Whole-batch download into pinned staging owned by the context; pointers stay valid until the next
 * process call.  Layout: [2*npairs][capacity] keypoints/descriptors, [npairs][capacity] matches. */
int sfe_download_batch(sfe_context* ctx, const sfe_keypoint** kps, const uint8_t** desc,
                       const sfe_dmatch** matches, const sfe_pair_counts** counts);

/* Streaming form of sfe_process_batch_host + sfe_download_batch: submit enqueues H2D copy, all kernels and the
 * D2H copy of the results on sfe_stream() and returns at once; wait blocks until they have landed in the
 * context's pinned result buffers (same layout as sfe_download_batch; entries beyond an image's count are unspecified).  Two contexts used alternately overlap
 * one batch's PCIe traffic with the other's kernels.  `images` should be pinned (sfe_host_staging). */
int sfe_submit_batch_host(sfe_context* ctx, const uint8_t* images, int npairs, size_t stride, size_t image_stride);
int sfe_wait_batch(sfe_context* ctx, const sfe_keypoint** kps, const uint8_t** desc,
                   const sfe_dmatch** matches, const sfe_pair_counts** counts);

/* Debug/parity taps: copy one pyramid (blurred=0) or blurred (blurred=1) level of image `img`
 * (index 2*pair+side) to host, tightly packed w*h. */
int sfe_get_level(sfe_context* ctx, int img, int level, int blurred, uint8_t* out);
/* Raw FAST+NMS corners of one level (unordered): packed (score<<24 | y<<12 | x). */
int sfe_get_raw_corners(sfe_context* ctx, int img, int level, uint32_t* out, int cap, int32_t* count);

/* Device pointers for zero-copy consumers (keyframe match / BoW on resident descriptors). */
int sfe_device_results(sfe_context* ctx, const void** kps, const void** desc, const void** counts,
                       const void** matches, const void** match_counts);

/* ---- brute-force Hamming match ---------------------------------------------------------------
 * Replaces: cv::BFMatcher(NORM_HAMMING,false).match(query, train, out) (EST.cpp:1447-1452) and the
 * stereo-vo inter-frame matcher ifmDescBF (EST.cpp:1137).  For each query row: argmin Hamming over
 * train rows, lowest train index on ties; out[q] = {q, t, 0, (float)dist}.  nt == 0 -> 0 matches.
 * Returns the number of matches through *n_out. */
int sfe_match_bf(sfe_context* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                 sfe_dmatch* out, int32_t* n_out);
/* Batched, device-resident: `npairs` independent (query,train) problems.  q_ofs/t_ofs are
 * npairs+1 prefix offsets (in descriptors) into d_query / d_train; d_best receives one packed
 * (dist<<16 | trainIdx) per query row (0xFFFFFFFF when the train set is empty).  All pointers are
 * device pointers; asynchronous on sfe_stream(). */
int sfe_match_bf_batch_device(sfe_context* ctx, const uint8_t* d_query, const int32_t* d_q_ofs,
                              const uint8_t* d_train, const int32_t* d_t_ofs, int npairs,
                              int max_nq, uint32_t* d_best);

/* ---- pair-list brute force over local / peer keyframe stores ---------------------------------------------
 * Loop-closure batches (BASELINE.json configs[3]): many (query keyframe, database keyframe) problems at once,
 * where the database keyframe may live on ANOTHER GPU of the box.  d_train_bases is a device array of per-rank
 * base pointers (the local one and cudaIpc-mapped peers, see sfe_ipc_*); the kernel loads peer descriptors
 * straight through NVLink -- no all-gather of the database.  Same match semantics as sfe_match_bf.
 * d_best (best_entries packed keys) is reset to 0xFFFFFFFF by the call; asynchronous on sfe_stream(). */
typedef struct sfe_bf_pair {
    int32_t q_start, nq;            /* query descriptors [q_start, q_start+nq) of d_query                     */
    int32_t t_rank, t_start, nt;    /* train descriptors [t_start, t_start+nt) of d_train_bases[t_rank]       */
    int32_t out_start;              /* d_best[out_start + q] receives (dist<<16 | train index within the kf)  */
} sfe_bf_pair;
int sfe_match_bf_pairs_device(sfe_context* ctx, const uint8_t* d_query, const uint8_t* const* d_train_bases,
                              const sfe_bf_pair* d_pairs, int npairs, int max_nq, uint32_t* d_best, size_t best_entries);
/* Raw device buffers that can be shared between the one-process-per-GPU ranks with CUDA IPC. */
int sfe_device_alloc(sfe_context* ctx, size_t bytes, void** out);
int sfe_device_free(sfe_context* ctx, void* p);
int sfe_copy_h2d(sfe_context* ctx, void* dst, const void* src, size_t bytes);
int sfe_copy_d2h(sfe_context* ctx, void* dst, const void* src, size_t bytes);
int sfe_ipc_export(sfe_context* ctx, const void* dptr, uint8_t handle[64]);
int sfe_ipc_import(sfe_context* ctx, const uint8_t handle[64], void** dptr);
int sfe_ipc_close(sfe_context* ctx, void* dptr);

/* ---- timing helpers (CUDA events on sfe_stream()) -------------------------------------------- */
int sfe_timer_start(sfe_context* ctx);
int sfe_timer_stop(sfe_context* ctx, float* ms);       /* synchronises on the stop event        */
/* Makes ctx's stream wait (on the device) for everything submitted so far on other's stream: joins two contexts that
 * are used round-robin, e.g. before sfe_timer_stop so that one event pair spans the work of both. */
int sfe_stream_wait(sfe_context* ctx, sfe_context* other);
/* Per-kernel stage times of the most recent sfe_process_batch_* call made with profiling on.
 * Stage order: see sfe_stage_name(). */
int sfe_set_profiling(sfe_context* ctx, int on);
int sfe_num_stages(void);
const char* sfe_stage_name(int stage);
int sfe_stage_times(sfe_context* ctx, float* ms /* sfe_num_stages() */, int32_t* launches);
int64_t sfe_kernel_launches(const sfe_context* ctx);   /* kernels launched (or replayed) so far  */
/* The downloads ship content, not capacity (the first ~nfeatures slots of every image, the recent maximum of matches per
 * pair; a remainder is fetched by sfe_wait_batch when a count exceeds what was shipped).  Bytes the last download moved,
 * and how often a remainder had to be fetched so far. */
int64_t sfe_last_d2h_bytes(const sfe_context* ctx);
int64_t sfe_d2h_refetches(const sfe_context* ctx);

/* ---- bag of words ----------------------------------------------------------------------------
 * Replaces: CBoWManager (CBoWManager.h:44-111) = DBoW2 BriefVocabulary + BriefDatabase.
 * The vocabulary is passed flattened, indexed by DBoW2 node id (0 = root), children in file order
 * (the C++ adapter parses voc.yml.gz).  node_desc uses OpenCV byte order (CBoWManager.h:102-106). */
typedef struct sfe_bow sfe_bow;
int sfe_bow_create(sfe_context* ctx, int nnodes, int nwords,
                   const int32_t* child_ofs /* nnodes+1 */, const int32_t* child_idx,
                   const uint8_t* node_desc /* nnodes*32 */, const double* node_weight,
                   const int32_t* node_word /* -1 for inner nodes */, sfe_bow** out);
void sfe_bow_destroy(sfe_bow* bow);
/* TemplatedVocabulary::transform(feature): word id + weight per descriptor (host buffers). */
int sfe_bow_transform(sfe_bow* bow, const uint8_t* desc, int n, int32_t* word, double* weight);
/* transform(features, BowVector): sorted distinct words + L1-normalised TF-IDF values. */
int sfe_bow_vector(sfe_bow* bow, const uint8_t* desc, int n, int32_t* out_word, double* out_val, int32_t* n_out);
/* BriefDatabase::add (CBoWManager.h:71-76): returns the new entry id through *entry_id. */
int sfe_bow_db_add(sfe_bow* bow, const uint8_t* desc, int n, int32_t* entry_id);
/* BriefDatabase::query (CBoWManager.h:83-88; EST.cpp:426 asks 4): best first, ties -> lower id. */
int sfe_bow_db_query(sfe_bow* bow, const uint8_t* desc, int n, int max_results,
                     int32_t* ids, double* scores, int32_t* n_out);
int sfe_bow_db_size(const sfe_bow* bow);
/* Device-resident batched transform (descriptors already in HBM), asynchronous. */
int sfe_bow_transform_device(sfe_bow* bow, const uint8_t* d_desc, int n, int32_t* d_word, double* d_weight);

/* ---- full front-end batch: BoW + keyframe store + data association ----------------------------------
 * Replaces, per candidate keyframe (EST.cpp): queryDB(new_kf, qResults, 4) (:424-427), the similar-KF list =
 * previous keyframe + BoW hits (:435, :1758), the matched-subset gathers (:1351-1359, :1432-1440),
 * BFMatcher.match against each (:1447-1452) and insertIntoDB (:919) -- batched, device-resident.
 * Every pair of the last processed stereo batch is appended as a keyframe; frame g only sees entries < g.
 * Results per frame: <= 4 query hits (ids are store indices, best first), 5 association slots
 * {previous keyframe, hit 0..3} with the keyframe index (-1 = empty / duplicate of the previous keyframe) and,
 * per slot, one packed (dist<<16 | train index) per stereo-matched feature (indices into the other
 * keyframe's match list, as in t_kf_da_info, srba-stereo-slam_utils.h:55-66). */
typedef struct sfe_kfdb sfe_kfdb;
int  sfe_kfdb_create(sfe_context* ctx, sfe_bow* bow, int capacity_frames, sfe_kfdb** out);
void sfe_kfdb_destroy(sfe_kfdb* db);
int  sfe_kfdb_size(const sfe_kfdb* db);
int  sfe_kfdb_reset(sfe_kfdb* db);
int  sfe_kfdb_process_batch(sfe_kfdb* db, int npairs);     /* asynchronous on sfe_stream() */
/* Same, taking the stereo batch of another context of the same GPU (double-buffered streaming): runs on that
 * context's stream; successive batches are ordered on the database by an event. */
int  sfe_kfdb_process_batch_on(sfe_kfdb* db, sfe_context* src, int npairs);
int  sfe_kfdb_results(sfe_kfdb* db, const int32_t** query_ids /* [n][4] */, const double** query_scores /* [n][4] */,
                      const int32_t** query_n /* [n] */, const int32_t** pair_kf /* [n][5] */,
                      const uint32_t** best /* [n][5][capacity] */, const int32_t** matched_n /* [n] */);

/* Streaming download of the last queried batch's results: submit enqueues the D2H copies behind the kernels (one of 4
 * pinned slots, `best` ships the filled prefix of every association row, not the capacity) and returns at once; wait
 * blocks until that slot has landed.  Layout as sfe_kfdb_results. */
int  sfe_kfdb_submit_results(sfe_kfdb* db, int32_t* slot_out);
int  sfe_kfdb_wait_results(sfe_kfdb* db, int slot, const int32_t** query_ids, const double** query_scores, const int32_t** query_n,
                           const int32_t** pair_kf, const uint32_t** best, const int32_t** matched_n, int32_t* n_frames);

/* ---- one logical keyframe database across the GPUs of a box (BASELINE.json configs[2]: frames sharded over 2/4/8 GPUs) -------
 * The reference has ONE database (CBoWManager.h:71-88, queried at EST.cpp:424-427, matched at :1432-1452); with the frames
 * sharded over ranks, rank r owns the global keyframes [starts[r], starts[r+1]) in its own stores.  Every rank exports its five
 * store arrays (matched-subset descriptors + counts, BoW words + values + counts) with sfe_ipc_export, maps its peers' with
 * sfe_ipc_import and hands the pointers over; the score / pair / brute-force kernels then read older keyframes of other ranks
 * IN-KERNEL through NVLink -- nothing is replicated, only the blocks a query touches cross the link.
 * A frame with global index g still only sees entries < g, so the results are identical to one rank processing the whole
 * sequence; that takes two passes: every rank first inserts all its frames (sfe_kfdb_insert_batch_on), a barrier, then queries
 * them (sfe_kfdb_query_batch_on).  ptrs: nranks x 5 device pointers in the order of sfe_kfdb_store_pointers (the own rank's
 * row is ignored). */
int  sfe_kfdb_store_pointers(sfe_kfdb* db, void* out[5]);
int  sfe_kfdb_set_peers(sfe_kfdb* db, int nranks, int rank, const int32_t* starts /* nranks+1 */, const void* const* ptrs /* nranks*5 */);
int  sfe_kfdb_insert_batch_on(sfe_kfdb* db, sfe_context* src, int npairs);                    /* append src's last stereo batch */
int  sfe_kfdb_query_batch_on(sfe_kfdb* db, sfe_context* stream_ctx, int first_local_frame, int npairs);

#ifdef __cplusplus
}
#endif
#endif /* SRBA_B200_H */
