// k_kfdb.cu -- the "full front-end" batch: for every stereo pair of a processed batch, BoW transform + database
// query + database add of the left descriptors, gather of the stereo-matched left descriptors into a
// device-resident keyframe store, and brute-force Hamming match against the previous keyframe and the BoW hits.
//
// Reference sequence per candidate keyframe (src/CSRBAStereoSLAMEstimator.cpp): getDataFromVOEngine (:413-418),
// m_bow_manager.queryDB(new_kf, qResults, 4) (:424-427), m_get_similar_kfs = previous KF + BoW hits (:435, :1758),
// m_data_association -> gather matched-subset descriptors (:1351-1359, :1432-1440) + BFMatcher.match (:1447-1452),
// m_bow_manager.insertIntoDB (:919).  Here every frame of the batch is treated as a keyframe (throughput mode,
// BASELINE.json configs[2]); frame g only sees database entries < g, exactly as if the frames had been
// processed one by one (DBoW2's query has the same max_id notion).
//
// BoW scoring is exact DBoW2 L1 in FP64, accumulated per entry in ascending word order (bit-identical to the
// single-keyframe path and the CPU oracle): one warp per database entry tests its sorted words against 16 query
// bitmaps held in shared memory (hits are ~2 % of the tests), hits are replayed in word order by the lane that
// owns the query.  Integer/popc work elsewhere; tensor cores unused by design.
#include "bow_internal.cuh"
#include <algorithm>
#include <cstring>

#define KF_QT 32                 // queries per score CTA (bitmaps in shared memory, one per lane)
#define KF_PAIRS 5               // previous keyframe + 4 BoW hits (EST.cpp:426, :1758)

#define KF_MAX_RANKS 16
#define KF_RESULT_SLOTS 4

struct BfPair { int q_start, nq, t_rank, t_start, nt, out_start; };

// One logical keyframe database over the GPUs of a box: rank r owns the global entries [start[r], start[r+1]) in its own
// stores; the other ranks' stores are mapped with CUDA IPC and read in-kernel through NVLink (no replication).
struct KfPeers {
    const uint8_t* msub[KF_MAX_RANKS]; const int32_t* msub_n[KF_MAX_RANKS];
    const int32_t* dbw[KF_MAX_RANKS]; const double* dbv[KF_MAX_RANKS]; const int32_t* dbn[KF_MAX_RANKS];
    int start[KF_MAX_RANKS + 1]; int nranks;
};
__device__ __forceinline__ int kf_rank_of(const KfPeers& P, int e)
{
    int r = 0;
#pragma unroll 1
    while (r + 1 < P.nranks && e >= P.start[r + 1]) ++r;
    return r;
}

struct sfe_kfdb {
    sfe_context* ctx = nullptr; sfe_bow* bow = nullptr;
    int cap_frames = 0, stride = 0, B = 0, n_frames = 0, nwords = 0, ntiles = 0;
    uint8_t* d_msub = nullptr; int32_t* d_msub_n = nullptr;
    int32_t* d_dbw = nullptr; double* d_dbv = nullptr; int32_t* d_dbn = nullptr;
    int32_t* d_word = nullptr; double* d_weight = nullptr;
    uint32_t* d_tmask = nullptr; double* d_dense = nullptr;      // [ntiles][nwords] word-major query masks; [B][nwords] dense query vectors
    double* d_score = nullptr; int32_t* d_common = nullptr; int* d_work_ctr = nullptr;
    int32_t* d_qid = nullptr; double* d_qscore = nullptr; int32_t* d_qn = nullptr;
    BfPair* d_pairs = nullptr; int32_t* d_pair_kf = nullptr; uint32_t* d_best = nullptr;
    int32_t* h_qid = nullptr; double* h_qscore = nullptr; int32_t* h_qn = nullptr; int32_t* h_pair_kf = nullptr; uint32_t* h_best = nullptr;
    int32_t* h_msub_n = nullptr;
    int last_base = 0, last_n = 0;
    cudaEvent_t ev_done = nullptr; bool ev_valid = false; cudaStream_t last_stream = nullptr;
    // multi-rank view (defaults: one rank, own stores)
    KfPeers peers; int rank = 0; int total_frames = 0;            // total_frames: score stride = capacity of the whole logical database
    // asynchronous result slots (sfe_kfdb_submit_results / sfe_kfdb_wait_results)
    struct Slot { int32_t* qid; double* qscore; int32_t* qn; int32_t* pair_kf; uint32_t* best; int32_t* msub_n; cudaEvent_t ev; int n; int ship; } slot[KF_RESULT_SLOTS] = {};
    int next_slot = 0; int ship_m = 0;
};

#define KF_CUDA(d, call) SFE_CUDA((d)->ctx, call)

__device__ __forceinline__ int kf_ham256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Same value with half the POPCs: the 8 xor words go through a carry-save adder tree (Harley-Seal), leaving
// ones/twos/fours/eights planes -> 4 POPC + 14 LOP3 instead of 8 POPC.  The brute-force kernel is bound by the
// POPC pipe (measured: 8 POPC per pair saturate it at ~44 % issue utilisation), LOP3 runs on the wider ALU pipe.
#define KF_CSA(h, l, a, b, c) { const uint32_t u_ = (a) ^ (b); h = ((a) & (b)) | (u_ & (c)); l = u_ ^ (c); }
__device__ __forceinline__ int kf_ham256_csa(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    // four carry-save adders leave two words of weight 1 (s3, x7), one of weight 2 and one of weight 4: 4 POPC like the full
    // reduction to weights 1/2/4/8, but 16 ALU operations instead of 22 - the ALU pipe (64 lanes/clk/SM), not the POPC pipe
    // (16 lanes/clk/SM, 4 x 2 cycles), was the longer of the two per distance
    uint32_t t1, s1, t2, s2, t3, s3, f1, tw1;
    KF_CSA(t1, s1, x0, x1, x2);
    KF_CSA(t2, s2, x3, x4, x5);
    KF_CSA(t3, s3, s1, s2, x6);
    KF_CSA(f1, tw1, t1, t2, t3);
    return __popc(s3) + __popc(x7) + 2 * __popc(tw1) + 4 * __popc(f1);
}

// ---- 1. gather the stereo-matched left descriptors of each pair into the keyframe store ----------------------
__global__ void kf_gather_kernel(const uint8_t* __restrict__ desc, const sfe_dmatch* __restrict__ matches,
                                 const int32_t* __restrict__ match_count, int cap, int base,
                                 uint8_t* __restrict__ msub, int32_t* __restrict__ msub_n)
{
    const int p = blockIdx.y;
    const int n = min(match_count[p], cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) msub_n[base + p] = n;
    const int k = blockIdx.x * (blockDim.x >> 1) + (threadIdx.x >> 1), half = threadIdx.x & 1;
    if (k >= n) return;
    const int qi = matches[(size_t)p * cap + k].queryIdx;
    const uint4* src = reinterpret_cast<const uint4*>(desc + ((size_t)(2 * p) * cap + qi) * 32);
    reinterpret_cast<uint4*>(msub + ((size_t)(base + p) * cap + k) * 32)[half] = __ldg(src + half);
}

// ---- 2. vocabulary transform of every left descriptor of the batch (8 lanes per descriptor) --------------------
__global__ void __launch_bounds__(256)
kf_transform_kernel(const int32_t* __restrict__ first, const int32_t* __restrict__ nchild, const uint4* __restrict__ ndesc,
                    const int32_t* __restrict__ nword, const double* __restrict__ nweight, int depth,
                    const uint8_t* __restrict__ desc, const int32_t* __restrict__ kp_count, int cap,
                    int32_t* __restrict__ word, double* __restrict__ weight)
{
    const int p = blockIdx.y;
    const int n = min(kp_count[2 * p], cap);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gtid >> 3, c = gtid & 7;
    if ((blockIdx.x * blockDim.x >> 3) >= n) return;
    const int ii = min(i, n - 1);
    const uint4* q = reinterpret_cast<const uint4*>(desc + ((size_t)(2 * p) * cap + ii) * 32);
    const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
    int node = 0;
    for (int d = 0; d < depth; ++d) {
        const int nc = __ldg(nchild + node), fc = __ldg(first + node);
        uint32_t key = 0xffffffffu;
        if (c < nc) {
            const uint4* pp = ndesc + 2 * (size_t)(fc + c);
            key = ((uint32_t)kf_ham256(q0, q1, __ldg(pp), __ldg(pp + 1)) << 4) | (uint32_t)c;
        }
        key = min(key, __shfl_xor_sync(0xffffffffu, key, 1));
        key = min(key, __shfl_xor_sync(0xffffffffu, key, 2));
        key = min(key, __shfl_xor_sync(0xffffffffu, key, 4));
        if (nc > 0) node = fc + (int)(key & 15u);
    }
    if (c == 0 && i < n) { word[(size_t)p * cap + i] = __ldg(nword + node); weight[(size_t)p * cap + i] = __ldg(nweight + node); }
}

// ---- 3. BoW vector of each pair, written straight into its (fixed-stride) database slot -------------------------
#define KV_THREADS 1024
__global__ void __launch_bounds__(KV_THREADS)
kf_vector_kernel(const int32_t* __restrict__ word, const double* __restrict__ weight, const int32_t* __restrict__ kp_count,
                 int cap, int npow2, int base, int32_t* __restrict__ dbw, double* __restrict__ dbv, int32_t* __restrict__ dbn)
{
    extern __shared__ uint32_t keys[];
    __shared__ int s_m;
    __shared__ double s_norm;
    const int p = blockIdx.x, tid = threadIdx.x;
    const int n = min(kp_count[2 * p], cap);
    const int32_t* w_in = word + (size_t)p * cap; const double* wt_in = weight + (size_t)p * cap;
    int32_t* out_w = dbw + (size_t)(base + p) * cap; double* out_v = dbv + (size_t)(base + p) * cap;
    for (int i = tid; i < npow2; i += KV_THREADS)
        keys[i] = (i < n && wt_in[i] > 0.0) ? (((uint32_t)w_in[i] << 16) | (uint32_t)i) : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += KV_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (tid == 0) s_m = 0;
    __syncthreads();
    for (int i = tid; i < npow2; i += KV_THREADS) {
        const uint32_t k = keys[i];
        if (k == 0xffffffffu) continue;
        const uint32_t w = k >> 16;
        if (i > 0 && (keys[i - 1] >> 16) == w) continue;
        double s = 0.0;
        for (int j = i; j < npow2 && keys[j] != 0xffffffffu && (keys[j] >> 16) == w; ++j) s += wt_in[keys[j] & 0xffffu];
        int rank = 0;
        for (int j = 1; j <= i; ++j) rank += (keys[j] >> 16) != (keys[j - 1] >> 16);
        out_w[rank] = (int32_t)w; out_v[rank] = s;
        atomicAdd(&s_m, 1);
    }
    __syncthreads();
    const int m = s_m;
    if (tid == 0) {
        double norm = 0.0;
        for (int i = 0; i < m; ++i) norm += fabs(out_v[i]);
        s_norm = norm; dbn[base + p] = m;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0) for (int i = tid; i < m; i += KV_THREADS) out_v[i] = out_v[i] / norm;
}

// ---- 4. query side: per 32-query tile a word-major hit mask T[tile][word] (bit q = query q of the tile contains the
//         word) + the dense vector of each pair's BoW vector (and the inverse, to clean up) -----------------------
__global__ void kf_query_prep_kernel(const int32_t* __restrict__ dbw, const double* __restrict__ dbv, const int32_t* __restrict__ dbn,
                                     int cap, int base, int nwords, uint32_t* __restrict__ tmask, double* __restrict__ dense, int clear)
{
    const int p = blockIdx.y, e = base + p;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dbn[e]) return;
    const int w = dbw[(size_t)e * cap + i];
    uint32_t* t = tmask + (size_t)(p >> 5) * nwords + w;
    if (clear) { dense[(size_t)p * nwords + w] = 0.0; *t = 0u; }
    else { dense[(size_t)p * nwords + w] = dbv[(size_t)e * cap + i]; atomicOr(t, 1u << (p & 31)); }
}

// ---- 5. exact L1 scores: lane = query (32 per CTA), warp = database entry ------------------------------------------
// Word-major hit masks per tile: ONE lookup T[w] tells which of the 32 queries contain word w.  A warp reads 32
// consecutive words of the entry (coalesced), looks their masks up, transposes the 32x32 bit matrix across the lanes
// (5 shuffle stages) so that lane q holds query q's hits, and every lane replays its own hits in ascending word order:
// word id and entry value by shuffle, the query value by one predicated gather, FP64 accumulation in ascending
// word order = DBoW2's order (bit-identical scores).  No shared memory, no queues.
#define KS_WARPS 8

__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane)      // out[q] bit i = in[i] bit q
{
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? (((y >> j) & m) | (x & ~m)) : ((x & m) | ((y << j) & ~m));
    }
    return x;
}

__global__ void __launch_bounds__(KS_WARPS * 32, 6)
kf_score_kernel(const __grid_constant__ KfPeers P, int cap,
                int base, int nq, int nwords, const uint32_t* __restrict__ tmask, const double* __restrict__ dense,
                int score_stride, double* __restrict__ score, int32_t* __restrict__ common, int ntiles, int n_items,
                int* __restrict__ work_ctr)
{
    // Persistent warps pull (entry, query tile) items from a global counter: the items are long (up to cap/32 chunks)
    // and uneven, a static split lost up to one item per warp to wave quantisation.
    const int lane = threadIdx.x & 31;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_ctr, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int e = item / ntiles, tile = item - e * ntiles;
        const int q0 = tile * KF_QT;
        const int nqt = min(KF_QT, nq - q0);
        if (e >= base + q0 + nqt - 1) continue;            // the newest query of the tile sees entries < base + q0 + nqt - 1
        // The tile's word-major masks (4 B per vocabulary word) and the dense query vectors stay in global memory: they
        // are L1/L2 resident, and a kernel without shared memory keeps 48 warps per SM to hide the gather latency.
        const uint32_t* __restrict__ T = tmask + (size_t)tile * nwords;
        const int q = q0 + lane;
        const double* __restrict__ myq = dense + (size_t)min(q, nq - 1) * nwords;
        const int er = kf_rank_of(P, e), le = e - P.start[er];        // owner rank and its local entry index
        const int n = P.dbn[er][le];
        const int32_t* ew = P.dbw[er] + (size_t)le * cap;
        const double* ev = P.dbv[er] + (size_t)le * cap;
        double s = 0.0; bool any = false;
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int i = c0 + lane;
            int w = 0; double dv = 0.0; uint32_t m = 0u;
            if (i < n) { w = __ldg(ew + i); dv = __ldg(ev + i); m = __ldg(T + w); }      // m: which queries contain word w
            uint32_t tm = transpose32(m, lane);                       // lane q: which of the chunk's 32 words hit query q
            any |= tm != 0u;
            // Every lane replays ITS OWN hits, lowest bit first = ascending word order (DBoW2's accumulation order): the entry's
            // word id and value come by shuffle from the lane that loaded them (a different source lane per lane), the query's
            // value is one predicated gather.  A lane has ~0.6 hits per 32-word chunk, the warp's worst lane ~3: the trip count
            // is that maximum, two hits per trip so that both gathers are in flight together.  (Round 1 walked all 32 words four
            // at a time for every lane - 2 % of those lane-slots held a hit: ~250 instead of ~35 instructions per chunk.  The
            // kernel stays bound by its chain of dependent loads - word, mask, query value - so the time moved by 1.4 % only;
            // loading the next chunk's words and masks ahead of the replay was measured again with this loop: slower.)
            const int trips = __reduce_max_sync(0xffffffffu, __popc(tm));
            for (int k = 0; k < trips; k += 2) {
                const bool a0 = tm != 0u; const int j0 = a0 ? __ffs(tm) - 1 : 0; tm &= tm - 1u;
                const bool a1 = tm != 0u; const int j1 = a1 ? __ffs(tm) - 1 : 0; tm &= tm - 1u;
                const int w0 = __shfl_sync(0xffffffffu, w, j0), w1 = __shfl_sync(0xffffffffu, w, j1);
                const double d0 = __shfl_sync(0xffffffffu, dv, j0), d1 = __shfl_sync(0xffffffffu, dv, j1);
                const double q0v = a0 ? __ldg(myq + w0) : 0.0, q1v = a1 ? __ldg(myq + w1) : 0.0;
                if (a0) s += fabs(q0v - d0) - fabs(q0v) - fabs(d0);
                if (a1) s += fabs(q1v - d1) - fabs(q1v) - fabs(d1);
            }
        }
        if (lane < nqt) {
            score[(size_t)q * score_stride + e] = s;
            common[(size_t)q * score_stride + e] = (any && e < base + q) ? 1 : 0;     // a frame only sees older entries
        }
    }
}

// ---- 6. top-k per query (best first, ties -> lower id), k <= 4 ----------------------------------------------------
__global__ void __launch_bounds__(256)
kf_topk_kernel(const double* __restrict__ score, int32_t* __restrict__ common, int base, int score_stride, int k,
               int32_t* __restrict__ qid, double* __restrict__ qscore, int32_t* __restrict__ qn)
{
    __shared__ double sv[8]; __shared__ int si[8];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ne = base + q;
    const double* sc = score + (size_t)q * score_stride; int32_t* cm = common + (size_t)q * score_stride;
    int found = 0;
    for (int r = 0; r < k; ++r) {
        double bv = 1e300; int bi = 0x7fffffff;
        for (int e = tid; e < ne; e += 256)
            if (cm[e] && (sc[e] < bv || (sc[e] == bv && e < bi))) { bv = sc[e]; bi = e; }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w) if (sv[w] < sv[0] || (sv[w] == sv[0] && si[w] < si[0])) { sv[0] = sv[w]; si[0] = si[w]; }
        }
        __syncthreads();
        bi = si[0]; bv = sv[0];
        if (bi == 0x7fffffff) break;
        if (tid == 0) { qid[q * 4 + found] = bi; qscore[q * 4 + found] = -bv / 2.0; cm[bi] = 0; }
        ++found;
        __syncthreads();
    }
    if (tid == 0) { qn[q] = found; for (int j = found; j < 4; ++j) { qid[q * 4 + j] = -1; qscore[q * 4 + j] = 0.0; } }
}

// ---- 7. pair list: previous keyframe + BoW hits (without duplicates of the previous keyframe) ---------------------
__global__ void kf_pairs_kernel(const __grid_constant__ KfPeers P, const int32_t* __restrict__ qid, const int32_t* __restrict__ qn,
                                int base /* global id of query 0 */, int lbase /* its local store index */, int nq, int cap,
                                BfPair* __restrict__ pairs, int32_t* __restrict__ pair_kf, int my_rank)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * KF_PAIRS) return;
    const int q = i / KF_PAIRS, slot = i - q * KF_PAIRS, g = base + q;
    int kf = -1;
    if (slot == 0) kf = g - 1;                                                         // may be -1 for the very first frame
    else if (slot - 1 < qn[q]) { kf = qid[q * 4 + slot - 1]; if (kf == g - 1) kf = -1; }
    BfPair pr;
    pr.q_start = (lbase + q) * cap; pr.nq = kf >= 0 ? P.msub_n[my_rank][lbase + q] : 0;
    const int tr = kf >= 0 ? kf_rank_of(P, kf) : 0, tl = max(kf, 0) - P.start[tr];
    pr.t_rank = tr; pr.t_start = tl * cap; pr.nt = kf >= 0 ? P.msub_n[tr][tl] : 0;
    pr.out_start = i * cap;
    pairs[i] = pr; pair_kf[i] = kf;
}

// ---- 8. brute-force Hamming over the pair list ----------------------------------------------------------------------
#define KBF_THREADS 128
#define KBF_CHUNK 128
#define KBF_SPLITS 4
__global__ void __launch_bounds__(KBF_THREADS)
kf_bf_kernel(const __grid_constant__ KfPeers P, const uint8_t* __restrict__ store, const BfPair* __restrict__ pairs, uint32_t* __restrict__ best)
{
    __shared__ uint4 ts[KBF_CHUNK * 2];
    const BfPair pr = pairs[blockIdx.z];
    if ((int)(blockIdx.x * KBF_THREADS) >= pr.nq || pr.nt == 0) return;
    const int per = (pr.nt + KBF_SPLITS - 1) / KBF_SPLITS;
    const int t0 = blockIdx.y * per, t1 = min(pr.nt, t0 + per);
    if (t0 >= t1) return;
    const int q = blockIdx.x * KBF_THREADS + threadIdx.x;
    const bool active = q < pr.nq;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (active) {
        const uint4* qp = reinterpret_cast<const uint4*>(store + (size_t)(pr.q_start + q) * 32);
        q0 = __ldg(qp); q1 = __ldg(qp + 1);
    }
    const uint4* tp = reinterpret_cast<const uint4*>(P.msub[pr.t_rank] + (size_t)pr.t_start * 32);   // possibly a peer GPU's store (NVLink)
    uint32_t key = 0xffffffffu;
    for (int c0 = t0; c0 < t1; c0 += KBF_CHUNK) {
        const int n = min(KBF_CHUNK, t1 - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += KBF_THREADS) ts[i] = __ldg(tp + 2 * (size_t)c0 + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const int d = kf_ham256_csa(q0, q1, ts[2 * j], ts[2 * j + 1]);
            key = min(key, ((uint32_t)d << 16) | (uint32_t)(c0 + j));
        }
    }
    if (active) atomicMin(&best[pr.out_start + q], key);
}

// ---- host side -------------------------------------------------------------------------------------------------------
static void kf_default_peers(sfe_kfdb* d)
{
    memset(&d->peers, 0, sizeof(d->peers));
    d->peers.nranks = 1; d->rank = 0;
    d->peers.msub[0] = d->d_msub; d->peers.msub_n[0] = d->d_msub_n; d->peers.dbw[0] = d->d_dbw; d->peers.dbv[0] = d->d_dbv; d->peers.dbn[0] = d->d_dbn;
    d->peers.start[0] = 0; d->peers.start[1] = 0x7fffffff;
}

static cudaError_t kf_alloc_scores(sfe_kfdb* d, int total_frames)
{
    cudaFree(d->d_score); cudaFree(d->d_common); d->d_score = nullptr; d->d_common = nullptr;
    cudaError_t e = cudaMalloc((void**)&d->d_score, (size_t)d->B * total_frames * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d->d_common, (size_t)d->B * total_frames * 4);
    if (e == cudaSuccess) d->total_frames = total_frames;
    return e;
}

extern "C" int sfe_kfdb_create(sfe_context* ctx, sfe_bow* bow, int capacity_frames, sfe_kfdb** out)
{
    SFE_BIND_DEVICE(ctx);
    if (!ctx || !bow || !out || capacity_frames < 1 || capacity_frames > (1 << 20)) return SFE_ERR_INVALID;
    sfe_kfdb* d = new sfe_kfdb(); d->ctx = ctx; d->bow = bow;
    d->cap_frames = capacity_frames; d->stride = ctx->pack.kp_cap; d->B = ctx->cfg.max_pairs;
    d->nwords = bow->nwords; d->ntiles = (ctx->cfg.max_pairs + KF_QT - 1) / KF_QT;
    const size_t F = capacity_frames, S = d->stride, B = d->B;
    cudaError_t e = cudaSuccess;
    auto al = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    auto ah = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMallocHost(p, bytes); };
    al((void**)&d->d_msub, F * S * 32); al((void**)&d->d_msub_n, F * 4);
    al((void**)&d->d_dbw, F * S * 4); al((void**)&d->d_dbv, F * S * 8); al((void**)&d->d_dbn, F * 4);
    al((void**)&d->d_word, B * S * 4); al((void**)&d->d_weight, B * S * 8);
    al((void**)&d->d_tmask, (size_t)d->ntiles * d->nwords * 4); al((void**)&d->d_dense, B * (size_t)d->nwords * 8);
    al((void**)&d->d_work_ctr, 256);
    al((void**)&d->d_qid, B * 4 * 4); al((void**)&d->d_qscore, B * 4 * 8); al((void**)&d->d_qn, B * 4);
    al((void**)&d->d_pairs, B * KF_PAIRS * sizeof(BfPair)); al((void**)&d->d_pair_kf, B * KF_PAIRS * 4);
    al((void**)&d->d_best, B * KF_PAIRS * S * 4);
    if (e == cudaSuccess) e = kf_alloc_scores(d, capacity_frames);
    ah((void**)&d->h_qid, B * 4 * 4); ah((void**)&d->h_qscore, B * 4 * 8); ah((void**)&d->h_qn, B * 4);
    ah((void**)&d->h_pair_kf, B * KF_PAIRS * 4); ah((void**)&d->h_best, B * KF_PAIRS * S * 4); ah((void**)&d->h_msub_n, B * 4);
    for (int k = 0; k < KF_RESULT_SLOTS; ++k) {
        sfe_kfdb::Slot& sl = d->slot[k];
        ah((void**)&sl.qid, B * 4 * 4); ah((void**)&sl.qscore, B * 4 * 8); ah((void**)&sl.qn, B * 4); ah((void**)&sl.pair_kf, B * KF_PAIRS * 4);
        ah((void**)&sl.best, B * KF_PAIRS * S * 4); ah((void**)&sl.msub_n, B * 4);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMemset(d->d_tmask, 0, (size_t)d->ntiles * d->nwords * 4);
    if (e == cudaSuccess) e = cudaMemset(d->d_dense, 0, B * (size_t)d->nwords * 8);
    if (e == cudaSuccess) e = cudaMemset(d->d_msub_n, 0, F * 4);
    if (e == cudaSuccess) e = cudaMemset(d->d_dbn, 0, F * 4);
    if (e != cudaSuccess) { ctx->err = std::string("sfe_kfdb_create: ") + cudaGetErrorString(e); sfe_kfdb_destroy(d); return SFE_ERR_CUDA; }
    sfe_raise_dynamic_smem((const void*)kf_vector_kernel, ctx->cfg.device, 32768 * 4);
    cudaGetLastError();
    kf_default_peers(d);
    d->ship_m = std::min((int)S, ctx->cfg.nfeatures + 16);
    *out = d;
    return SFE_OK;
}

extern "C" void sfe_kfdb_destroy(sfe_kfdb* d)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    if (!d) return;
    cudaFree(d->d_msub); cudaFree(d->d_msub_n); cudaFree(d->d_dbw); cudaFree(d->d_dbv); cudaFree(d->d_dbn);
    cudaFree(d->d_word); cudaFree(d->d_weight); cudaFree(d->d_tmask); cudaFree(d->d_dense); cudaFree(d->d_score);
    cudaFree(d->d_common); cudaFree(d->d_work_ctr); cudaFree(d->d_qid); cudaFree(d->d_qscore); cudaFree(d->d_qn); cudaFree(d->d_pairs);
    cudaFree(d->d_pair_kf); cudaFree(d->d_best);
    cudaFreeHost(d->h_qid); cudaFreeHost(d->h_qscore); cudaFreeHost(d->h_qn); cudaFreeHost(d->h_pair_kf); cudaFreeHost(d->h_best);
    cudaFreeHost(d->h_msub_n);
    for (int k = 0; k < KF_RESULT_SLOTS; ++k) {
        sfe_kfdb::Slot& sl = d->slot[k];
        cudaFreeHost(sl.qid); cudaFreeHost(sl.qscore); cudaFreeHost(sl.qn); cudaFreeHost(sl.pair_kf); cudaFreeHost(sl.best); cudaFreeHost(sl.msub_n);
        if (sl.ev) cudaEventDestroy(sl.ev);
    }
    if (d->ev_done) cudaEventDestroy(d->ev_done);
    delete d;
}

extern "C" int sfe_kfdb_size(const sfe_kfdb* d) { return d ? d->n_frames : 0; }
extern "C" int sfe_kfdb_reset(sfe_kfdb* d) { if (!d) return SFE_ERR_INVALID; d->n_frames = 0; return SFE_OK; }

// ---- one logical database over several GPUs ----------------------------------------------------------------------------------
extern "C" int sfe_kfdb_store_pointers(sfe_kfdb* d, void* out[5])
{
    if (!d || !out) return SFE_ERR_INVALID;
    out[0] = d->d_msub; out[1] = d->d_msub_n; out[2] = d->d_dbw; out[3] = d->d_dbv; out[4] = d->d_dbn;
    return SFE_OK;
}

extern "C" int sfe_kfdb_set_peers(sfe_kfdb* d, int nranks, int rank, const int32_t* starts, const void* const* ptrs)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    if (!d) return SFE_ERR_INVALID;
    if (nranks < 1 || nranks > KF_MAX_RANKS || rank < 0 || rank >= nranks || !starts || !ptrs) { d->ctx->err = "sfe_kfdb_set_peers: bad arguments"; return SFE_ERR_INVALID; }
    for (int r = 0; r < nranks; ++r) if (starts[r + 1] < starts[r] || (r == rank && starts[r + 1] - starts[r] > d->cap_frames)) { d->ctx->err = "sfe_kfdb_set_peers: shard does not fit the local store"; return SFE_ERR_INVALID; }
    KF_CUDA(d, cudaStreamSynchronize(d->last_stream ? d->last_stream : d->ctx->stream));
    memset(&d->peers, 0, sizeof(d->peers));
    d->peers.nranks = nranks; d->rank = rank;
    for (int r = 0; r < nranks; ++r) {
        const void* const* p = ptrs + 5 * r;
        d->peers.msub[r] = (const uint8_t*)(r == rank ? d->d_msub : p[0]); d->peers.msub_n[r] = (const int32_t*)(r == rank ? d->d_msub_n : p[1]);
        d->peers.dbw[r] = (const int32_t*)(r == rank ? d->d_dbw : p[2]); d->peers.dbv[r] = (const double*)(r == rank ? d->d_dbv : p[3]);
        d->peers.dbn[r] = (const int32_t*)(r == rank ? d->d_dbn : p[4]);
        d->peers.start[r] = starts[r];
    }
    d->peers.start[nranks] = starts[nranks];
    if (starts[nranks] > d->total_frames) {
        const cudaError_t e = kf_alloc_scores(d, starts[nranks]);
        if (e != cudaSuccess) { d->ctx->err = std::string("sfe_kfdb_set_peers: ") + cudaGetErrorString(e); return SFE_ERR_CUDA; }
    }
    return SFE_OK;
}

static int kf_check_src(sfe_kfdb* d, sfe_context* c, int npairs)
{
    if (!d || !c || npairs < 1 || npairs > d->B) return SFE_ERR_INVALID;
    if (c->pack.kp_cap != d->stride || c->cfg.device != d->ctx->cfg.device) { c->err = "source context does not match the keyframe database"; return SFE_ERR_INVALID; }
    return SFE_OK;
}

// Pass 1 (steps 1-3): append every pair of the source context's last stereo batch to the LOCAL stores.
static int kf_insert(sfe_kfdb* d, sfe_context* c, int npairs, cudaStream_t s)
{
    sfe_bow* b = d->bow;
    if (!c->have_results || c->last_npairs < npairs) { c->err = "sfe_kfdb: process the stereo batch first"; return SFE_ERR_STATE; }
    if (d->n_frames + npairs > d->cap_frames) { c->err = "keyframe store is full"; return SFE_ERR_CAPACITY; }
    const int cap = d->stride, base = d->n_frames;
    kf_gather_kernel<<<dim3((cap + 127) / 128, npairs), 256, 0, s>>>(c->d_desc, c->d_matches, c->d_match_count, cap, base, d->d_msub, d->d_msub_n);
    kf_transform_kernel<<<dim3((cap * 8 + 255) / 256, npairs), 256, 0, s>>>(b->d_first, b->d_nchild, b->d_desc, b->d_word, b->d_weight, b->depth,
                                                                            c->d_desc, c->d_kp_count, cap, d->d_word, d->d_weight);
    int np2 = 1; while (np2 < cap) np2 <<= 1;
    if (np2 > 32768) { c->err = "keypoint capacity too large for the BoW vector kernel"; return SFE_ERR_INVALID; }
    kf_vector_kernel<<<npairs, KV_THREADS, (size_t)np2 * 4, s>>>(d->d_word, d->d_weight, c->d_kp_count, cap, np2, base, d->d_dbw, d->d_dbv, d->d_dbn);
    c->launches += 3;
    d->n_frames = base + npairs;
    return SFE_OK;
}

// Pass 2 (steps 4-8): query + associate the local frames [lbase, lbase + npairs) against every OLDER entry of the logical
// database (all ranks), top 4, pair list, brute force.
static int kf_query(sfe_kfdb* d, sfe_context* c, int lbase, int npairs, cudaStream_t s)
{
    if (lbase < 0 || lbase + npairs > d->n_frames) { c->err = "sfe_kfdb_query: frames not in the store"; return SFE_ERR_INVALID; }
    const int cap = d->stride, gbase = d->peers.start[d->rank] + lbase;
    if (gbase + npairs > d->total_frames) { c->err = "sfe_kfdb_query: score buffer smaller than the logical database"; return SFE_ERR_CAPACITY; }
    kf_query_prep_kernel<<<dim3((cap + 255) / 256, npairs), 256, 0, s>>>(d->d_dbw, d->d_dbv, d->d_dbn, cap, lbase, d->nwords, d->d_tmask, d->d_dense, 0);
    const int ne_vis = gbase + npairs - 1;                      // entries any query of this batch can see
    if (ne_vis > 0) {
        const int tiles = (npairs + KF_QT - 1) / KF_QT;
        const int n_items = ne_vis * tiles;
        const int gx = std::max(1, std::min((n_items + KS_WARPS - 1) / KS_WARPS, 148 * 6));       // persistent: 6 CTAs per SM
        KF_CUDA(d, cudaMemsetAsync(d->d_work_ctr, 0, sizeof(int), s));
        kf_score_kernel<<<gx, KS_WARPS * 32, 0, s>>>(d->peers, cap, gbase, npairs, d->nwords, d->d_tmask, d->d_dense,
                                                     d->total_frames, d->d_score, d->d_common, tiles, n_items, d->d_work_ctr);
    }
    kf_topk_kernel<<<npairs, 256, 0, s>>>(d->d_score, d->d_common, gbase, d->total_frames, 4, d->d_qid, d->d_qscore, d->d_qn);
    kf_query_prep_kernel<<<dim3((cap + 255) / 256, npairs), 256, 0, s>>>(d->d_dbw, d->d_dbv, d->d_dbn, cap, lbase, d->nwords, d->d_tmask, d->d_dense, 1);
    kf_pairs_kernel<<<(npairs * KF_PAIRS + 127) / 128, 128, 0, s>>>(d->peers, d->d_qid, d->d_qn, gbase, lbase, npairs, cap, d->d_pairs, d->d_pair_kf, d->rank);
    KF_CUDA(d, cudaMemsetAsync(d->d_best, 0xff, (size_t)npairs * KF_PAIRS * cap * 4, s));
    kf_bf_kernel<<<dim3((cap + KBF_THREADS - 1) / KBF_THREADS, KBF_SPLITS, npairs * KF_PAIRS), KBF_THREADS, 0, s>>>(d->peers, d->d_msub, d->d_pairs, d->d_best);
    c->launches += ne_vis > 0 ? 6 : 5;
    d->last_base = lbase; d->last_n = npairs;
    return SFE_OK;
}

extern "C" int sfe_kfdb_process_batch_on(sfe_kfdb* d, sfe_context* src, int npairs);
extern "C" int sfe_kfdb_process_batch(sfe_kfdb* d, int npairs)
{
    return d ? sfe_kfdb_process_batch_on(d, d->ctx, npairs) : SFE_ERR_INVALID;
}

// mode bit 0: insert, bit 1: query.  Runs on the SOURCE context's stream (so that context's next batch cannot overwrite the
// descriptors being read); batches from different contexts are serialised on the database through an event.
static int kf_run(sfe_kfdb* d, sfe_context* c, int lbase, int npairs, int mode)
{
    int r = kf_check_src(d, c, npairs); if (r) return r;
    cudaStream_t s = c->stream;
    if (d->ev_valid && d->last_stream != s) KF_CUDA(d, cudaStreamWaitEvent(s, d->ev_done, 0));
    if (mode & 1) { lbase = d->n_frames; r = kf_insert(d, c, npairs, s); if (r) return r; }
    if (mode & 2) { r = kf_query(d, c, lbase, npairs, s); if (r) return r; }
    KF_CUDA(d, cudaGetLastError());
    KF_CUDA(d, cudaEventRecord(d->ev_done, s));
    d->ev_valid = true; d->last_stream = s;
    return SFE_OK;
}

extern "C" int sfe_kfdb_process_batch_on(sfe_kfdb* d, sfe_context* c, int npairs)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    return kf_run(d, c, 0, npairs, 3);
}
extern "C" int sfe_kfdb_insert_batch_on(sfe_kfdb* d, sfe_context* c, int npairs)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    return kf_run(d, c, 0, npairs, 1);
}
extern "C" int sfe_kfdb_query_batch_on(sfe_kfdb* d, sfe_context* c, int first_local_frame, int npairs)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    return kf_run(d, c, first_local_frame, npairs, 2);
}

extern "C" int sfe_kfdb_results(sfe_kfdb* d, const int32_t** query_ids, const double** query_scores, const int32_t** query_n,
                                const int32_t** pair_kf, const uint32_t** best, const int32_t** matched_n)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    if (!d) return SFE_ERR_INVALID;
    cudaStream_t s = d->last_stream ? d->last_stream : d->ctx->stream; const int n = d->last_n; const size_t cap = d->stride;
    if (n == 0) { d->ctx->err = "no batch processed"; return SFE_ERR_STATE; }
    KF_CUDA(d, cudaMemcpyAsync(d->h_qid, d->d_qid, sizeof(int32_t) * 4 * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(d->h_qscore, d->d_qscore, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(d->h_qn, d->d_qn, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(d->h_pair_kf, d->d_pair_kf, sizeof(int32_t) * KF_PAIRS * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(d->h_best, d->d_best, sizeof(uint32_t) * KF_PAIRS * n * cap, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(d->h_msub_n, d->d_msub_n + d->last_base, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaStreamSynchronize(s));
    if (query_ids) *query_ids = d->h_qid; if (query_scores) *query_scores = d->h_qscore; if (query_n) *query_n = d->h_qn;
    if (pair_kf) *pair_kf = d->h_pair_kf; if (best) *best = d->h_best; if (matched_n) *matched_n = d->h_msub_n;
    return SFE_OK;
}

// Streaming form: enqueue the download of the last queried batch's results right behind its kernels (same stream), into one of
// KF_RESULT_SLOTS pinned slots; the next batch may overwrite the device buffers as soon as this copy has run.  `best` ships the
// first ship_m slots of every association row (content, not capacity); a remainder is fetched by the wait if a frame had more matches.
extern "C" int sfe_kfdb_submit_results(sfe_kfdb* d, int32_t* slot_out)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    if (!d || !slot_out) return SFE_ERR_INVALID;
    if (d->last_n == 0 || !d->last_stream) { d->ctx->err = "no batch processed"; return SFE_ERR_STATE; }
    const int k = d->next_slot; d->next_slot = (k + 1) % KF_RESULT_SLOTS;
    sfe_kfdb::Slot& sl = d->slot[k];
    cudaStream_t s = d->last_stream; const int n = d->last_n; const size_t cap = d->stride;
    KF_CUDA(d, cudaMemcpyAsync(sl.qid, d->d_qid, sizeof(int32_t) * 4 * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(sl.qscore, d->d_qscore, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(sl.qn, d->d_qn, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(sl.pair_kf, d->d_pair_kf, sizeof(int32_t) * KF_PAIRS * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpyAsync(sl.msub_n, d->d_msub_n + d->last_base, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaMemcpy2DAsync(sl.best, cap * 4, d->d_best, cap * 4, (size_t)d->ship_m * 4, (size_t)KF_PAIRS * n, cudaMemcpyDeviceToHost, s));
    KF_CUDA(d, cudaEventRecord(sl.ev, s));
    KF_CUDA(d, cudaEventRecord(d->ev_done, s));              // the next batch's kernels (any stream) wait for this copy too
    sl.n = n; sl.ship = d->ship_m;
    *slot_out = k;
    return SFE_OK;
}

extern "C" int sfe_kfdb_wait_results(sfe_kfdb* d, int slot, const int32_t** query_ids, const double** query_scores, const int32_t** query_n,
                                     const int32_t** pair_kf, const uint32_t** best, const int32_t** matched_n, int32_t* n_frames)
{
    SFE_BIND_DEVICE((d ? d->ctx : nullptr));
    if (!d || slot < 0 || slot >= KF_RESULT_SLOTS) return SFE_ERR_INVALID;
    sfe_kfdb::Slot& sl = d->slot[slot];
    if (sl.n == 0) { d->ctx->err = "sfe_kfdb_wait_results: empty slot"; return SFE_ERR_STATE; }
    KF_CUDA(d, cudaEventSynchronize(sl.ev));
    int mmax = 0; for (int i = 0; i < sl.n; ++i) mmax = std::max(mmax, std::min(sl.msub_n[i], d->stride));
    if (mmax > sl.ship) {
        // mis-speculation: the rows are still in d_best only if no later batch has been queried -- which the streaming protocol
        // (submit right after the query, at most KF_RESULT_SLOTS in flight) cannot guarantee; ship the full rows next time and say so.
        d->ship_m = d->stride;
        d->ctx->err = "sfe_kfdb_wait_results: a frame had more matches than the shipped prefix; resubmit (the prefix now covers the capacity)";
        return SFE_ERR_CAPACITY;
    }
    d->ship_m = std::min(d->stride, std::max(d->ship_m, mmax + 64));
    if (query_ids) *query_ids = sl.qid; if (query_scores) *query_scores = sl.qscore; if (query_n) *query_n = sl.qn;
    if (pair_kf) *pair_kf = sl.pair_kf; if (best) *best = sl.best; if (matched_n) *matched_n = sl.msub_n; if (n_frames) *n_frames = sl.n;
    return SFE_OK;
}
