// k_match.cu -- Hamming matchers on 256-bit descriptors: row-band left/right match and brute force.
//
// (1) Row-band L/R match: replaces stereo-vo's smDescRbR stage forced at reference EST.cpp:1136 and run
//     inside processNewImagePair (EST.cpp:112,267) with [MATCH] max_y_diff / orb_max_distance
//     (kitti.ini:17,24) and the run-time ORB threshold (EST.cpp:284,304,311).  Conventions (stereo-vo is
//     not vendored): candidates for left i are right j with |yl-yr| <= max_y_diff and
//     min_disp <= xl-xr <= max_disp; best = min Hamming, lowest j on ties; accepted iff
//     dist <= threshold; optional mutual-best (robust 1-to-1).  Output DMatch{left idx, right idx} in
//     left order (reference KF.cpp:43-44, EST.cpp:147-151).
// (2) Brute force: replaces cv::BFMatcher(NORM_HAMMING,false).match (reference EST.cpp:1447-1452) and
//     the inter-frame matcher ifmDescBF (EST.cpp:1137): argmin over all train rows, lowest index on ties.
//
// Integer/popc work: xor + __popc on eight 32-bit words per pair, packed (dist<<16 | idx) keys so the
// argmin with the lowest-index tie rule is a plain unsigned min (warp shuffles / atomicMin).
// Tensor cores are deliberately not used (BASELINE.json north_star).
#include "sfe_internal.cuh"
#include <algorithm>

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ------------------------------------------------------------------------------------------------
// Row-band match.  Kernel 1: grid (slices, pairs) -- every CTA buckets the right keypoints of its pair by
// image row in shared memory (counting sort, parallel scan) and matches its slice of the left keypoints,
// one half-warp per left keypoint over the +-max_y_diff band.  Kernel 2: one CTA per pair, ordered compaction
// (+ the mutual-best check) into the cv::DMatch list.
// ------------------------------------------------------------------------------------------------
#define RB_THREADS 256
#ifndef RB_LANES
#define RB_LANES 16               // lanes per left keypoint
#endif
#ifndef RB_WAVE
#define RB_WAVE (148 * 8)
#endif

struct RowBandArgs {
    const sfe_keypoint* kps; const uint8_t* desc; const int32_t* kp_count;
    sfe_dmatch* matches; int32_t* match_count;
    uint32_t* best; uint32_t* owner;           // [pairs][kp_cap] packed (dist<<16 | idx) / (dist<<16 | left idx)
    int kp_cap, height, slices;
    float max_y_diff, min_disp, max_disp;
    int max_dist, robust;
};

__global__ void __launch_bounds__(RB_THREADS, 8)     // 32 registers: the slice count below assumes 8 CTAs per SM (one wave); 39 registers cost a second wave, +16 us
rowband_match_kernel(const RowBandArgs a)
{
    extern __shared__ uint32_t dynm[];
    // layout: row_start[height+2] | sorted[kp_cap]
    int* row_start = reinterpret_cast<int*>(dynm);
    int* sorted = row_start + a.height + 2;
    __shared__ int s_warp[RB_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.y;
    const int il = 2 * pair, ir = 2 * pair + 1;
    pdl_wait();                                                   // keypoints and descriptors are the previous kernel's output
    const int nl = min(a.kp_count[il], a.kp_cap), nr = min(a.kp_count[ir], a.kp_cap);
    const sfe_keypoint* kl = a.kps + (size_t)il * a.kp_cap;
    const sfe_keypoint* kr = a.kps + (size_t)ir * a.kp_cap;
    const uint4* dl = reinterpret_cast<const uint4*>(a.desc + (size_t)il * a.kp_cap * 32);
    const uint4* dr = reinterpret_cast<const uint4*>(a.desc + (size_t)ir * a.kp_cap * 32);
    const int H = a.height;
    const int chunk = (nl + a.slices - 1) / a.slices;
    const int i0 = blockIdx.x * chunk, i1 = min(nl, i0 + chunk);
    if (i0 >= i1) return;

    // ---- bucket the right keypoints by image row (counting sort) ----------------------------
    for (int i = tid; i < H + 2; i += RB_THREADS) row_start[i] = 0;
    __syncthreads();
    for (int j = tid; j < nr; j += RB_THREADS) {
        const int r = min(max((int)floorf(kr[j].y), 0), H - 1);
        atomicAdd(&row_start[r + 2], 1);
    }
    __syncthreads();
    {   // inclusive scan of row_start[2 .. H+1] in place: per-thread segments + warp/block scan of the segment sums
        const int seg = (H + RB_THREADS - 1) / RB_THREADS;
        const int b = 2 + tid * seg, e = min(b + seg, H + 2);
        int sum = 0;
        for (int r = b; r < e; ++r) sum += row_start[r];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int wofs = 0;
        for (int w = 0; w < warp; ++w) wofs += s_warp[w];
        int run = wofs + incl - sum;
        for (int r = b; r < e; ++r) { run += row_start[r]; row_start[r] = run; }
    }
    __syncthreads();
    // row_start[r+1] = begin of row r (after the scatter below: its end = begin of row r+1)
    for (int j = tid; j < nr; j += RB_THREADS) {
        const int r = min(max((int)floorf(kr[j].y), 0), H - 1);
        const int p = atomicAdd(&row_start[r + 1], 1);
        sorted[p] = j;
    }
    __syncthreads();
    // now row r occupies sorted[row_start[r] .. row_start[r+1])

    // ---- one warp per left keypoint of this slice ------------------------------------------------
    uint32_t* best = a.best + (size_t)pair * a.kp_cap;
    uint32_t* owner = a.owner + (size_t)pair * a.kp_cap;
    // RB_LANES lanes per left keypoint: a +-2 row band holds ~13 candidates at 1000 keypoints, a half-warp is enough, and two
    // independent keypoints per warp halve the number of trips through the chain of dependent loads (0.044 -> 0.038 ms; measured
    // at 32 registers - at 39 the grid no longer fits one wave and every gain is lost)
    constexpr int KPW = 32 / RB_LANES;                            // left keypoints a warp matches at once
    const int half = lane / RB_LANES, hl = lane % RB_LANES;
    for (int ib = i0 + KPW * warp; ib < i1; ib += KPW * (RB_THREADS / 32)) {     // warp-uniform trip count
        const int i = ib + half;
        const bool valid = i < i1;
        uint32_t key = 0xffffffffu;
        if (valid) {
            const float xl = kl[i].x, yl = kl[i].y;
            const uint4 q0 = __ldg(dl + 2 * i), q1 = __ldg(dl + 2 * i + 1);
            const int rlo = min(max((int)floorf(yl - a.max_y_diff), 0), H - 1);
            const int rhi = min(max((int)floorf(yl + a.max_y_diff), 0), H - 1);
            const int beg = row_start[rlo], end = row_start[rhi + 1];
            for (int p = beg + hl; p < end; p += RB_LANES) {
                const int j = sorted[p];
                const float dy = fabsf(yl - kr[j].y), disp = xl - kr[j].x;
                if (dy <= a.max_y_diff && disp >= a.min_disp && disp <= a.max_disp) {
                    const int d = hamming256(q0, q1, __ldg(dr + 2 * j), __ldg(dr + 2 * j + 1));
                    key = min(key, ((uint32_t)d << 16) | (uint32_t)j);
                }
            }
        }
#pragma unroll
        for (int o = RB_LANES / 2; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));     // stays inside the lane group
        if (valid && hl == 0) {
            if (key != 0xffffffffu && (int)(key >> 16) > a.max_dist) key = 0xffffffffu;
            best[i] = key;
            if (a.robust && key != 0xffffffffu) atomicMin(&owner[key & 0xffffu], (key & 0xffff0000u) | (uint32_t)i);
        }
    }
}

__global__ void __launch_bounds__(RB_THREADS)
rowband_compact_kernel(const RowBandArgs a)
{
    __shared__ int s_scan[RB_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.x;
    pdl_wait();
    const int nl = min(a.kp_count[2 * pair], a.kp_cap);
    const uint32_t* best = a.best + (size_t)pair * a.kp_cap;
    const uint32_t* owner = a.owner + (size_t)pair * a.kp_cap;
    sfe_dmatch* out = a.matches + (size_t)pair * a.kp_cap;
    if (tid == 0) s_total = 0;
    __syncthreads();
    for (int base = 0; base < nl; base += RB_THREADS) {
        const int i = base + tid;
        bool ok = false; uint32_t key = 0xffffffffu;
        if (i < nl) {
            key = best[i];
            ok = key != 0xffffffffu;
            if (ok && a.robust) ok = (owner[key & 0xffffu] & 0xffffu) == (uint32_t)i;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_scan[warp] = __popc(bal);
        __syncthreads();
        int wofs = s_total;
        for (int w = 0; w < warp; ++w) wofs += s_scan[w];
        if (ok) {
            sfe_dmatch m;
            m.queryIdx = i; m.trainIdx = (int)(key & 0xffffu); m.imgIdx = 0; m.distance = (float)(key >> 16);
            out[wofs + __popc(bal & ((1u << lane) - 1u))] = m;
        }
        __syncthreads();
        if (tid == 0) { int t = s_total; for (int w = 0; w < RB_THREADS / 32; ++w) t += s_scan[w]; s_total = t; }
        __syncthreads();
    }
    if (tid == 0) a.match_count[pair] = s_total;
}

static size_t rowband_smem(const sfe_context* c)
{
    return sizeof(int) * ((size_t)c->cfg.height + 2 + (size_t)c->pack.kp_cap);
}

void prepare_rowband(sfe_context* c)
{
    // the attribute belongs to the kernel function (per device), not to the context: never lower it, or a second,
    // smaller context would break the launches of an earlier, larger one
    sfe_raise_dynamic_smem((const void*)rowband_match_kernel, c->cfg.device, (int)rowband_smem(c));
}

void launch_match_rowband(sfe_context* c, int npairs)
{
    RowBandArgs a;
    a.kps = c->d_kp; a.desc = c->d_desc; a.kp_count = c->d_kp_count;
    a.matches = c->d_matches; a.match_count = c->d_match_count;
    a.best = c->d_rb_best; a.owner = c->d_rb_owner;
    a.kp_cap = c->pack.kp_cap; a.height = c->cfg.height;
    a.max_y_diff = c->cfg.max_y_diff; a.min_disp = c->cfg.min_disparity; a.max_disp = c->cfg.max_disparity;
    a.max_dist = c->orb_th; a.robust = c->cfg.robust_1to1;
    // As many slices as still make ONE wave of 8 CTAs per SM (the kernel is a latency chain per CTA: a second, partial wave
    // doubles its time), at least 64 left keypoints per slice.  128 pairs: 5 / 6 / 7 / 9 / 10 slices measured
    // 0.054 / 0.050 / 0.046 / 0.044 / 0.054 ms.
    a.slices = std::max(1, std::min({ RB_WAVE / npairs, (c->cfg.nfeatures + 63) / 64, 32 }));
    if (a.robust) cudaMemsetAsync(c->d_rb_owner, 0xff, sizeof(uint32_t) * (size_t)npairs * a.kp_cap, c->stream);
    sfe_launch(c, !a.robust && (c->pdl_mask & 16), rowband_match_kernel, dim3(a.slices, npairs), dim3(RB_THREADS), rowband_smem(c), a);     // robust: a memset precedes it
    sfe_launch(c, (c->pdl_mask & 32) != 0, rowband_compact_kernel, dim3(npairs), dim3(RB_THREADS), 0, a);
}

// ------------------------------------------------------------------------------------------------
// Brute force: grid (query tiles, train splits, pairs).  Each thread owns one query descriptor in
// registers; the train split is streamed through shared memory in chunks and read as broadcasts.
// Partial results merge with atomicMin on the packed key (d_best must be pre-set to 0xFFFFFFFF).
// ------------------------------------------------------------------------------------------------
#define BF_THREADS 128
#define BF_CHUNK 128
#define BF_SPLITS 8

__global__ void bf_init_kernel(uint32_t* best, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) best[i] = 0xffffffffu;
}

__global__ void __launch_bounds__(BF_THREADS)
bf_kernel(const uint8_t* __restrict__ dq, const int32_t* __restrict__ q_ofs,
          const uint8_t* __restrict__ dt, const int32_t* __restrict__ t_ofs, uint32_t* __restrict__ best)
{
    __shared__ uint4 ts[BF_CHUNK * 2];
    const int pair = blockIdx.z;
    const int qb = q_ofs[pair], nq = q_ofs[pair + 1] - qb;
    const int tb = t_ofs[pair], nt = t_ofs[pair + 1] - tb;
    const int q = blockIdx.x * BF_THREADS + threadIdx.x;
    if (blockIdx.x * BF_THREADS >= nq) return;
    // this CTA's slice of the train set
    const int per = (nt + BF_SPLITS - 1) / BF_SPLITS;
    const int t0 = blockIdx.y * per, t1 = min(nt, t0 + per);
    if (t0 >= t1) return;
    const bool active = q < nq;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (active) {
        const uint4* qp = reinterpret_cast<const uint4*>(dq + (size_t)(qb + q) * 32);
        q0 = __ldg(qp); q1 = __ldg(qp + 1);
    }
    const uint4* tp = reinterpret_cast<const uint4*>(dt + (size_t)tb * 32);
    uint32_t key = 0xffffffffu;
    for (int c0 = t0; c0 < t1; c0 += BF_CHUNK) {
        const int n = min(BF_CHUNK, t1 - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += BF_THREADS) ts[i] = __ldg(tp + 2 * (size_t)c0 + i);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const int d = hamming256(q0, q1, ts[2 * j], ts[2 * j + 1]);
            key = min(key, ((uint32_t)d << 16) | (uint32_t)(c0 + j));
        }
    }
    if (active) atomicMin(&best[qb + q], key);
}

void launch_match_bf(cudaStream_t s, const uint8_t* d_query, const int32_t* d_q_ofs, const uint8_t* d_train,
                     const int32_t* d_t_ofs, int npairs, int max_nq, uint32_t* d_best)
{
    if (npairs <= 0 || max_nq <= 0) return;
    dim3 grid((max_nq + BF_THREADS - 1) / BF_THREADS, BF_SPLITS, npairs);
    bf_kernel<<<grid, BF_THREADS, 0, s>>>(d_query, d_q_ofs, d_train, d_t_ofs, d_best);
}

void launch_bf_init(cudaStream_t s, uint32_t* best, int n)
{
    if (n <= 0) return;
    bf_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(best, n);
}

// Same, for the batched entry point whose total query count q_ofs[npairs] only exists on the device.
__global__ void bf_init_ofs_kernel(uint32_t* best, const int32_t* __restrict__ q_ofs, int npairs)
{
    const int n = q_ofs[npairs];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) best[i] = 0xffffffffu;
}
void launch_bf_init_ofs(cudaStream_t s, uint32_t* best, const int32_t* d_q_ofs, int npairs)
{
    bf_init_ofs_kernel<<<148 * 4, 256, 0, s>>>(best, d_q_ofs, npairs);
}


// ------------------------------------------------------------------------------------------------
// Pair-list brute force with the train side possibly living in a PEER GPU's memory (loop-closure
// batches, BASELINE.json configs[3]): every pair names the rank that owns its train keyframe and the
// kernel loads those descriptors straight through NVLink (ld.global on cudaIpc-mapped peer pointers),
// staged through shared memory exactly like the local case.  No all-gather, no staging copy: only the
// candidate blocks that are actually matched cross the link, overlapped with the popc work of other CTAs.
// ------------------------------------------------------------------------------------------------
#define PBF_THREADS 128
#define PBF_CHUNK 128
#define PBF_SPLITS 4
#define PBF_CSA(h, l, a, b, c) { const uint32_t u_ = (a) ^ (b); h = ((a) & (b)) | (u_ & (c)); l = u_ ^ (c); }
__device__ __forceinline__ int hamming256_csa(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    // four carry-save adders leave two words of weight 1 (s3, x7), one of weight 2 and one of weight 4: 4 POPC like the full
    // reduction to weights 1/2/4/8, but 16 ALU operations instead of 22 (the ALU pipe was the longer one per distance)
    uint32_t t1, s1, t2, s2, t3, s3, f1, tw1;
    PBF_CSA(t1, s1, x0, x1, x2);
    PBF_CSA(t2, s2, x3, x4, x5);
    PBF_CSA(t3, s3, s1, s2, x6);
    PBF_CSA(f1, tw1, t1, t2, t3);
    return __popc(s3) + __popc(x7) + 2 * __popc(tw1) + 4 * __popc(f1);
}

__global__ void __launch_bounds__(PBF_THREADS)
bf_pairs_kernel(const uint8_t* __restrict__ dq, const uint8_t* const* __restrict__ train_bases,
                const sfe_bf_pair* __restrict__ pairs, uint32_t* __restrict__ best)
{
    __shared__ uint4 ts[PBF_CHUNK * 2];
    const sfe_bf_pair pr = pairs[blockIdx.z];
    if ((int)(blockIdx.x * PBF_THREADS) >= pr.nq || pr.nt == 0) return;
    const int per = (pr.nt + PBF_SPLITS - 1) / PBF_SPLITS;
    const int t0 = blockIdx.y * per, t1 = min(pr.nt, t0 + per);
    if (t0 >= t1) return;
    const int q = blockIdx.x * PBF_THREADS + threadIdx.x;
    const bool active = q < pr.nq;
    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    if (active) {
        const uint4* qp = reinterpret_cast<const uint4*>(dq + (size_t)(pr.q_start + q) * 32);
        q0 = __ldg(qp); q1 = __ldg(qp + 1);
    }
    // possibly a peer pointer: plain (coherent) loads, not the read-only path
    const uint4* tp = reinterpret_cast<const uint4*>(train_bases[pr.t_rank] + (size_t)pr.t_start * 32);
    uint32_t key = 0xffffffffu;
    for (int c0 = t0; c0 < t1; c0 += PBF_CHUNK) {
        const int n = min(PBF_CHUNK, t1 - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += PBF_THREADS) ts[i] = tp[2 * (size_t)c0 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const int d = hamming256_csa(q0, q1, ts[2 * j], ts[2 * j + 1]);
            key = min(key, ((uint32_t)d << 16) | (uint32_t)(c0 + j));
        }
    }
    if (active) atomicMin(&best[pr.out_start + q], key);
}

void launch_bf_pairs(cudaStream_t s, const uint8_t* dq, const uint8_t* const* train_bases, const sfe_bf_pair* pairs,
                     int npairs, int max_nq, uint32_t* best)
{
    if (npairs <= 0 || max_nq <= 0) return;
    dim3 grid((max_nq + PBF_THREADS - 1) / PBF_THREADS, PBF_SPLITS, npairs);
    bf_pairs_kernel<<<grid, PBF_THREADS, 0, s>>>(dq, train_bases, pairs, best);
}
