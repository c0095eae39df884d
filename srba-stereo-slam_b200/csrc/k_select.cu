// k_select.cu -- per (image, level): retain-best by FAST score, Harris response, retain-best by
// Harris response, canonical ordering.
//
// Replaces, inside cv::ORB::computeKeyPoints (OpenCV 2.4; reached through
// m_voEngine.processNewImagePair, reference EST.cpp:112,267; nfeatures from EST.cpp:1142):
//   KeyPointsFilter::retainBest(2*quota)  ->  HarrisResponses(7, 0.04)  ->  retainBest(quota).
// retainBest keeps the n best PLUS every point whose response equals the n-th best;
// with integer FAST scores that is a 256-bin histogram cut, with float Harris responses a 4-pass
// radix select.  Survivors are written in canonical order (y, then x) -- the reference's own order
// is an std::nth_element artefact, so parity is defined on this order.
//
// One CTA per (image, level).  Harris: one thread per candidate, 9x9 window via aligned word loads and a
// 3-row sliding register window (integer sums, exact); the float expression is evaluated unfused
// (__fmul_rn/__fsub_rn) to match the CPU bit for bit.  The kernel is latency-bound (a chain of short phases), so the
// phases are arranged for few barriers (9 on the common path, was 55): every histogram is zeroed once at the start,
// every warp evaluates the k-th-from-top scan itself (no broadcast through shared memory, no barrier after it), and
// the canonical order comes from rank counting over the kept points (unique keys: rank = #smaller, broadcast
// 128-bit shared loads, no barriers) instead of a 36-stage bitonic sort.
#include "sfe_internal.cuh"

#define SEL_SORT_CAP 1024

__device__ __forceinline__ uint32_t float_order_key(float f)   // monotone map float -> uint32
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_key(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Warp-parallel "k-th from the top" over a 256-bin shared histogram (call with all 32 lanes of one warp):
// returns the largest bin b with sum_{i >= b} hist[i] >= k, and `above` = sum_{i > b} hist[i].
// If the total is below k, bin 0 is returned with above = total - hist[0].
__device__ __forceinline__ void kth_from_top(const int* hist, int k, int lane, int& bin, int& above)
{
    const int base = 8 * (31 - lane);                 // lane 0 owns the top 8 bins
    int h[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = hist[base + j]; s += h[j]; }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    const int excl = incl - s;
    const uint32_t bal = __ballot_sync(0xffffffffu, incl >= k && excl < k);
    const int src = bal ? __ffs(bal) - 1 : 31;
    int b = base, ab = excl + s - h[0];                // default: lowest bin of the group
    int cum = excl;
#pragma unroll
    for (int j = 7; j >= 0; --j) {
        if (cum + h[j] >= k) { b = base + j; ab = cum; break; }
        cum += h[j];
    }
    bin = __shfl_sync(0xffffffffu, b, src);
    above = __shfl_sync(0xffffffffu, ab, src);
}

#ifndef SEL_OCC
#define SEL_OCC 4                        // 64 registers: four 256-thread CTAs per SM
#endif
__global__ void __launch_bounds__(SEL_THREADS, SEL_OCC)
select_kernel(const __grid_constant__ LevelPack P, const uint8_t* __restrict__ pyr,
              const uint32_t* __restrict__ raw, const int32_t* __restrict__ raw_count,
              uint32_t* __restrict__ sel_xy, float* __restrict__ sel_resp, int32_t* __restrict__ sel_count,
              int32_t* __restrict__ status, const int cand_cap, int32_t* __restrict__ next_ctr)
{
    extern __shared__ uint32_t dyn[];
    uint32_t* cand_xy = dyn;                                    // [cand_cap]  (y<<16 | x)
    float* cand_resp = reinterpret_cast<float*>(dyn + cand_cap);// [cand_cap]
    __shared__ int hist[256];
    __shared__ int histp[4][256];                               // radix-select passes of the Harris cut
    __shared__ int s_n, s_n2, s_k;
    __shared__ uint32_t s_prefix;
    __shared__ alignas(16) uint32_t skey[SEL_SORT_CAP];
    __shared__ float sval[SEL_SORT_CAP];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTAs are issued x-fastest: every image's level 0 (the most candidates, the longest chain of phases) first, the cheap top
    // levels last - longest-processing-time order trims the tail of this latency-bound kernel (0.083 -> 0.076 ms)
    const int level = blockIdx.y, img = blockIdx.x;
    const LevelGeom& g = P.g[level];
    const int q = g.quota;
    const uint32_t* rl = raw + (size_t)img * P.raw_img_stride + g.raw_ofs;
    // shared-memory set-up first: it does not depend on the FAST kernel, whose tail this CTA may be running under
    for (int i = tid; i < 256; i += SEL_THREADS) { hist[i] = 0; histp[0][i] = 0; histp[1][i] = 0; histp[2][i] = 0; histp[3][i] = 0; }
    for (int i = tid; i < SEL_SORT_CAP; i += SEL_THREADS) skey[i] = 0xffffffffu;     // padding of the rank-counting loop
    if (tid == 0) { s_n = 0; s_n2 = 0; }
    pdl_wait();
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) *next_ctr = 0;      // the describe kernel's work counter (it starts after this grid)
    int n = raw_count[(size_t)img * P.nlevels + level];
    if (n > g.raw_cap) { n = g.raw_cap; if (tid == 0) atomicOr(&status[img], SFE_ERR_CAPACITY); }

    // ---- retainBest(2q) on the integer FAST score: histogram cut ---------------------------
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) atomicAdd(&hist[rl[i] >> 24], 1);
    __syncthreads();
    int thr = 0, kcut = -1;                                       // every warp evaluates the cut itself: identical values, no broadcast
    if (n > 2 * q) {
        int above = 0;
        kth_from_top(hist, 2 * q, lane, thr, above);
        if (q == 0) thr = 256;
        if (thr < 256 && above + hist[thr] > cand_cap) kcut = cand_cap - above;
    }
    // Ties at the cut that do not fit the shared candidate list (binary / saturated images: most corners score the
    // same): OpenCV would carry every tie into the Harris stage; here the ties are cut DETERMINISTICALLY to the ones
    // with the smallest (y, x) -- a 3-pass radix select on the inverted 24-bit position key -- instead of failing the
    // frame with an order-dependent truncation.
    uint32_t tie_max = 0x00ffffffu;
    if (kcut >= 0) {                                              // uniform, rare
        __syncthreads();                                          // every warp has read hist[]
        if (tid == 0) { s_prefix = 0u; s_k = kcut; }
        uint32_t mask = 0u;
        for (int shift = 16; shift >= 0; shift -= 8) {
            __syncthreads();
            for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < n; i += SEL_THREADS) {
                const uint32_t r = rl[i];
                const uint32_t k = 0x00ffffffu - (r & 0x00ffffffu);
                if ((int)(r >> 24) == thr && (k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
            }
            __syncthreads();
            if (warp == 0) {
                int b, above;
                kth_from_top(hist, s_k, lane, b, above);
                __syncwarp();
                if (lane == 0) { s_k = s_k - above; s_prefix = prefix | ((uint32_t)b << shift); }
            }
            mask |= 255u << shift;
        }
        __syncthreads();
        tie_max = 0x00ffffffu - s_prefix;
    }
    for (int base = 0; base < n; base += SEL_THREADS) {          // warp-aggregated compaction
        const int i = base + tid;
        uint32_t r = 0; bool keepc = false;
        if (i < n) { r = rl[i]; const int sc = (int)(r >> 24); keepc = sc > thr || (sc == thr && (r & 0x00ffffffu) <= tie_max); }
        const uint32_t bal = __ballot_sync(0xffffffffu, keepc);
        if (bal) {
            int wb = 0;
            if (lane == 0) wb = atomicAdd(&s_n, __popc(bal));
            wb = __shfl_sync(0xffffffffu, wb, 0);
            const int idx = wb + __popc(bal & ((1u << lane) - 1u));
            if (keepc && idx < cand_cap) cand_xy[idx] = (((r >> 12) & 0xfffu) << 16) | (r & 0xfffu);
        }
    }
    __syncthreads();
    int nc = s_n;
    if (nc > cand_cap) { nc = cand_cap; if (tid == 0) atomicOr(&status[img], SFE_ERR_CAPACITY); }

    // ---- Harris response, one thread per candidate: 9x9 window through a 3-row sliding register window ----
    const uint8_t* im = pyr + g.lvl_ofs + (size_t)img * g.img_bytes;
    const int pitch = g.pitch;
    for (int c = tid; c < nc; c += SEL_THREADS) {
        const int x = cand_xy[c] & 0xffff, y = cand_xy[c] >> 16;
        const uint8_t* rowp = im + (size_t)(y - 4) * pitch + (x - 4);
        const int sh = (int)((size_t)rowp & 3) * 8;                      // pitch % 4 == 0: same alignment on every row
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(rowp - (sh >> 3));
        int r0[9], r1[9], r2[9];
        auto load_row = [&](int* r, int i) {
            const uint32_t* w = wp + (size_t)i * (pitch >> 2);
            const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
            const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh);
            r[0] = a & 0xff; r[1] = (a >> 8) & 0xff; r[2] = (a >> 16) & 0xff; r[3] = a >> 24;
            r[4] = b & 0xff; r[5] = (b >> 8) & 0xff; r[6] = (b >> 16) & 0xff; r[7] = b >> 24;
            r[8] = (w2 >> sh) & 0xff;                                    // ninth byte: bits [sh, sh+8) of the third word
        };
        load_row(r0, 0); load_row(r1, 1);
        int a = 0, b = 0, cc = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            load_row(r2, i + 2);
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int Ix = (r1[j + 2] - r1[j]) * 2 + (r0[j + 2] - r0[j]) + (r2[j + 2] - r2[j]);
                const int Iy = (r2[j + 1] - r0[j + 1]) * 2 + (r2[j] - r0[j]) + (r2[j + 2] - r0[j + 2]);
                a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) { r0[j] = r1[j]; r1[j] = r2[j]; }
        }
        const float scale = __fdiv_rn(1.0f, 4.0f * 7.0f * 255.0f);
        const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
        const float fa = (float)a, fb = (float)b, fc = (float)cc;
        const float t3 = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
        const float sm = __fadd_rn(fa, fb);
        const float t5 = __fmul_rn(__fmul_rn(0.04f, sm), sm);
        cand_resp[c] = __fmul_rn(__fsub_rn(t3, t5), s4);
    }
    __syncthreads();

    // ---- retainBest(q) on the float response: radix select of the q-th largest ------------
    float boundary = -__int_as_float(0x7f800000);   // -inf: keep everything
    if (nc > q && q > 0) {
        uint32_t prefix = 0u, mask = 0u;
        int kk = q;
#pragma unroll 1
        for (int p = 0; p < 4; ++p) {                             // one barrier per pass: histp[p] was zeroed at the start
            const int shift = 24 - 8 * p;
            for (int i = tid; i < nc; i += SEL_THREADS) {
                const uint32_t k = float_order_key(cand_resp[i]);
                if ((k & mask) == prefix) atomicAdd(&histp[p][(k >> shift) & 255u], 1);
            }
            __syncthreads();
            int b, above;
            kth_from_top(histp[p], kk, lane, b, above);           // every warp: same result
            kk -= above; prefix |= (uint32_t)b << shift; mask |= 255u << shift;
        }
        boundary = float_from_key(prefix);
    } else if (q == 0) {
        nc = 0;
    }

    // ---- canonical order (y, then x): compact the kept points, then rank counting on the packed key ----------
    uint32_t* out_xy = sel_xy + (size_t)img * P.sel_img_stride + g.sel_ofs;
    float* out_r = sel_resp + (size_t)img * P.sel_img_stride + g.sel_ofs;
    // Ties at the Harris cut that do not fit the selection list (identical structures repeated over the image: every response equal):
    // as at the FAST-score cut above, the ties with the smallest (y, x) are kept -- deterministic, and the frame succeeds.
    uint32_t harris_tie_max = 0xffffffffu;
    for (int attempt = 0; attempt < 2; ++attempt) {
        for (int base = 0; base < nc; base += SEL_THREADS) {      // warp-aggregated compaction of the kept points
            const int i = base + tid;
            const bool kp = i < nc && (cand_resp[i] > boundary || (cand_resp[i] == boundary && cand_xy[i] <= harris_tie_max));
            const uint32_t bal = __ballot_sync(0xffffffffu, kp);
            if (bal) {
                int wb = 0;
                if (lane == 0) wb = atomicAdd(&s_n2, __popc(bal));
                wb = __shfl_sync(0xffffffffu, wb, 0);
                const int slot = wb + __popc(bal & ((1u << lane) - 1u));
                if (kp && slot < SEL_SORT_CAP) { skey[slot] = cand_xy[i]; sval[slot] = cand_resp[i]; }
            }
        }
        __syncthreads();
        if (s_n2 <= g.sel_cap || attempt == 1) break;              // uniform
        // overflow (rare): need = sel_cap - #(response > boundary) ties; 4-pass radix select of the need-th smallest (y<<16|x) among the ties
        __syncthreads();
        if (tid == 0) { s_k = 0; s_prefix = 0u; s_n2 = 0; }
        for (int i = tid; i < SEL_SORT_CAP; i += SEL_THREADS) skey[i] = 0xffffffffu;
        __syncthreads();
        int above = 0;
        for (int i = tid; i < nc; i += SEL_THREADS) above += cand_resp[i] > boundary;
        if (above) atomicAdd(&s_k, above);
        __syncthreads();
        const int need = max(g.sel_cap - s_k, 1);
        __syncthreads();
        if (tid == 0) s_k = need;
        uint32_t mask = 0u;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < nc; i += SEL_THREADS) {
                const uint32_t k = ~cand_xy[i];                   // inverted: the k-th from the top = the k-th smallest position
                if (cand_resp[i] == boundary && (k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
            }
            __syncthreads();
            if (warp == 0) {
                int b, ab;
                kth_from_top(hist, s_k, lane, b, ab);
                __syncwarp();
                if (lane == 0) { s_k = s_k - ab; s_prefix = prefix | ((uint32_t)b << shift); }
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        harris_tie_max = ~s_prefix;
    }
    const int nk_all = s_n2;
    if (nk_all <= SEL_SORT_CAP) {
        // keys are unique (one per pixel): rank = number of smaller keys; entries past nk_all hold 0xffffffff and never count
        for (int i = tid; i < nk_all; i += SEL_THREADS) {
            const uint32_t key = skey[i];
            int rank = 0;
            for (int j = 0; j < nk_all; j += 4) {
                const uint4 k4 = *reinterpret_cast<const uint4*>(skey + j);
                rank += (int)(k4.x < key) + (int)(k4.y < key) + (int)(k4.z < key) + (int)(k4.w < key);
            }
            if (rank < g.sel_cap) { out_xy[rank] = key; out_r[rank] = sval[i]; }
        }
    } else {
        // very large quotas: rank by counting over the candidate list (O(n^2), exact)
        for (int i = tid; i < nc; i += SEL_THREADS) {
            const uint32_t key = cand_xy[i];
            if (!(cand_resp[i] > boundary || (cand_resp[i] == boundary && key <= harris_tie_max))) continue;
            int rank = 0;
            for (int j = 0; j < nc; ++j) rank += (cand_resp[j] > boundary || (cand_resp[j] == boundary && cand_xy[j] <= harris_tie_max)) && (cand_xy[j] < key);
            if (rank < g.sel_cap) { out_xy[rank] = key; out_r[rank] = cand_resp[i]; }
        }
    }
    if (tid == 0) {
        int nk = nk_all;
        if (nk > g.sel_cap) { nk = g.sel_cap; atomicOr(&status[img], SFE_ERR_CAPACITY); }
        sel_count[(size_t)img * P.nlevels + level] = nk;
    }
}

void prepare_select(sfe_context* c)
{
    sfe_raise_dynamic_smem((const void*)select_kernel, c->cfg.device, c->sel_cand_cap * 8);   // per function, never lowered
}

void launch_select(sfe_context* c, int nimg)
{
    const int cap = c->sel_cand_cap;
    const size_t smem = (size_t)cap * 8;
    dim3 grid(nimg, c->pack.nlevels);
    sfe_launch(c, (c->pdl_mask & 4) != 0, select_kernel, grid, dim3(SEL_THREADS), smem, c->pack, (const uint8_t*)c->d_pyr, (const uint32_t*)c->d_raw, (const int32_t*)c->d_raw_count,
               c->d_sel_xy, c->d_sel_resp, c->d_sel_count, c->d_status, cap, c->d_desc_ctr);
}
