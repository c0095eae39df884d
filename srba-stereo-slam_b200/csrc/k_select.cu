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
// One CTA per (image, level).  Warp-shuffle reductions for the 7x7 Harris sums (integer, exact);
// the float expression is evaluated unfused (__fmul_rn/__fsub_rn) to match the CPU bit for bit.
#include "sfe_internal.cuh"

__device__ __forceinline__ uint32_t float_order_key(float f)   // monotone map float -> uint32
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_key(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const __grid_constant__ LevelPack P, const uint8_t* __restrict__ pyr,
              const uint32_t* __restrict__ raw, const int32_t* __restrict__ raw_count,
              uint32_t* __restrict__ sel_xy, float* __restrict__ sel_resp, int32_t* __restrict__ sel_count,
              int32_t* __restrict__ status, const int cand_cap)
{
    extern __shared__ uint32_t dyn[];
    uint32_t* cand_xy = dyn;                                    // [cand_cap]  (y<<16 | x)
    float* cand_resp = reinterpret_cast<float*>(dyn + cand_cap);// [cand_cap]
    __shared__ int hist[256];
    __shared__ int s_n, s_thr, s_k;
    __shared__ uint32_t s_prefix;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int level = blockIdx.x, img = blockIdx.y;
    const LevelGeom& g = P.g[level];
    const int q = g.quota;
    const uint32_t* rl = raw + (size_t)img * P.raw_img_stride + g.raw_ofs;
    int n = raw_count[(size_t)img * P.nlevels + level];
    if (n > g.raw_cap) { n = g.raw_cap; if (tid == 0) atomicOr(&status[img], SFE_ERR_CAPACITY); }

    // ---- retainBest(2q) on the integer FAST score: histogram cut ---------------------------
    for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) atomicAdd(&hist[rl[i] >> 24], 1);
    __syncthreads();
    if (tid == 0) {
        int thr = 0;
        if (n > 2 * q) {
            int cum = 0;
            for (int s = 255; s >= 0; --s) { cum += hist[s]; if (cum >= 2 * q) { thr = s; break; } }
            if (q == 0) thr = 256;
        }
        s_thr = thr;
    }
    __syncthreads();
    const int thr = s_thr;
    for (int i = tid; i < n; i += SEL_THREADS) {
        const uint32_t r = rl[i];
        if ((int)(r >> 24) >= thr) {
            const int idx = atomicAdd(&s_n, 1);
            if (idx < cand_cap) cand_xy[idx] = (((r >> 12) & 0xfffu) << 16) | (r & 0xfffu);
        }
    }
    __syncthreads();
    int nc = s_n;
    if (nc > cand_cap) { nc = cand_cap; if (tid == 0) atomicOr(&status[img], SFE_ERR_CAPACITY); }

    // ---- Harris response, one warp per candidate ------------------------------------------
    const uint8_t* im = pyr + g.lvl_ofs + (size_t)img * g.img_bytes;
    const int pitch = g.pitch;
    for (int c = warp; c < nc; c += SEL_THREADS / 32) {
        const int x = cand_xy[c] & 0xffff, y = cand_xy[c] >> 16;
        int a = 0, b = 0, cc = 0;
        for (int k = lane; k < 49; k += 32) {
            const int i = k / 7, j = k - i * 7;
            const uint8_t* p = im + (size_t)(y - 3 + i) * pitch + (x - 3 + j);
            const int p00 = __ldg(p - pitch - 1), p01 = __ldg(p - pitch), p02 = __ldg(p - pitch + 1);
            const int p10 = __ldg(p - 1), p12 = __ldg(p + 1);
            const int p20 = __ldg(p + pitch - 1), p21 = __ldg(p + pitch), p22 = __ldg(p + pitch + 1);
            const int Ix = (p12 - p10) * 2 + (p02 - p00) + (p22 - p20);
            const int Iy = (p21 - p01) * 2 + (p20 - p00) + (p22 - p02);
            a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            cc += __shfl_xor_sync(0xffffffffu, cc, o);
        }
        if (lane == 0) {
            float scale = __fdiv_rn(1.0f, 4.0f * 7.0f * 255.0f);
            const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
            const float fa = (float)a, fb = (float)b, fc = (float)cc;
            const float t3 = __fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc));
            const float s = __fadd_rn(fa, fb);
            const float t5 = __fmul_rn(__fmul_rn(0.04f, s), s);
            cand_resp[c] = __fmul_rn(__fsub_rn(t3, t5), s4);
        }
    }
    __syncthreads();

    // ---- retainBest(q) on the float response: radix select of the q-th largest ------------
    float boundary = -__int_as_float(0x7f800000);   // -inf: keep everything
    if (nc > q && q > 0) {
        if (tid == 0) { s_prefix = 0u; s_k = q; }
        uint32_t mask = 0u;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < nc; i += SEL_THREADS) {
                const uint32_t k = float_order_key(cand_resp[i]);
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int k = s_k, cum = 0, b = 255;
                for (; b > 0; --b) { if (cum + hist[b] >= k) break; cum += hist[b]; }
                s_k = k - cum;
                s_prefix = prefix | ((uint32_t)b << shift);
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        boundary = float_from_key(s_prefix);
    } else if (q == 0) {
        nc = 0;
    }

    // ---- canonical order: rank kept points by (y, x) --------------------------------------
    uint32_t* out_xy = sel_xy + (size_t)img * P.sel_img_stride + g.sel_ofs;
    float* out_r = sel_resp + (size_t)img * P.sel_img_stride + g.sel_ofs;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int i = tid; i < nc; i += SEL_THREADS) {
        if (!(cand_resp[i] >= boundary)) continue;
        const uint32_t key = cand_xy[i];
        int rank = 0;
        for (int j = 0; j < nc; ++j) rank += (cand_resp[j] >= boundary) && (cand_xy[j] < key);
        atomicAdd(&s_n, 1);
        if (rank < g.sel_cap) { out_xy[rank] = key; out_r[rank] = cand_resp[i]; }
    }
    __syncthreads();
    if (tid == 0) {
        int nk = s_n;
        if (nk > g.sel_cap) { nk = g.sel_cap; atomicOr(&status[img], SFE_ERR_CAPACITY); }
        sel_count[(size_t)img * P.nlevels + level] = nk;
    }
}

void prepare_select(sfe_context* c)
{
    cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->sel_cand_cap * 8);
}

void launch_select(sfe_context* c, int nimg)
{
    const int cap = c->sel_cand_cap;
    const size_t smem = (size_t)cap * 8;
    dim3 grid(c->pack.nlevels, nimg);
    select_kernel<<<grid, SEL_THREADS, smem, c->stream>>>(c->pack, c->d_pyr, c->d_raw, c->d_raw_count,
                                                           c->d_sel_xy, c->d_sel_resp, c->d_sel_count,
                                                           c->d_status, cap);
    c->launches++;
}
