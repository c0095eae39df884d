// k_describe.cu -- intensity-centroid orientation + steered BRIEF (rBRIEF) for every selected keypoint,
// and assembly of the final cv::KeyPoint-layout records.
//
// Replaces computeOrientation (IC_Angle + fastAtan2) and computeOrbDescriptor (WTA_K=2, patch 31)
// inside cv::ORB::operator() (OpenCV 2.4; reached through m_voEngine.processNewImagePair, reference
// EST.cpp:112,267).  The outputs are what voEngine.getValues hands to CStereoSLAMKF
// (reference CStereoSLAMKF.h:68-70,99-104): 28-byte keypoints and N x 32 CV_8U descriptors.
//
// One warp per keypoint slot.  Both patches are staged with 8-byte cp.async copies of 8-byte aligned spans (5-6 lanes per
// row): the 31-row UNBLURRED patch for the orientation and the 37-row BLURRED patch for the tests.  Orientation: lane =
// patch row; the row's 32 aligned bytes are masked to the radius-15 disc and reduced with DP4A (weights 1 and u), the
// integer moments are exact; OpenCV's fastAtan2 polynomial is evaluated unfused.  Descriptor: each lane evaluates 8 of
// the 256 tests (one descriptor byte) with the pattern rotated in float, unfused, round-half-even through the
// 1.5*2^23 magic add (bit-identical to cvRound of the CPU arithmetic); the biased integers index shared memory directly.
#include "sfe_internal.cuh"
#include <algorithm>

__device__ const int8_t d_brief_pattern[256 * 4] = {
#include "brief_pattern_31.inc"
};
__constant__ int c_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

#define DESC_WARPS 8
#define PATCH_R 18
#define BP_ROWS (2 * PATCH_R + 1)   // blurred patch: 37 rows, 48 bytes copied (37 + 7 bytes of alignment slack <= 44) into a
#define BP_PITCH 56                 // 14-word pitch: the random test gathers spread over 16 bank offsets
#define OP_ROWS 31                  // orientation patch: 31 rows of 40 bytes (31 + 7 <= 38)
#define OP_PITCH 40

__device__ __forceinline__ int dp4a_u8_s8(uint32_t pix, uint32_t wgt, int acc)     // unsigned pixels x signed weights
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wgt), "r"(acc));
    return d;
}
// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, so all copies of a lane are in flight at once
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// weights u = 4k + b - 15 of the four bytes of aligned word k, as packed int8
__host__ __device__ constexpr uint32_t orient_wu(int k)
{
    return ((uint32_t)(uint8_t)(int8_t)(4 * k - 15)) | ((uint32_t)(uint8_t)(int8_t)(4 * k - 14) << 8) |
           ((uint32_t)(uint8_t)(int8_t)(4 * k - 13) << 16) | ((uint32_t)(uint8_t)(int8_t)(4 * k - 12) << 24);
}

__device__ __forceinline__ float cv_fast_atan2(float y, float x)
{
    // cv::fastAtan2 (degrees); constants are float products as in mathfuncs.cpp
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, k), p3 = __fmul_rn(-0.3258083974640975f, k);
    const float p5 = __fmul_rn(0.1555786518463281f, k), p7 = __fmul_rn(-0.04432655554792128f, k);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(DESC_WARPS * 32, 6)
describe_kernel(const __grid_constant__ LevelPack P, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                const uint32_t* __restrict__ sel_xy, const float* __restrict__ sel_resp,
                const int32_t* __restrict__ sel_count,
                sfe_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int32_t* __restrict__ kp_count)
{
    __shared__ uint32_t pat_s[256];                                   // [test-in-byte j][lane] -> packed int8 (ax, ay, bx, by)
    __shared__ uint32_t disc_mask[16 * 9];                            // [|v|][aligned word k], pitch 9: conflict-free
    __shared__ alignas(16) uint8_t bpatch[DESC_WARPS][BP_ROWS * BP_PITCH];
    __shared__ alignas(16) uint8_t opatch[DESC_WARPS][OP_ROWS * OP_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.y;
    {   // pattern entry t = 8*L + j  ->  pat_s[j*32 + L] (the strided side is the cached global read)
        pat_s[tid] = reinterpret_cast<const uint32_t*>(d_brief_pattern)[8 * (tid & 31) + (tid >> 5)];
        if (tid < 128) {                                              // byte masks of the radius-15 disc: |u| <= umax[|v|]
            const int av = tid >> 3, k = tid & 7, m = c_umax[av];
            uint32_t mk = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (abs(4 * k + b - 15) <= m) mk |= 0xffu << (8 * b);
            disc_mask[av * 9 + k] = mk;
        }
    }
    __syncthreads();

    // slot -> (level, index): lanes 0..7 hold the per-level counts, an 8-lane scan gives the level boundaries
    const int slot = blockIdx.x * DESC_WARPS + warp;
    const int cnt = lane < P.nlevels ? sel_count[(size_t)img * P.nlevels + lane] : 0;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    const uint32_t below = __ballot_sync(0xffffffffu, lane < P.nlevels && slot < inc);
    if (blockIdx.x == 0 && tid == 7) kp_count[img] = inc;            // lane 7: total over all levels (SFE_MAX_LEVELS == 8)
    if (below == 0u) return;                                          // warp-uniform
    const int level = __ffs(below) - 1;
    const int idx = slot - (__shfl_sync(0xffffffffu, inc, level) - __shfl_sync(0xffffffffu, cnt, level));

    const LevelGeom& g = P.g[level];
    const size_t so = (size_t)img * P.sel_img_stride + g.sel_ofs + idx;
    const uint32_t xy = sel_xy[so];
    const int x = xy & 0xffff, y = xy >> 16;
    const size_t lofs = g.lvl_ofs + (size_t)img * g.img_bytes;
    const int pitch = g.pitch;                                        // multiple of 128: aligned spans stay inside the row

    // ---- stage both patches: 8-byte aligned spans, one 8-byte cp.async per lane and step ------------------------------
    uint8_t* op = opatch[warp];
    uint8_t* bp = bpatch[warp];
    const int xo = (x - 15) & ~7, osh = (x - 15) - xo;               // orientation patch: columns x-15 .. x+15
    const int xb = (x - PATCH_R) & ~7, bsh = (x - PATCH_R) - xb;     // blurred patch: columns x-18 .. x+18
    {
        const int sub = lane / 5, part = lane - 5 * sub;              // 6 rows x 5 pieces per step (lanes 30, 31 idle)
        const uint8_t* src = pyr + lofs + (size_t)(y - 15 + sub) * pitch + xo + 8 * part;
        uint8_t* dst = op + sub * OP_PITCH + 8 * part;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (lane < 30 && sub + 6 * k < OP_ROWS) cp_async8(dst + 6 * k * OP_PITCH, src + (size_t)(6 * k) * pitch);
    }
    {
        const int sub = lane / 6, part = lane - 6 * sub;              // 5 rows x 6 pieces per step (lanes 30, 31 idle)
        const uint8_t* src = blur + lofs + (size_t)(y - PATCH_R + sub) * pitch + xb + 8 * part;
        uint8_t* dst = bp + sub * BP_PITCH + 8 * part;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (lane < 30 && sub + 5 * k < BP_ROWS) cp_async8(dst + 5 * k * BP_PITCH, src + (size_t)(5 * k) * pitch);
    }
    cp_async_wait_all();
    __syncwarp();

    // ---- orientation: m10 = sum u*I, m01 = sum v*I over the radius-15 disc; lane = row v ----------------------------
    int m10 = 0, m01 = 0;
    if (lane < OP_ROWS) {
        const int v = lane - 15;
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(op + lane * OP_PITCH) + (osh >> 2);
        const uint32_t* mk = disc_mask + abs(v) * 9;
        const int bs = (osh & 3) * 8;
        uint32_t w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = rw[k];
        int s1 = 0, su = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t a = __funnelshift_r(w[k], w[k + 1], bs) & mk[k];      // pixels u = 4k-15 .. 4k-12 of row v
            s1 = (int)__dp4a(a, 0x01010101u, (uint32_t)s1);
            su = dp4a_u8_s8(a, orient_wu(k), su);
        }
        m10 = su; m01 = v * s1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = cv_fast_atan2((float)m01, (float)m10);

    // ---- 256 tests, 8 per lane -------------------------------------------------------------
    const float arad = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    const float ca = (float)cos((double)arad), sa = (float)sin((double)arad);
    // cvRound through the magic add: bits(v + 1.5*2^23) = 0x4B400000 + round_half_even(v); the bias is folded into the base
    const float MAGIC = 12582912.f;
    const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(bp) + PATCH_R * BP_PITCH + PATCH_R + bsh - 0x4B400000u * (BP_PITCH + 1u);
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t e = pat_s[j * 32 + lane];
        const float ax = (float)(int8_t)(e & 0xff), ay = (float)(int8_t)((e >> 8) & 0xff);
        const float bx = (float)(int8_t)((e >> 16) & 0xff), by = (float)(int8_t)(e >> 24);
        const float fx0 = __fadd_rn(__fsub_rn(__fmul_rn(ax, ca), __fmul_rn(ay, sa)), MAGIC);
        const float fy0 = __fadd_rn(__fadd_rn(__fmul_rn(ax, sa), __fmul_rn(ay, ca)), MAGIC);
        const float fx1 = __fadd_rn(__fsub_rn(__fmul_rn(bx, ca), __fmul_rn(by, sa)), MAGIC);
        const float fy1 = __fadd_rn(__fadd_rn(__fmul_rn(bx, sa), __fmul_rn(by, ca)), MAGIC);
        const uint32_t t0 = lds_u8(__float_as_uint(fy0) * BP_PITCH + __float_as_uint(fx0) + cbase);
        const uint32_t t1 = lds_u8(__float_as_uint(fy1) * BP_PITCH + __float_as_uint(fx1) + cbase);
        byte |= (uint32_t)(t0 < t1) << j;
    }
    // gather 4 bytes per word: lanes 4k..4k+3 -> word k
    uint32_t wv = byte << (8 * (lane & 3));
    wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
    wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
    const size_t ko = (size_t)img * P.kp_cap + slot;
    if ((lane & 3) == 0) reinterpret_cast<uint32_t*>(desc + ko * 32)[lane >> 2] = wv;

    if (lane == 0) {
        sfe_keypoint k;
        k.x = (float)x; k.y = (float)y;
        if (level != 0) { k.x = __fmul_rn(k.x, g.scale); k.y = __fmul_rn(k.y, g.scale); }
        k.size = __fmul_rn(31.f, g.scale);
        k.angle = angle;
        k.response = sel_resp[so];
        k.octave = level; k.class_id = -1;
        kps[ko] = k;
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// Fused variant (SFE_BLUR_CV24_FIXED, edge_threshold >= 22): the 7x7 Gaussian is evaluated INSIDE this kernel, on the one
// unblurred 43x43 patch a keypoint needs, instead of blurring every pyramid level into a second arena first.  The
// OpenCV-2.4 fixed-point blur has no intermediate rounding (row sums <= 255*257 fit 16 bits, one (s + 2^15) >> 16 at the
// end), so the separable pass over the patch gives bit for bit the bytes the whole-image blur would give at those
// positions.  What it removes: blur_kernel (read P + write P per image), the blurred arena (half the context's
// footprint), the second patch staging (43 rows instead of 31 + 37).  Per warp:
//   stage   rows y-21..y+21 of the 8-byte aligned span that holds columns x-21..x+21 (7 cp.async of 8 bytes per row)
//   orient  intensity centroid on rows y-15..y+15 of the same patch (as above)
//   rows    7-tap row pass, two rows x four outputs per work item: DP4A on the ALIGNED words with shifted tap patterns,
//           the sums of rows 2p and 2p+1 packed into one word (k_blur.cu's formulation)
//   cols    7-tap column pass: lane = (4-column group, block of 12/13 rows) sliding down its rows with a window of four
//           packed row pairs, 4 DP2A per output, saturate, pack -> the 37x40 blurred patch overwrites the raw one
//   tests   as above, on the blurred patch.
// ------------------------------------------------------------------------------------------------------------------------
#define FZ_R 21                          // raw patch radius: 18 (sampler reach) + 3 (blur)
#define FZ_ROWS 44                       // 43 rows + one spare (the last row pair is (42, 43))
#define FZ_PITCH 88                      // bytes: 8 pad | 56 data | 24 pad; 22 words: consecutive row PAIRS land 12 banks apart (row pass: lanes span 4 pairs x 10 groups)
#define FZ_PAIRS 22
#ifndef FZ_KPW
#define FZ_KPW 8                         // keypoint slots a warp describes in turn (2: 0.424, 4: 0.404, 8: 0.400, 16: 0.412 ms)
#endif
#ifndef FZ_WARPS
#define FZ_WARPS 6                       // 6 x 7.2 KB of patch buffers per CTA, four resident CTAs (24 warps) per SM.  Static grid: 2 / 3 / 4 / 5 / 6 / 10 / 15 warps measured 0.411 / 0.408 / 0.400 / 0.392 / 0.400 / 0.417 / 0.424 ms; with the work counter 4 / 5 / 6 / 7 / 8 / 12 warps 0.404 / 0.390 / 0.384 / 0.406 / 0.390 / 0.394 ms
#endif
#ifndef FZ_CTAS
#define FZ_CTAS 4
#endif
#ifndef FZ_RESIDENT
#define FZ_RESIDENT 4                    // CTAs of FZ_WARPS warps that fit an SM (shared memory)
#endif
#define FZ_GROUPS 10                     // 4-column output groups covering x-18 .. x+18 for every alignment
#define FZ_RB_PITCH (FZ_GROUPS * 4)      // row-sum words per row pair
#define FZ_WARP_BYTES (FZ_ROWS * FZ_PITCH + FZ_PAIRS * FZ_RB_PITCH * 4)

__global__ void __launch_bounds__(FZ_WARPS * 32, FZ_CTAS)
describe_fused_kernel(const __grid_constant__ LevelPack P, const uint8_t* __restrict__ pyr,
                      const uint32_t* __restrict__ sel_xy, const float* __restrict__ sel_resp,
                      const int32_t* __restrict__ sel_count,
                      sfe_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, int32_t* __restrict__ kp_count,
                      int32_t* __restrict__ work_ctr, const int nimg)
{
    extern __shared__ __align__(16) uint8_t fz_smem[];
    __shared__ uint32_t pat_s[256];
    __shared__ uint32_t disc_mask[16 * 9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < 256; t += FZ_WARPS * 32) pat_s[t] = reinterpret_cast<const uint32_t*>(d_brief_pattern)[8 * (t & 31) + (t >> 5)];
    for (int t = tid; t < 128; t += FZ_WARPS * 32) {
        const int av = t >> 3, k = t & 7, m = c_umax[av];
        uint32_t mk = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) if (abs(4 * k + b - 15) <= m) mk |= 0xffu << (8 * b);
        disc_mask[av * 9 + k] = mk;
    }
    pdl_wait();                                                                  // the selection lists are the previous kernel's output
    __syncthreads();

    // Work items = (image, chunk of FZ_KPW keypoint slots), claimed by the warps from a global counter (the select kernel of the
    // batch zeroed it): whichever warp is free takes the next item, so the grid - one resident wave - drains evenly instead of
    // ending on a partial wave of 45-us CTAs.
    const int chunks = (P.kp_cap + FZ_KPW - 1) / FZ_KPW;
    const int nitems = nimg * chunks;
    for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(work_ctr, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    const int img = item / chunks, chunk = item - img * chunks;
    // slot -> (level, index): lanes 0..7 hold the per-level counts, an 8-lane scan gives the level boundaries
    const int cnt = lane < P.nlevels ? sel_count[(size_t)img * P.nlevels + lane] : 0;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) { const int t_ = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t_; }
    if (chunk == 0 && lane == 7) kp_count[img] = inc;
#pragma unroll 1
    for (int kk = 0; kk < FZ_KPW; ++kk) {
    const int slot = chunk * FZ_KPW + kk;
    const uint32_t below = __ballot_sync(0xffffffffu, lane < P.nlevels && slot < inc);
    if (below == 0u) break;                                                      // warp-uniform; later slots are empty too
    const int level = __ffs(below) - 1;
    const int idx = slot - (__shfl_sync(0xffffffffu, inc, level) - __shfl_sync(0xffffffffu, cnt, level));

    const LevelGeom& g = P.g[level];
    const size_t so = (size_t)img * P.sel_img_stride + g.sel_ofs + idx;
    const uint32_t xy = sel_xy[so];
    const int x = xy & 0xffff, y = xy >> 16;
    const size_t lofs = g.lvl_ofs + (size_t)img * g.img_bytes;
    const int pitch = g.pitch;

    uint8_t* raw = fz_smem + (size_t)warp * FZ_WARP_BYTES;                       // [FZ_ROWS][FZ_PITCH], data at byte 8 of a row
    uint32_t* rowbuf = reinterpret_cast<uint32_t*>(raw + FZ_ROWS * FZ_PITCH);    // [FZ_PAIRS][FZ_RB_PITCH]
    const int xb = (x - FZ_R) & ~7, bsh = (x - FZ_R) - xb;                       // data byte c of a row <-> image column xb + c
    // ---- stage the raw patch: lane = (row of a group of four, 8-byte piece); 11 steps of 4 rows, immediate offsets only -----------
    {
        const int part = lane & 7, rsub = lane >> 3;                             // piece 7 does not exist: those lanes idle
        const uint8_t* src = pyr + lofs + (size_t)(y - FZ_R + rsub) * pitch + xb + 8 * part;
        uint8_t* dst = raw + rsub * FZ_PITCH + 8 + 8 * part;
        const size_t step = (size_t)4 * pitch;
        if (part < 7) {
#pragma unroll
            for (int k = 0; k < FZ_ROWS / 4; ++k) { cp_async8(dst + 4 * k * FZ_PITCH, src); src += step; }     // rows 0 .. 43 (row 43 only feeds zero-weight taps)
        }
    }
    cp_async_wait_all();
    __syncwarp();

    // ---- orientation on rows y-15 .. y+15, columns x-15 .. x+15 of the raw patch --------------------------------------------
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int v = lane - 15;
        const int ob = 8 + bsh + (FZ_R - 15);                                    // byte of column x-15 in a row
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw + (lane + FZ_R - 15) * FZ_PITCH) + (ob >> 2);
        const uint32_t* mk = disc_mask + abs(v) * 9;
        const int bs = (ob & 3) * 8;
        uint32_t w[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) w[k] = rw[k];
        int s1 = 0, su = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t a = __funnelshift_r(w[k], w[k + 1], bs) & mk[k];
            s1 = (int)__dp4a(a, 0x01010101u, (uint32_t)s1);
            su = dp4a_u8_s8(a, orient_wu(k), su);
        }
        m10 = su; m01 = v * s1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = cv_fast_atan2((float)m01, (float)m10);

    // ---- blur, row pass: item = (row pair, 4-column group) ---------------------------------------------------------------------
    const int g0 = (bsh + 3) >> 2;                                               // first group that holds column x-18
    {
        auto row4 = [](uint32_t A, uint32_t B, uint32_t C, uint32_t (&o)[4]) {    // A|B|C = pixels x-4..x-1 | x..x+3 | x+4..x+7
            o[0] = __dp4a(B, 0x12223137u, __dp4a(A, 0x31221200u, 0u));
            o[1] = __dp4a(C, 0x00000012u, __dp4a(B, 0x22313731u, __dp4a(A, 0x22120000u, 0u)));
            o[2] = __dp4a(C, 0x00001222u, __dp4a(B, 0x31373122u, __dp4a(A, 0x12000000u, 0u)));
            o[3] = __dp4a(C, 0x00122231u, __dp4a(B, 0x37312212u, 0u));
        };
        // lane = (group gq of 10, row pair of three): 8 steps of three row pairs, immediate offsets, no index arithmetic; the three
        // pairs of a step sit 12 banks apart (FZ_PITCH), so the 30 active lanes read conflict-free
        const int prl = lane / FZ_GROUPS, gq = lane - prl * FZ_GROUPS;
        const uint32_t* c = reinterpret_cast<const uint32_t*>(raw) + 2 + g0 + 2 * prl * (FZ_PITCH / 4) + gq;   // data starts at word 2 of a row
        uint32_t* ob = rowbuf + prl * FZ_RB_PITCH + 4 * gq;
        if (lane < 3 * FZ_GROUPS) {
#pragma unroll
            for (int it = 0; it < (FZ_PAIRS + 2) / 3; ++it) {
                if (3 * it + 2 >= FZ_PAIRS && 3 * it + prl >= FZ_PAIRS) break;   // last step: only pair 21 exists
                const uint32_t* cc = c + 6 * it * (FZ_PITCH / 4);
                uint32_t lo[4], hi[4];
                row4(cc[-1], cc[0], cc[1], lo);
                row4(cc[FZ_PITCH / 4 - 1], cc[FZ_PITCH / 4], cc[FZ_PITCH / 4 + 1], hi);
                *reinterpret_cast<uint4*>(ob + 3 * it * FZ_RB_PITCH) =
                    make_uint4(__byte_perm(lo[0], hi[0], 0x5410), __byte_perm(lo[1], hi[1], 0x5410),
                               __byte_perm(lo[2], hi[2], 0x5410), __byte_perm(lo[3], hi[3], 0x5410));
            }
        }
    }
    __syncwarp();
    // ---- blur, column pass: lane = (group, row block); output row yy <-> image row y-18+yy reads raw rows yy .. yy+6 -------------------
    uint8_t* bp = raw;                                                           // the blurred patch [37][BP_PITCH] overwrites the raw one
    if (lane < 3 * FZ_GROUPS) {
        const int rb = lane / FZ_GROUPS, gq = lane - rb * FZ_GROUPS;
        const int r0 = 12 * rb;                                                  // row blocks [0,12) [12,24) [24,37): even starts
        auto col = [](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, bool odd) -> uint32_t {
            uint32_t s = odd ? __dp2a_lo(w3, 0x1222u, __dp2a_lo(w2, 0x3137u, __dp2a_lo(w1, 0x3122u, __dp2a_lo(w0, 0x1200u, 1u << 15))))
                             : __dp2a_lo(w3, 0x0012u, __dp2a_lo(w2, 0x2231u, __dp2a_lo(w1, 0x3731u, __dp2a_lo(w0, 0x2212u, 1u << 15))));
            return min(s, 0x00ffffffu);
        };
        const uint4* rp = reinterpret_cast<const uint4*>(rowbuf + 4 * gq) + (r0 >> 1) * (FZ_RB_PITCH / 4);
        uint32_t* outp = reinterpret_cast<uint32_t*>(bp + r0 * BP_PITCH + 4 * gq);
        uint4 W[10];                                                             // row pairs j0 .. j0+9: 7 steps x 2 rows, fully unrolled (no window moves)
#pragma unroll
        for (int k = 0; k < 3; ++k) W[k] = rp[k * (FZ_RB_PITCH / 4)];
#pragma unroll
        for (int it = 0; it < 7; ++it) {
            if (it < 6 || rb == 2) W[it + 3] = rp[(it + 3) * (FZ_RB_PITCH / 4)];   // the 7th step only exists for the 13-row block
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                const int yy = 2 * it + o;                                       // row within the block
                if (yy >= 12 && (rb != 2 || yy > 12)) continue;
                const uint32_t s0 = col(W[it].x, W[it + 1].x, W[it + 2].x, W[it + 3].x, o), s1 = col(W[it].y, W[it + 1].y, W[it + 2].y, W[it + 3].y, o);
                const uint32_t s2 = col(W[it].z, W[it + 1].z, W[it + 2].z, W[it + 3].z, o), s3 = col(W[it].w, W[it + 1].w, W[it + 2].w, W[it + 3].w, o);
                outp[yy * (BP_PITCH / 4)] = __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);   // nobody reads the raw patch any more
            }
        }
    }
    __syncwarp();

    // ---- 256 tests, 8 per lane, on the blurred patch: byte (yy, 4*gq + e) <-> pixel (xb + 4*(g0 + gq) + e, y - 18 + yy) ---------------
    const float arad = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    double dsa, dca; sincos((double)arad, &dsa, &dca);                           // one range reduction for both (same values as sin / cos)
    const float ca = (float)dca, sa = (float)dsa;
    const float MAGIC = 12582912.f;
    const uint32_t cbase = (uint32_t)__cvta_generic_to_shared(bp) + PATCH_R * BP_PITCH + (bsh + FZ_R - 4 * g0) - 0x4B400000u * (BP_PITCH + 1u);
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t e = pat_s[j * 32 + lane];
        const float ax = (float)(int8_t)(e & 0xff), ay = (float)(int8_t)((e >> 8) & 0xff);
        const float bx = (float)(int8_t)((e >> 16) & 0xff), by = (float)(int8_t)(e >> 24);
        const float fx0 = __fadd_rn(__fsub_rn(__fmul_rn(ax, ca), __fmul_rn(ay, sa)), MAGIC);
        const float fy0 = __fadd_rn(__fadd_rn(__fmul_rn(ax, sa), __fmul_rn(ay, ca)), MAGIC);
        const float fx1 = __fadd_rn(__fsub_rn(__fmul_rn(bx, ca), __fmul_rn(by, sa)), MAGIC);
        const float fy1 = __fadd_rn(__fadd_rn(__fmul_rn(bx, sa), __fmul_rn(by, ca)), MAGIC);
        const uint32_t t0 = lds_u8(__float_as_uint(fy0) * BP_PITCH + __float_as_uint(fx0) + cbase);
        const uint32_t t1 = lds_u8(__float_as_uint(fy1) * BP_PITCH + __float_as_uint(fx1) + cbase);
        byte |= (uint32_t)(t0 < t1) << j;
    }
    uint32_t wv = byte << (8 * (lane & 3));
    wv |= __shfl_xor_sync(0xffffffffu, wv, 1);
    wv |= __shfl_xor_sync(0xffffffffu, wv, 2);
    const size_t ko = (size_t)img * P.kp_cap + slot;
    if ((lane & 3) == 0) reinterpret_cast<uint32_t*>(desc + ko * 32)[lane >> 2] = wv;

    if (lane == 0) {
        sfe_keypoint k;
        k.x = (float)x; k.y = (float)y;
        if (level != 0) { k.x = __fmul_rn(k.x, g.scale); k.y = __fmul_rn(k.y, g.scale); }
        k.size = __fmul_rn(31.f, g.scale);
        k.angle = angle;
        k.response = sel_resp[so];
        k.octave = level; k.class_id = -1;
        kps[ko] = k;
    }
    __syncwarp();                                                                // the next keypoint's staging overwrites the patch buffers
    }
    }
}

bool describe_is_fused(const sfe_context* c)
{
    return c->cfg.blur_mode == SFE_BLUR_CV24_FIXED && c->cfg.edge_threshold >= FZ_R + 1 && !c->force_unfused;
}

void prepare_describe(sfe_context* c)
{
    sfe_raise_dynamic_smem((const void*)describe_fused_kernel, c->cfg.device, FZ_WARPS * FZ_WARP_BYTES);
}

static_assert(FZ_ROWS * FZ_PITCH >= BP_ROWS * BP_PITCH, "the blurred patch must fit over the raw one");
void launch_describe(sfe_context* c, int nimg)
{
    if (describe_is_fused(c)) {
        const int chunks = (c->pack.kp_cap + FZ_KPW - 1) / FZ_KPW;
        dim3 grid(std::min((nimg * chunks + FZ_WARPS - 1) / FZ_WARPS, 148 * FZ_RESIDENT));     // one resident wave of persistent CTAs
        sfe_launch(c, (c->pdl_mask & 8) != 0, describe_fused_kernel, grid, dim3(FZ_WARPS * 32), FZ_WARPS * FZ_WARP_BYTES, c->pack, (const uint8_t*)c->d_pyr, (const uint32_t*)c->d_sel_xy,
                   (const float*)c->d_sel_resp, (const int32_t*)c->d_sel_count, c->d_kp, c->d_desc, c->d_kp_count, c->d_desc_ctr, nimg);
        return;
    }
    dim3 grid((c->pack.kp_cap + DESC_WARPS - 1) / DESC_WARPS, nimg);
    describe_kernel<<<grid, DESC_WARPS * 32, 0, c->stream>>>(c->pack, c->d_pyr, c->d_blur, c->d_sel_xy, c->d_sel_resp,
                                                             c->d_sel_count, c->d_kp, c->d_desc, c->d_kp_count);
    c->launches++;
}
