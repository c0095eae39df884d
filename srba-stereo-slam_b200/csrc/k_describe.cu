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

void launch_describe(sfe_context* c, int nimg)
{
    dim3 grid((c->pack.kp_cap + DESC_WARPS - 1) / DESC_WARPS, nimg);
    describe_kernel<<<grid, DESC_WARPS * 32, 0, c->stream>>>(c->pack, c->d_pyr, c->d_blur, c->d_sel_xy, c->d_sel_resp,
                                                             c->d_sel_count, c->d_kp, c->d_desc, c->d_kp_count);
    c->launches++;
}
