// k_describe.cu -- intensity-centroid orientation + steered BRIEF (rBRIEF) for every selected keypoint,
// and assembly of the final cv::KeyPoint-layout records.
//
// Replaces computeOrientation (IC_Angle + fastAtan2) and computeOrbDescriptor (WTA_K=2, patch 31)
// inside cv::ORB::operator() (OpenCV 2.4; reached through m_voEngine.processNewImagePair, reference
// EST.cpp:112,267).  The outputs are what voEngine.getValues hands to CStereoSLAMKF
// (reference CStereoSLAMKF.h:68-70,99-104): 28-byte keypoints and N x 32 CV_8U descriptors.
//
// One warp per keypoint slot.  Orientation: lanes = patch columns, 31 coalesced row reads from the
// UNBLURRED level, integer moments reduced with warp shuffles (exact), then OpenCV's fastAtan2
// polynomial evaluated unfused.  Descriptor: the 37x37 BLURRED patch is staged in shared memory,
// each lane evaluates 8 of the 256 tests (one descriptor byte) with the pattern rotated in float,
// unfused, round-half-even -- bit-identical to the CPU arithmetic.
#include "sfe_internal.cuh"

__device__ const int8_t d_brief_pattern[256 * 4] = {
#include "brief_pattern_31.inc"
};
__constant__ int c_umax[16] = { 15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3 };

#define DESC_WARPS 8
#define PATCH_R 18
#define PATCH_PITCH 44      // 11 words: the 37-byte row plus up to 3 bytes of alignment slack on the left

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
    __shared__ uint32_t pat_s[256];                                   // [test-in-byte j][lane] -> char4
    __shared__ alignas(16) uint8_t patch_s[DESC_WARPS][(2 * PATCH_R + 1) * PATCH_PITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.y;
    {   // pattern entry t = 8*L + j  ->  pat_s[j*32 + L]
        const uint32_t e = reinterpret_cast<const uint32_t*>(d_brief_pattern)[tid];
        pat_s[(tid & 7) * 32 + (tid >> 3)] = e;
    }
    __syncthreads();

    // slot -> (level, index)
    const int slot = blockIdx.x * DESC_WARPS + warp;
    int level = -1, idx = 0, cum = 0;
    for (int l = 0; l < P.nlevels; ++l) {
        const int c = sel_count[(size_t)img * P.nlevels + l];
        if (level < 0 && slot < cum + c) { level = l; idx = slot - cum; }
        cum += c;
    }
    if (blockIdx.x == 0 && tid == 0) kp_count[img] = cum;
    if (level < 0) return;

    const LevelGeom& g = P.g[level];
    const size_t so = (size_t)img * P.sel_img_stride + g.sel_ofs + idx;
    const uint32_t xy = sel_xy[so];
    const int x = xy & 0xffff, y = xy >> 16;
    const size_t lofs = g.lvl_ofs + (size_t)img * g.img_bytes;
    const uint8_t* im = pyr + lofs;
    const uint8_t* bl = blur + lofs;
    const int pitch = g.pitch;

    // ---- orientation: m10 = sum u*I, m01 = sum v*I over the radius-15 disc ------------------
    // lane = column u; the disc is symmetric (umax[umax[v]] >= v), so column u spans rows |v| <= umax[|u|]
    int m10 = 0, m01 = 0;
    {
        const int u = lane - 15;
        const int vm = lane < 31 ? c_umax[abs(u)] : -1;
        const uint8_t* p = im + (size_t)(y - 15) * pitch + x + u;
        int sum = 0;
#pragma unroll
        for (int v = -15; v <= 15; ++v) {
            if (abs(v) <= vm) { const int I = __ldg(p); sum += I; m01 += v * I; }
            p += pitch;
        }
        m10 = u * sum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
    }
    const float angle = cv_fast_atan2((float)m01, (float)m10);

    // ---- stage the blurred 37x37 patch with aligned 32-bit loads: 37 rows x 11 words ----------------
    uint8_t* ps = patch_s[warp];
    const int xa = (x - PATCH_R) & ~3, xsh = (x - PATCH_R) - xa;
    {
        const uint8_t* b0 = bl + (size_t)(y - PATCH_R) * pitch + xa;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const int idx = lane + 32 * k;
            if (idx < (2 * PATCH_R + 1) * 11) {
                const int row = idx / 11, col = idx - row * 11;
                reinterpret_cast<uint32_t*>(ps)[row * (PATCH_PITCH / 4) + col] =
                    __ldg(reinterpret_cast<const uint32_t*>(b0 + (size_t)row * pitch) + col);
            }
        }
    }
    __syncwarp();

    // ---- 256 tests, 8 per lane -------------------------------------------------------------
    const float arad = __fmul_rn(angle, (float)(3.14159265358979323846 / 180.f));
    const float ca = (float)cos((double)arad), sa = (float)sin((double)arad);
    const uint8_t* ctr = ps + PATCH_R * PATCH_PITCH + PATCH_R + xsh;
    uint32_t byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t e = pat_s[j * 32 + lane];
        const float ax = (float)(int8_t)(e & 0xff), ay = (float)(int8_t)((e >> 8) & 0xff);
        const float bx = (float)(int8_t)((e >> 16) & 0xff), by = (float)(int8_t)(e >> 24);
        const int x0 = __float2int_rn(__fsub_rn(__fmul_rn(ax, ca), __fmul_rn(ay, sa)));
        const int y0 = __float2int_rn(__fadd_rn(__fmul_rn(ax, sa), __fmul_rn(ay, ca)));
        const int x1 = __float2int_rn(__fsub_rn(__fmul_rn(bx, ca), __fmul_rn(by, sa)));
        const int y1 = __float2int_rn(__fadd_rn(__fmul_rn(bx, sa), __fmul_rn(by, ca)));
        const int t0 = ctr[y0 * PATCH_PITCH + x0], t1 = ctr[y1 * PATCH_PITCH + x1];
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
