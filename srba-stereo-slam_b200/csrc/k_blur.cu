// k_blur.cu -- 7x7 sigma=2 Gaussian blur of every pyramid level (feeds the rBRIEF sampler only).
//
// Replaces GaussianBlur(level, Size(7,7), 2, 2, BORDER_REFLECT_101) inside cv::ORB::operator()
// (OpenCV 2.4, after detection; reached through m_voEngine.processNewImagePair, reference
// EST.cpp:112,267).  Two arithmetic modes (sfe_config.blur_mode):
//   SFE_BLUR_CV24_FIXED  taps round(256*g) = 18 34 49 55 49 34 18, int32 row pass, column pass
//                        (sum + 2^15) >> 16, saturated -- the OpenCV-2.4 8-bit path the reference links.
//   SFE_BLUR_FLOAT       float taps, float row pass accumulated left to right, symmetric float column
//                        pass, round-half-even -- cv::sepFilter2D since OpenCV 3.4 (what cv2 4.13's ORB
//                        runs; pinned bit-exactly by tests/golden).
// One launch covers all levels and images.  A CTA stages a 134x22 footprint with one 3-D TMA load;
// the reflect-101 border is resolved by index remapping inside the tile (TMA zero-fills outside).
// HBM-bound: 1 byte read + 1 byte written per pixel.
#include "sfe_internal.cuh"

__constant__ int c_taps_i[7] = { 18, 34, 49, 55, 49, 34, 18 };
__constant__ float c_taps_f[7];

void sfe_upload_blur_taps(const float* taps) { cudaMemcpyToSymbol(c_taps_f, taps, 7 * sizeof(float)); }

template <int MODE>
__global__ void __launch_bounds__(256)
blur_kernel(const __grid_constant__ LevelPack P, uint8_t* __restrict__ blur)
{
    __shared__ alignas(128) uint8_t tile[BLUR_BOX_H * BLUR_BOX_W];
    __shared__ alignas(8) uint64_t bar;
    __shared__ int rowbuf[BLUR_BOX_H * BLUR_TW];

    const int tid = threadIdx.x;
    int level = 0;
#pragma unroll
    for (int l = 1; l < SFE_MAX_LEVELS; ++l)
        if (l < P.nlevels && (int)blockIdx.x >= P.btile_prefix[l]) level = l;
    const LevelGeom& g = P.g[level];
    const int t = blockIdx.x - P.btile_prefix[level];
    const int x0 = (t % g.btiles_x) * BLUR_TW, y0 = (t / g.btiles_x) * BLUR_TH;
    const int img = blockIdx.y;
    const int w = g.w, h = g.h;
    const int bx0 = (x0 - 3) & ~15;               // TMA: box start must be 16-byte aligned in x

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, BLUR_BOX_H * BLUR_BOX_W);
        tma_load_3d(tile, &P.pyr_map[level], bx0, y0 - 3, img, &bar);
    }
    mbar_wait(&bar, 0);

    // ---- row pass: rowbuf[r][x] for the 22 tile rows ---------------------------------------
    for (int i = tid; i < BLUR_BOX_H * BLUR_TW; i += 256) {
        const int r = i >> 7, x = i & 127;
        int iy = y0 - 3 + r;                               // image row of this tile row (reflect-101)
        iy = iy < 0 ? -iy : (iy >= h ? 2 * h - 2 - iy : iy);
        const int tr = min(max(iy - (y0 - 3), 0), BLUR_BOX_H - 1);   // only rows feeding y >= h can clamp
        const uint8_t* row = tile + tr * BLUR_BOX_W;
        int acc_i = 0; float acc_f = 0.f;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            int ix = x0 + x - 3 + k;
            ix = ix < 0 ? -ix : (ix >= w ? 2 * w - 2 - ix : ix);
            int tx = ix - bx0;
            tx = min(max(tx, 0), BLUR_BOX_W - 1);          // only padding columns (x >= w) can clamp
            const int v = row[tx];
            if (MODE == SFE_BLUR_CV24_FIXED) acc_i += c_taps_i[k] * v;
            else acc_f = (k == 0) ? __fmul_rn(c_taps_f[0], (float)v) : __fadd_rn(acc_f, __fmul_rn(c_taps_f[k], (float)v));
        }
        rowbuf[i] = (MODE == SFE_BLUR_CV24_FIXED) ? acc_i : __float_as_int(acc_f);
    }
    __syncthreads();

    // ---- column pass: 4 adjacent pixels per thread, one 32-bit store ------------------------
    uint8_t* dst = blur + g.lvl_ofs + (size_t)img * g.img_bytes;
    for (int i = tid; i < BLUR_TH * BLUR_TW / 4; i += 256) {
        const int y = i >> 5, xq = (i & 31) * 4;
        if (y0 + y >= h) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int x = xq + e;
            int v;
            if (MODE == SFE_BLUR_CV24_FIXED) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k) s += c_taps_i[k] * rowbuf[(y + k) * BLUR_TW + x];
                s = (s + (1 << 15)) >> 16;
                v = min(s, 255);
            } else {
                float s = __fmul_rn(c_taps_f[3], __int_as_float(rowbuf[(y + 3) * BLUR_TW + x]));
#pragma unroll
                for (int j = 1; j <= 3; ++j) {
                    const float pr = __fadd_rn(__int_as_float(rowbuf[(y + 3 + j) * BLUR_TW + x]),
                                               __int_as_float(rowbuf[(y + 3 - j) * BLUR_TW + x]));
                    s = __fadd_rn(s, __fmul_rn(c_taps_f[3 + j], pr));
                }
                v = min(max(__float2int_rn(s), 0), 255);
            }
            if (x0 + x >= w) v = 0;
            packed |= (uint32_t)v << (8 * e);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + y) * g.pitch + x0 + xq) = packed;
    }
}

void launch_blur(sfe_context* c, int nimg)
{
    const int ntiles = c->pack.btile_prefix[c->pack.nlevels];
    dim3 grid(ntiles, nimg);
    if (c->cfg.blur_mode == SFE_BLUR_FLOAT) blur_kernel<SFE_BLUR_FLOAT><<<grid, 256, 0, c->stream>>>(c->pack, c->d_blur);
    else blur_kernel<SFE_BLUR_CV24_FIXED><<<grid, 256, 0, c->stream>>>(c->pack, c->d_blur);
    c->launches++;
}
