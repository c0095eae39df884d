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
// One launch covers all levels and images.  A CTA stages a 134x38 footprint with one 3-D TMA load;
// border tiles patch the reflect-101 halo inside shared memory (TMA zero-fills outside the image), so the
// passes are branch-free: the fixed-point row pass is two DP4A per output on byte-permuted words, the
// column pass works on 4x4 register blocks and stores packed 32-bit words.
// HBM-bound: 1 byte read + 1 byte written per pixel.
#include "sfe_internal.cuh"

__constant__ int c_taps_i[7] = { 18, 34, 49, 55, 49, 34, 18 };
__constant__ float c_taps_f[7];

void sfe_upload_blur_taps(const float* taps) { cudaMemcpyToSymbol(c_taps_f, taps, 7 * sizeof(float)); }

// Row pass of the fixed-point mode with DP4A: 7 taps = two dot products of 4 packed bytes.
__device__ __forceinline__ int row7_dp4a(uint32_t lo4, uint32_t hi4)
{
    // taps 18 34 49 55 | 49 34 18 0, little-endian packed
    return (int)__dp4a(hi4, 0x00122231u, __dp4a(lo4, 0x37312212u, 0u));
}

template <int MODE>
__global__ void __launch_bounds__(256, MODE == SFE_BLUR_CV24_FIXED ? 8 : 6)
blur_kernel(const __grid_constant__ LevelPack P, const int4* __restrict__ tile_tab, uint8_t* __restrict__ blur)
{
    __shared__ alignas(128) uint8_t tile[BLUR_BOX_H * BLUR_BOX_W];
    __shared__ alignas(8) uint64_t bar;
    // fixed point: row sums are <= 255*257 = 65535, so two vertically adjacent rows share a 32-bit word (DP2A operand)
    __shared__ alignas(16) int rowbuf[(MODE == SFE_BLUR_CV24_FIXED ? BLUR_ROWS / 2 : BLUR_ROWS) * BLUR_TW];

    const int tid = threadIdx.x;
    const int4 ti = __ldg(tile_tab + blockIdx.x);
    const int level = ti.x, x0 = ti.y, y0 = ti.z;
    const LevelGeom& g = P.g[level];
    const int img = blockIdx.y;
    const int w = g.w, h = g.h;
    const int bx0 = x0 - 16;                      // TMA box start (16-byte aligned); tile byte 12 = pixel x0-4

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, BLUR_BOX_H * BLUR_BOX_W);
        tma_load_3d(tile, &P.pyr_map[level], bx0, y0 - 3, img, &bar);
    }
    // one warp polls the mbarrier, the others park in the hardware barrier (a spinning warp burns issue slots that the
    // co-resident CTAs need: 8.4 M of 298 M warp instructions were try_wait retries)
    if (tid < 32) mbar_wait(&bar, 0);
    __syncthreads();

    // ---- reflect-101 fix-up, border tiles only (TMA zero-filled what lies outside the image) ----------
    // tile row r <-> image row y0-3+r (r < BLUR_ROWS used); tile byte c <-> image column bx0+c
    const bool bl = x0 == 0, br = x0 + BLUR_TW + 3 > w, bt = y0 == 0, bb = y0 + BLUR_TH + 3 > h;
    if (bl || br) {                                // columns first (rows still raw)
        for (int i = tid; i < BLUR_ROWS * 8; i += 256) {
            const int r = i >> 3, k = i & 7;
            uint8_t* row = tile + r * BLUR_BOX_W;
            if (bl && k < 3) row[16 - 1 - k] = row[16 + 1 + k];                       // x = -1-k  <- x = 1+k
            if (br && k >= 4 && k < 7) { const int x = w + (k - 4); const int c = x - bx0;   // x = w+j <- w-2-j
                if (c >= 0 && c < BLUR_BOX_W) row[c] = row[min(max(2 * w - 2 - x - bx0, 0), BLUR_BOX_W - 1)]; }
        }
        __syncthreads();
    }
    if (bt || bb) {
        for (int i = tid; i < 6 * (BLUR_BOX_W / 4); i += 256) {
            const int k = i / (BLUR_BOX_W / 4), cw = i - k * (BLUR_BOX_W / 4);
            uint32_t* tw = reinterpret_cast<uint32_t*>(tile);
            if (bt && k < 3) tw[(3 - 1 - k) * (BLUR_BOX_W / 4) + cw] = tw[(3 + 1 + k) * (BLUR_BOX_W / 4) + cw];   // y = -1-k <- 1+k
            if (bb && k >= 3) {
                const int y = h + (k - 3), r = y - (y0 - 3), rs = 2 * h - 2 - y - (y0 - 3);
                if (r >= 0 && r < BLUR_ROWS && rs >= 0) tw[r * (BLUR_BOX_W / 4) + cw] = tw[rs * (BLUR_BOX_W / 4) + cw];
            }
        }
        __syncthreads();
    }

    if constexpr (MODE == SFE_BLUR_CV24_FIXED) {
        // ---- row pass, fixed point: 2 rows x 4 adjacent outputs per thread.  The 7 taps are applied to the ALIGNED words
        //      with shifted tap patterns (10 DP4A per 4 outputs, no byte permutes); the sums of rows 2p and 2p+1 are packed
        //      into one word (low | high << 16) for the column pass' DP2A ------------------------------------------------
        const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile) + 3;      // word 3 = pixels x0-4 .. x0-1
        auto row4 = [](uint32_t A, uint32_t B, uint32_t C, uint32_t (&o)[4]) {  // A|B|C = pixels x-4..x-1 | x..x+3 | x+4..x+7
            o[0] = __dp4a(B, 0x12223137u, __dp4a(A, 0x31221200u, 0u));                             // x-3 .. x+3
            o[1] = __dp4a(C, 0x00000012u, __dp4a(B, 0x22313731u, __dp4a(A, 0x22120000u, 0u)));     // x-2 .. x+4
            o[2] = __dp4a(C, 0x00001222u, __dp4a(B, 0x31373122u, __dp4a(A, 0x12000000u, 0u)));     // x-1 .. x+5
            o[3] = __dp4a(C, 0x00122231u, __dp4a(B, 0x37312212u, 0u));                             // x   .. x+6
        };
#pragma unroll
        for (int it = 0; it < ((BLUR_ROWS / 2) * (BLUR_TW / 4) + 255) / 256; ++it) {     // unrolled: tap words stay in registers
            const int i = tid + 256 * it;
            if (i >= (BLUR_ROWS / 2) * (BLUR_TW / 4)) break;
            const int pr = i >> 5, q = i & 31;
            const uint32_t* c = tw + 2 * pr * (BLUR_BOX_W / 4) + q;
            uint32_t lo[4], hi[4];
            row4(c[0], c[1], c[2], lo);
            row4(c[BLUR_BOX_W / 4], c[BLUR_BOX_W / 4 + 1], c[BLUR_BOX_W / 4 + 2], hi);
            *reinterpret_cast<uint4*>(rowbuf + pr * BLUR_TW + 4 * q) =
                make_uint4(__byte_perm(lo[0], hi[0], 0x5410), __byte_perm(lo[1], hi[1], 0x5410),
                           __byte_perm(lo[2], hi[2], 0x5410), __byte_perm(lo[3], hi[3], 0x5410));
        }
        __syncthreads();

        // ---- column pass, fixed point: 4 columns x 4 rows per thread from 5 packed row pairs; 4 DP2A per output (the
        //      accumulator starts at 2^15), min(s, 2^24-1) so that byte 2 is the saturated result, PRMT packs 4 pixels ----
        uint8_t* dst = blur + g.lvl_ofs + (size_t)img * g.img_bytes;
        {
            const int yq = (tid >> 5) * 4, xq = (tid & 31) * 4;               // (BLUR_TH/4) * (BLUR_TW/4) == 256 items
            uint4 W[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) W[k] = *reinterpret_cast<const uint4*>(rowbuf + ((yq >> 1) + k) * BLUR_TW + xq);
            const int nvalid = min(max(w - (x0 + xq), 0), 4);
            const uint32_t cmask = nvalid >= 4 ? 0xffffffffu : ((1u << (8 * nvalid)) - 1u);     // zero beyond the image width
            // taps (t_a | t_b << 8) on a pair (row 2j, row 2j+1):  even output rows  (18,34)(49,55)(49,34)(18,0)
            //                                                      odd output rows   (0,18)(34,49)(55,49)(34,18)
            auto col = [](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, bool odd) -> uint32_t {
                uint32_t s = odd ? __dp2a_lo(w3, 0x1222u, __dp2a_lo(w2, 0x3137u, __dp2a_lo(w1, 0x3122u, __dp2a_lo(w0, 0x1200u, 1u << 15))))
                                 : __dp2a_lo(w3, 0x0012u, __dp2a_lo(w2, 0x2231u, __dp2a_lo(w1, 0x3731u, __dp2a_lo(w0, 0x2212u, 1u << 15))));
                return min(s, 0x00ffffffu);
            };
#pragma unroll
            for (int yy = 0; yy < 4; ++yy) {
                if (y0 + yq + yy >= h) break;
                const int j = yy >> 1; const bool odd = yy & 1;
                const uint32_t s0 = col(W[j].x, W[j + 1].x, W[j + 2].x, W[j + 3].x, odd);
                const uint32_t s1 = col(W[j].y, W[j + 1].y, W[j + 2].y, W[j + 3].y, odd);
                const uint32_t s2 = col(W[j].z, W[j + 1].z, W[j + 2].z, W[j + 3].z, odd);
                const uint32_t s3 = col(W[j].w, W[j + 1].w, W[j + 2].w, W[j + 3].w, odd);
                const uint32_t out = __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
                *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + yq + yy) * g.pitch + x0 + xq) = out & cmask;
            }
        }
        return;
    }

    // ---- row pass: 4 adjacent outputs per thread from 3 aligned words ------------------------------------
    {
        const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile) + 3;      // word 3 = pixels x0-4 .. x0-1
        for (int i = tid; i < BLUR_ROWS * (BLUR_TW / 4); i += 256) {
            const int r = i >> 5, q = i & 31;
            const uint32_t* c = tw + r * (BLUR_BOX_W / 4) + q;
            const uint32_t A = c[0], B = c[1], C = c[2];                       // pixels x-4..x-1 | x..x+3 | x+4..x+7
            int4 o;
            if (MODE == SFE_BLUR_CV24_FIXED) {
                o.x = row7_dp4a(__byte_perm(A, B, 0x4321), __byte_perm(B, C, 0x4321));
                o.y = row7_dp4a(__byte_perm(A, B, 0x5432), __byte_perm(B, C, 0x5432));
                o.z = row7_dp4a(__byte_perm(A, B, 0x6543), __byte_perm(B, C, 0x6543));
                o.w = row7_dp4a(B, C);
            } else {
                float px[10];
                px[0] = (float)((A >> 8) & 0xff); px[1] = (float)((A >> 16) & 0xff); px[2] = (float)(A >> 24);
                px[3] = (float)(B & 0xff); px[4] = (float)((B >> 8) & 0xff); px[5] = (float)((B >> 16) & 0xff); px[6] = (float)(B >> 24);
                px[7] = (float)(C & 0xff); px[8] = (float)((C >> 8) & 0xff); px[9] = (float)((C >> 16) & 0xff);
                float acc[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float a = __fmul_rn(c_taps_f[0], px[e]);
#pragma unroll
                    for (int k = 1; k < 7; ++k) a = __fadd_rn(a, __fmul_rn(c_taps_f[k], px[e + k]));
                    acc[e] = a;
                }
                o.x = __float_as_int(acc[0]); o.y = __float_as_int(acc[1]); o.z = __float_as_int(acc[2]); o.w = __float_as_int(acc[3]);
            }
            *reinterpret_cast<int4*>(rowbuf + r * BLUR_TW + 4 * q) = o;
        }
    }
    __syncthreads();

    // ---- column pass: 4 columns x 4 rows per thread; one 32-bit store per row ---------------------------
    uint8_t* dst = blur + g.lvl_ofs + (size_t)img * g.img_bytes;
    for (int i = tid; i < (BLUR_TH / 4) * (BLUR_TW / 4); i += 256) {
        const int yq = (i >> 5) * 4, xq = (i & 31) * 4;
        int4 rv[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) rv[k] = *reinterpret_cast<const int4*>(rowbuf + (yq + k) * BLUR_TW + xq);
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
            if (y0 + yq + yy >= h) break;
            int v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#define RV(k) (e == 0 ? rv[k].x : e == 1 ? rv[k].y : e == 2 ? rv[k].z : rv[k].w)
                if (MODE == SFE_BLUR_CV24_FIXED) {
                    int s = 55 * RV(yy + 3) + 49 * (RV(yy + 2) + RV(yy + 4)) + 34 * (RV(yy + 1) + RV(yy + 5)) + 18 * (RV(yy) + RV(yy + 6));
                    s = (s + (1 << 15)) >> 16;
                    v[e] = min(s, 255);
                } else {
                    float s = __fmul_rn(c_taps_f[3], __int_as_float(RV(yy + 3)));
#pragma unroll
                    for (int j = 1; j <= 3; ++j)
                        s = __fadd_rn(s, __fmul_rn(c_taps_f[3 + j], __fadd_rn(__int_as_float(RV(yy + 3 + j)), __int_as_float(RV(yy + 3 - j)))));
                    v[e] = min(max(__float2int_rn(s), 0), 255);
                }
#undef RV
                if (x0 + xq + e >= w) v[e] = 0;
            }
            *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + yq + yy) * g.pitch + x0 + xq) =
                (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
        }
    }
}

void launch_blur(sfe_context* c, int nimg)
{
    const int ntiles = c->pack.btile_prefix[c->pack.nlevels];
    dim3 grid(ntiles, nimg);
    if (c->cfg.blur_mode == SFE_BLUR_FLOAT) blur_kernel<SFE_BLUR_FLOAT><<<grid, 256, 0, c->stream>>>(c->pack, c->d_btile_tab, c->d_blur);
    else blur_kernel<SFE_BLUR_CV24_FIXED><<<grid, 256, 0, c->stream>>>(c->pack, c->d_btile_tab, c->d_blur);
    c->launches++;
}
