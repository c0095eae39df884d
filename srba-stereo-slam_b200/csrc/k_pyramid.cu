// k_pyramid.cu -- one pyramid level: 8-bit bilinear resize in OpenCV's fixed-point arithmetic.
//
// Replaces the pyramid construction inside cv::ORB::operator() reached through
// m_voEngine.processNewImagePair (reference src/CSRBAStereoSLAMEstimator.cpp:112,267; nlevels from :1141).
// Arithmetic: horizontal pass S[sx]*a0 + S[sx+1]*a1 with 11-bit coefficients into
// int32, vertical pass (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.  The coefficient tables
// are computed on the host exactly as OpenCV does (double/float) so the kernel is pure integer.
//
// One CTA = one 128x16 destination tile of one image.  Its source footprint (<= 176x24 u8 incl. alignment slack) is
// fetched with a single 3-D TMA load (out-of-bounds rows/columns arrive as zeros and are never
// weighted).  HBM-bound by design: per destination pixel ~1.44 source bytes read + 1 byte written.
#include "sfe_internal.cuh"

struct PyrArgs {
    uint8_t* dst; int dpitch; size_t dimg_bytes;
    int dw, dh, sh;
    const int2* xtab; const int2* ytab;
};

__global__ void __launch_bounds__(256)
pyramid_kernel(const __grid_constant__ CUtensorMap src_map, const PyrArgs a)
{
    __shared__ alignas(128) uint8_t tile[PYR_BOX_H * PYR_BOX_W];
    __shared__ alignas(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * PYR_TW, y0 = blockIdx.y * PYR_TH, img = blockIdx.z;
    const int sx0 = a.xtab[x0].x;
    const int bx0 = sx0 & ~15;                    // TMA: box start must be 16-byte aligned in x
    const int sy0 = max(a.ytab[y0].x, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, PYR_BOX_H * PYR_BOX_W);
        tma_load_3d(tile, &src_map, bx0, sy0, img, &bar);
    }
    // while the tile is in flight: this thread's 8 destination columns and its row
    const int tx = (tid & 15) * 8, ty = tid >> 4;
    const int dy = y0 + ty;
    int2 xt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xt[i] = a.xtab[min(x0 + tx + i, a.dw - 1)];
    const int2 yt = a.ytab[min(dy, a.dh - 1)];
    mbar_wait(&bar, 0);
    if (dy >= a.dh) return;
    const int r0 = min(max(yt.x, 0), a.sh - 1) - sy0;
    const int r1 = min(max(yt.x + 1, 0), a.sh - 1) - sy0;
    const int b0 = (short)(yt.y & 0xffff), b1 = (short)(yt.y >> 16);
    const uint8_t* S0 = tile + r0 * PYR_BOX_W;
    const uint8_t* S1 = tile + r1 * PYR_BOX_W;
    uint32_t out[2] = { 0u, 0u };
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int sx = xt[i].x - bx0;
        const int a0 = (short)(xt[i].y & 0xffff), a1 = (short)(xt[i].y >> 16);
        const int h0 = S0[sx] * a0 + S0[sx + 1] * a1;
        const int h1 = S1[sx] * a0 + S1[sx + 1] * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        if (x0 + tx + i >= a.dw) v = 0;             // keep the pitch padding deterministic
        out[i >> 2] |= (uint32_t)(v & 0xff) << (8 * (i & 3));
    }
    uint8_t* D = a.dst + (size_t)img * a.dimg_bytes + (size_t)dy * a.dpitch + x0 + tx;
    *reinterpret_cast<uint2*>(D) = make_uint2(out[0], out[1]);
}

void launch_pyramid_level(sfe_context* c, int level, int nimg)
{
    const LevelGeom& g = c->pack.g[level];
    PyrArgs a;
    a.dst = c->d_pyr + g.lvl_ofs; a.dpitch = g.pitch; a.dimg_bytes = g.img_bytes;
    a.dw = g.w; a.dh = g.h; a.sh = c->pack.g[level - 1].h;
    a.xtab = c->d_xtab[level]; a.ytab = c->d_ytab[level];
    dim3 grid((g.w + PYR_TW - 1) / PYR_TW, (g.h + PYR_TH - 1) / PYR_TH, nimg);
    pyramid_kernel<<<grid, 256, 0, c->stream>>>(c->pyrsrc_map[level - 1], a);
    c->launches++;
}

// ---- level 0 ingest: tightly packed host layout [img][H][W] -> pitched level-0 arena -----------------------
// One linear H2D copy followed by this kernel replaces a row-by-row pitched DMA (W = 1241 is not a multiple of
// 16, which the TMA descriptors of the consumers require of the row stride).
__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ dst, int w, int h, int pitch, size_t dimg_bytes)
{
    const int y = blockIdx.y, img = blockIdx.z;
    const int x = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x >= pitch) return;
    const uint8_t* src = packed + ((size_t)img * h + y) * w;
    uint32_t v = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) if (x + e < w) v |= (uint32_t)__ldg(src + x + e) << (8 * e);
    *reinterpret_cast<uint32_t*>(dst + (size_t)img * dimg_bytes + (size_t)y * pitch + x) = v;
}

void launch_unpack(sfe_context* c, int nimg)
{
    const LevelGeom& g = c->pack.g[0];
    dim3 grid((g.pitch / 4 + 255) / 256, g.h, nimg);
    unpack_kernel<<<grid, 256, 0, c->stream>>>(c->d_packed, c->d_pyr + g.lvl_ofs, g.w, g.h, g.pitch, g.img_bytes);
    c->launches++;
}
