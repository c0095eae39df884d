// k_fast.cu -- FAST-9/16 corner score + 3x3 strict non-max suppression over every pyramid level.
//
// Replaces cv::FastFeatureDetector(threshold, true) run per level inside cv::ORB (OpenCV 2.4),
// reached through m_voEngine.processNewImagePair (reference EST.cpp:112,267); the threshold is the
// run-time value of setFASTThreshold (EST.cpp:277,302,310).  Semantics (oracle
// orc_fast9_nms): s(p) = max over the 16 arcs of 9 contiguous ring pixels of min(|I(q)-I(p)|) with
// one sign; corner iff s > th; score = s-1; scores are 0 outside the 3-px margin; keep iff the
// score is strictly greater than all 8 neighbours.  Only corners inside the border-filtered ROI
// (runByImageBorder(edgeThreshold)) are emitted -- the others can never survive.
//
// One launch covers all levels and images: blockIdx.x -> (level, tile) through a prefix table,
// blockIdx.y -> image.  A CTA stages its 136x24 pixel footprint with one 3-D TMA load, then:
//   A. every thread runs the opposite-pair quick reject on its pixels and survivors are compacted
//      into a shared candidate list with warp ballots;
//   B. the compacted candidates get the exact 16-pixel test and, if corners, the arc score;
//   C. 3x3 NMS on the shared score tile; survivors are appended to the per-(image,level) raw list
//      with one warp-aggregated atomic per warp.
// Integer-only; bounded by HBM reads of the level (1 byte/pixel) plus issue slots for the tests.
#include "sfe_internal.cuh"

__device__ __forceinline__ int arc9_any(uint32_t m)      // m: 16-bit ring mask; any 9 contiguous set bits (circular)?
{
    m |= m << 16;
    uint32_t x = m & (m >> 1);      // runs of 2
    x &= x >> 2;                    // runs of 4
    x &= x >> 4;                    // runs of 8
    x &= m >> 8;                    // runs of 9
    return (x & 0xffffu) != 0;
}

__global__ void __launch_bounds__(256)
fast_kernel(const __grid_constant__ LevelPack P, uint32_t* __restrict__ raw, int32_t* __restrict__ raw_count,
            const int th)
{
    __shared__ alignas(128) uint8_t tile[FAST_BOX_H * FAST_BOX_W];
    __shared__ alignas(8) uint64_t bar;
    __shared__ uint8_t score[FAST_SH * FAST_SW];
    __shared__ uint16_t cand[FAST_SH * 130];
    __shared__ int ncand;

    const int tid = threadIdx.x, lane = tid & 31;
    int level = 0;
#pragma unroll
    for (int l = 1; l < SFE_MAX_LEVELS; ++l)
        if (l < P.nlevels && (int)blockIdx.x >= P.ftile_prefix[l]) level = l;
    const LevelGeom& g = P.g[level];
    const int t = blockIdx.x - P.ftile_prefix[level];
    const int tx0 = g.rx0 + (t % g.ftiles_x) * FAST_TW;
    const int ty0 = g.ry0 + (t / g.ftiles_x) * FAST_TH;
    const int img = blockIdx.y;
    const int bx0 = (tx0 - 4) & ~15;              // TMA: box start must be 16-byte aligned in x
    const int xoff = tx0 - 4 - bx0;               // tile column of pixel tx0-4

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); ncand = 0; }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, FAST_BOX_H * FAST_BOX_W);
        tma_load_3d(tile, &P.pyr_map[level], bx0, ty0 - 4, img, &bar);
    }
    for (int i = tid; i < FAST_SH * FAST_SW / 4; i += 256) reinterpret_cast<uint32_t*>(score)[i] = 0u;
    mbar_wait(&bar, 0);

    // ---- A: quick reject on the 130x18 score positions ------------------------------------
    // score position (sx, sy) <-> pixel (tx0-1+sx, ty0-1+sy) <-> tile byte (sx+3+xoff, sy+3)
    const int w = g.w, h = g.h;
    for (int base = 0; base < FAST_SH * 130; base += 256) {
        const int pos = base + tid;
        bool pass = false;
        if (pos < FAST_SH * 130) {
            const int sy = pos / 130, sx = pos - sy * 130;
            const int px = tx0 - 1 + sx, py = ty0 - 1 + sy;
            if (px >= 3 && px < w - 3 && py >= 3 && py < h - 3) {
                const uint8_t* c = tile + (sy + 3) * FAST_BOX_W + sx + 3 + xoff;
                const int v = c[0];
                const int lo = v - th, hi = v + th;
                const int p0 = c[3 * FAST_BOX_W], p8 = c[-3 * FAST_BOX_W], p4 = c[3], p12 = c[-3];
                // every 9-arc contains one pixel of each opposite pair
                const bool br = (p0 > hi || p8 > hi) && (p4 > hi || p12 > hi);
                const bool dk = (p0 < lo || p8 < lo) && (p4 < lo || p12 < lo);
                pass = br || dk;
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&ncand, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (pass) cand[wbase + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)pos;
        }
    }
    __syncthreads();

    // ---- B: exact test + score on the compacted candidates --------------------------------
    const int nc = ncand;
    for (int i = tid; i < nc; i += 256) {
        const int pos = cand[i];
        const int sy = pos / 130, sx = pos - sy * 130;
        const uint8_t* c = tile + (sy + 3) * FAST_BOX_W + sx + 3 + xoff;
        const int v = c[0];
        int d[16];
        d[0] = c[3 * FAST_BOX_W] - v;       d[1] = c[3 * FAST_BOX_W + 1] - v;
        d[2] = c[2 * FAST_BOX_W + 2] - v;   d[3] = c[FAST_BOX_W + 3] - v;
        d[4] = c[3] - v;                    d[5] = c[-FAST_BOX_W + 3] - v;
        d[6] = c[-2 * FAST_BOX_W + 2] - v;  d[7] = c[-3 * FAST_BOX_W + 1] - v;
        d[8] = c[-3 * FAST_BOX_W] - v;      d[9] = c[-3 * FAST_BOX_W - 1] - v;
        d[10] = c[-2 * FAST_BOX_W - 2] - v; d[11] = c[-FAST_BOX_W - 3] - v;
        d[12] = c[-3] - v;                  d[13] = c[FAST_BOX_W - 3] - v;
        d[14] = c[2 * FAST_BOX_W - 2] - v;  d[15] = c[3 * FAST_BOX_W - 1] - v;
        uint32_t mb = 0, md = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { mb |= (uint32_t)(d[k] > th) << k; md |= (uint32_t)(d[k] < -th) << k; }
        const bool isb = arc9_any(mb), isd = arc9_any(md);
        if (!(isb || isd)) continue;
        // strength of the passing polarity (the other one cannot also hold a 9-arc):
        // max over arcs of the min over the arc of e[k] = +-d[k]  (sliding-window min, window 9)
        int e[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) e[k] = isb ? d[k] : -d[k];
        int m2[16], m4[16], best = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) m2[k] = min(e[k], e[(k + 1) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int m8 = min(m4[k], m4[(k + 4) & 15]);
            best = max(best, min(m8, e[(k + 8) & 15]));
        }
        score[sy * FAST_SW + sx] = (uint8_t)(best - 1);
    }
    __syncthreads();

    // ---- C: 3x3 strict NMS + emit ----------------------------------------------------------
    int32_t* cnt = raw_count + (size_t)img * P.nlevels + level;
    uint32_t* out = raw + (size_t)img * P.raw_img_stride + g.raw_ofs;
    for (int base = 0; base < FAST_TW * FAST_TH; base += 256) {
        const int pos = base + tid;
        const int oy = pos >> 7, ox = pos & 127;
        const int px = tx0 + ox, py = ty0 + oy;
        const uint8_t* s = score + (oy + 1) * FAST_SW + ox + 1;
        const int sc = s[0];
        bool keep = false;
        if (sc != 0 && px < g.rx1 && py < g.ry1) {
            keep = sc > s[-1] && sc > s[1] && sc > s[-FAST_SW - 1] && sc > s[-FAST_SW] && sc > s[-FAST_SW + 1] &&
                   sc > s[FAST_SW - 1] && sc > s[FAST_SW] && sc > s[FAST_SW + 1];
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (bal) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(cnt, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (keep) {
                const int idx = wbase + __popc(bal & ((1u << lane) - 1u));
                if (idx < g.raw_cap) out[idx] = ((uint32_t)sc << 24) | ((uint32_t)py << 12) | (uint32_t)px;
            }
        }
    }
}

void launch_fast(sfe_context* c, int nimg)
{
    const int ntiles = c->pack.ftile_prefix[c->pack.nlevels];
    if (ntiles == 0) return;
    dim3 grid(ntiles, nimg);
    fast_kernel<<<grid, 256, 0, c->stream>>>(c->pack, c->d_raw, c->d_raw_count, c->fast_th);
    c->launches++;
}
