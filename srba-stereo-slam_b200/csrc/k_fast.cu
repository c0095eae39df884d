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
// One launch covers all levels and images: blockIdx.x -> (level, tile) through a tile table,
// blockIdx.y -> image.  A CTA (128x32 keypoint tile) stages its 142x40 pixel footprint with one 3-D TMA load, then:
//   A.  SWAR quick reject, 4 pixels per thread (VABSDIFF4 + byte-parallel compares) on the opposite
//       pairs (0,8) and (4,12); survivors are compacted into a shared list (warp scan + one atomic);
//   B.  exact corner strength of the survivors, both polarities at once in packed s16x2 (VIMNMX.S16x2
//       sliding-window min over the ring); corners get their score and are compacted again;
//   C.  3x3 strict NMS evaluated only at the corners; survivors are appended to the per-(image,level)
//       raw list with one warp-aggregated atomic per warp.
// Integer-only; bounded by HBM reads of the level (1 byte/pixel) plus issue slots for the tests.
#include "sfe_internal.cuh"

// SWAR helpers on 4 packed bytes; flags are returned in bit 7 of each byte.
__device__ __forceinline__ uint32_t gt80_const(uint32_t x, uint32_t k7f_minus_t)   // x > t per byte, t < 128
{
    return (((x & 0x7f7f7f7fu) + k7f_minus_t) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t nibble_of_flags(uint32_t f80)                  // bit 7 of byte i -> bit i
{
    return (f80 * 0x00204081u) >> 28;
}

#define FAST_WJ 34                       // score words per row: pixel columns tx0-4 .. tx0+131
#define FAST_NPOS (FAST_SH * FAST_WJ * 4)
#define FAST_TP (FAST_BOX_W / 4)         // tile pitch in words
#define FAST_AK ((FAST_SH * FAST_WJ + 255) / 256)   // score words per thread in phase A (<= 8: 4 flag bits each)

__global__ void __launch_bounds__(256, 6)
fast_kernel(const __grid_constant__ LevelPack P, const int4* __restrict__ tile_tab, uint32_t* __restrict__ raw,
            int32_t* __restrict__ raw_count, const int th)
{
    __shared__ alignas(128) uint8_t tile[FAST_BOX_H * FAST_BOX_W];
    __shared__ alignas(8) uint64_t bar;
    __shared__ alignas(16) uint8_t score[FAST_SH * FAST_WJ * 4];
    __shared__ uint16_t list1[FAST_NPOS];       // quick-test survivors: (row*34 + word)*4 + byte
    __shared__ uint16_t list2[FAST_NPOS];       // corners: same | polarity << 15
    __shared__ int n1, n2;

    const int tid = threadIdx.x, lane = tid & 31;
    const int4 ti = __ldg(tile_tab + blockIdx.x);
    const int level = ti.x, tx0 = ti.y, ty0 = ti.z;                 // tx0 is a multiple of 4: pixel words are smem words
    const LevelGeom& g = P.g[level];
    const int img = blockIdx.y;
    const int bx0 = (tx0 - 8) & ~15;              // TMA: box start must be 16-byte aligned in x
    const int xoff = tx0 - 8 - bx0;               // 0, 4, 8 or 12: tile byte column of pixel tx0-8

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); n1 = 0; n2 = 0; }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, FAST_BOX_H * FAST_BOX_W);
        tma_load_3d(tile, &P.pyr_map[level], bx0, ty0 - 4, img, &bar);
    }
    for (int i = tid; i < FAST_SH * FAST_WJ; i += 256) reinterpret_cast<uint32_t*>(score)[i] = 0u;
    mbar_wait(&bar, 0);

    // score word (j, r) <-> pixels (tx0-4+4j .. +3, ty0-1+r) <-> tile word (xoff/4 + 1 + j, r+3)
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile) + (xoff >> 2) + 1;
    const uint8_t* tb = tile + xoff + 4;

    // ---- A: SWAR quick reject, 4 pixels per thread: every 9-arc holds one pixel of each opposite pair,
    //         so a corner needs |p-v| > th on (0 or 8) and on (4 or 12) -------------------------------
    const uint32_t kq = (uint32_t)(0x7f - th) * 0x01010101u;
    {
        uint32_t m[FAST_AK];
#pragma unroll
        for (int k = 0; k < FAST_AK; ++k) {
            const int idx = tid + 256 * k;
            m[k] = 0;
            if (idx < FAST_SH * FAST_WJ) {
                const int r = idx / FAST_WJ, j = idx - r * FAST_WJ;
                if (tx0 - 4 + 4 * j <= g.rx1 && ty0 - 1 + r <= g.ry1) {
                    const uint32_t* c = tw + (r + 3) * FAST_TP + j;
                    const uint32_t v = c[0];
                    const uint32_t a0 = __vabsdiffu4(c[3 * FAST_TP], v), a8 = __vabsdiffu4(c[-3 * FAST_TP], v);
                    const uint32_t a4 = __vabsdiffu4(__byte_perm(v, c[1], 0x6543), v);
                    const uint32_t a12 = __vabsdiffu4(__byte_perm(c[-1], v, 0x4321), v);
                    m[k] = (gt80_const(a0, kq) | gt80_const(a8, kq)) & (gt80_const(a4, kq) | gt80_const(a12, kq));
                }
            }
        }
        // compaction: the byte flags of this thread fold into one bit mask (4 bits per word); one shared atomic per thread
        // that has survivors (ptxas aggregates same-address atomics across the warp), then a short bit loop
        uint32_t cm = 0;
#pragma unroll
        for (int k = 0; k < FAST_AK; ++k) cm |= nibble_of_flags(m[k]) << (4 * k);
        if (cm) {
            int o = atomicAdd(&n1, __popc(cm));
            do {
                const int b = __ffs(cm) - 1;
                cm &= cm - 1;
                list1[o++] = (uint16_t)((tid + 256 * (b >> 2)) * 4 + (b & 3));
            } while (cm);
        }
    }
    __syncthreads();

    // ---- B: exact corner strength of the survivors, both polarities at once in packed s16x2 ---------------------
    // P[k] = (ring_k - v | v - ring_k << 16); strength = max over the 16 arcs of the min over 9 contiguous P
    // (sliding-window min by doubling: 4 x 16 VIMNMX.S16x2), corner iff max(lo, hi) > th, score = strength - 1.
    const int nc1 = n1;
    for (int base = (tid >> 5) * 32; base < nc1; base += 256) {        // warp-strided: idle warps skip the loop
        const int i = base + lane;
        bool corner = false; int pos = 0;
        if (i < nc1) {
            pos = list1[i];
            const int wi = pos >> 2, r = wi / FAST_WJ, j = wi - r * FAST_WJ;
            const uint8_t* c = tb + (r + 3) * FAST_BOX_W + 4 * j + (pos & 3);
            const uint32_t v = c[0];
            const uint32_t v2 = v * 0xFFFF0001u + 0x00FF0000u;          // (v | (255 - v) << 16)
            uint32_t P[16];
#define RINGP(k, off) P[k] = __vsub2((uint32_t)c[off] * 0xFFFF0001u + 0x00FF0000u, v2)
            RINGP(0, 3 * FAST_BOX_W);       RINGP(1, 3 * FAST_BOX_W + 1);   RINGP(2, 2 * FAST_BOX_W + 2);   RINGP(3, FAST_BOX_W + 3);
            RINGP(4, 3);                    RINGP(5, -FAST_BOX_W + 3);      RINGP(6, -2 * FAST_BOX_W + 2);  RINGP(7, -3 * FAST_BOX_W + 1);
            RINGP(8, -3 * FAST_BOX_W);      RINGP(9, -3 * FAST_BOX_W - 1);  RINGP(10, -2 * FAST_BOX_W - 2); RINGP(11, -FAST_BOX_W - 3);
            RINGP(12, -3);                  RINGP(13, FAST_BOX_W - 3);      RINGP(14, 2 * FAST_BOX_W - 2);  RINGP(15, 3 * FAST_BOX_W - 1);
#undef RINGP
            // window-9 min as 3 x 3 with three-input packed min (VIMNMX3.S16x2): 16 + 16 ops, then a max3 tree
            uint32_t m3[16], best;
#pragma unroll
            for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(P[k], P[(k + 1) & 15], P[(k + 2) & 15]);
            uint32_t m9[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) m9[k] = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
            uint32_t t5[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) t5[k] = __vimax3_s16x2(m9[3 * k], m9[3 * k + 1], m9[3 * k + 2]);
            best = __vimax3_s16x2(__vimax3_s16x2(t5[0], t5[1], t5[2]), __vimax3_s16x2(t5[3], t5[4], m9[15]), 0x80008000u);
            const int strength = max((int)(short)(best & 0xffffu), (int)(short)(best >> 16));
            corner = strength > th;
            if (corner) score[pos] = (uint8_t)(strength - 1);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, corner);
        if (bal) {
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&n2, __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (corner) list2[wbase + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)pos;
        }
    }
    __syncthreads();
    const int nc2 = n2;

    // ---- C: 3x3 strict NMS, only at the corners, + emit ----------------------------------------------------
    int32_t* cnt = raw_count + (size_t)img * P.nlevels + level;
    uint32_t* out = raw + (size_t)img * P.raw_img_stride + g.raw_ofs;
    const int SPB = FAST_WJ * 4;
    const int y_end = min(ty0 + FAST_TH, g.ry1);
    for (int base = (tid >> 5) * 32; base < nc2; base += 256) {        // warp-strided: idle warps skip the loop
        const int i = base + lane;
        bool keep = false; int px = 0, py = 0, sc = 0;
        if (i < nc2) {
            const int pos = list2[i] & 0x7fff;
            const int r = pos / SPB, cb = pos - r * SPB;
            px = tx0 - 4 + cb; py = ty0 - 1 + r;
            if (px >= max(tx0, g.rx0) && px < min(tx0 + FAST_TW, g.rx1) && py >= ty0 && py < y_end) {
                const uint8_t* s = score + pos;
                sc = s[0];
                keep = sc > s[-1] && sc > s[1] && sc > s[-SPB - 1] && sc > s[-SPB] && sc > s[-SPB + 1] &&
                       sc > s[SPB - 1] && sc > s[SPB] && sc > s[SPB + 1];
            }
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
    fast_kernel<<<grid, 256, 0, c->stream>>>(c->pack, c->d_ftile_tab, c->d_raw, c->d_raw_count, c->fast_th);
    c->launches++;
}
