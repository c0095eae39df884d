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
//   A.  SWAR quick reject, 4 pixels per thread-word (VABSDIFF4 + byte-parallel compares) on the opposite
//       pairs (0,8) and (4,12); a thread owns one word column of 4 adjacent rows (10 loads feed 4 tests, no
//       branches); survivors are compacted into a shared list (one atomic per warp + straight-line predicated stores);
//   B.  exact corner strength of the survivors, both polarities at once in packed s16x2 (one IMAD per ring
//       pixel, VIMNMX3.S16x2 window min); corners get their score and go to a warp-private list (no atomics);
//   C.  3x3 strict NMS evaluated only at the corners; survivors are appended to the per-(image,level)
//       raw list with one atomic per warp.
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
#define FAST_TP (FAST_BOX_W / 4)         // tile pitch in words; the score array uses the same pitch
#define FAST_L2W (((FAST_NPOS + 255) / 256) * 32)   // a warp tests at most this many survivors

// global atomic add issued by ONE lane under a predicate (no branch)
__device__ __forceinline__ int lane0_atom_add_global(int* p, int v, int lane)
{
    int old = 0;
    asm volatile("{ .reg .pred q; setp.eq.s32 q, %3, 0; @q atom.global.add.s32 %0, [%1], %2; }"
                 : "+r"(old) : "l"(p), "r"(v), "r"(lane) : "memory");
    return __shfl_sync(0xffffffffu, old, 0);
}

#ifndef FAST_OCC
#define FAST_OCC 7                       // 32 registers, 30.9 KB of shared memory: seven CTAs per SM (measured against 6: -0.8 %)
#endif
__global__ void __launch_bounds__(256, FAST_OCC)
fast_kernel(const __grid_constant__ LevelPack P, const int4* __restrict__ tile_tab, uint32_t* __restrict__ raw,
            int32_t* __restrict__ raw_count, const int th)
{
    __shared__ alignas(128) uint8_t tile[FAST_BOX_H * FAST_BOX_W];
    __shared__ alignas(8) uint64_t bar;
    __shared__ alignas(16) uint8_t score[FAST_SH * FAST_BOX_W];   // same pitch as the tile: one offset serves both
    __shared__ uint16_t list1[FAST_NPOS];       // quick-test survivors: row * FAST_BOX_W + byte column
    __shared__ uint16_t list2[8][FAST_L2W];     // corners, same encoding; one private segment per warp (no atomics)
    __shared__ int n1;
    __shared__ alignas(16) int s_nc2[8];        // corners found by each warp

    const int tid = threadIdx.x, lane = tid & 31;
    const int4 ti = __ldg(tile_tab + blockIdx.x);
    const int level = ti.x, tx0 = ti.y, ty0 = ti.z;                 // tx0 is a multiple of 4: pixel words are smem words
    const LevelGeom& g = P.g[level];
    const int img = blockIdx.y;
    const int bx0 = (tx0 - 8) & ~15;              // TMA: box start must be 16-byte aligned in x
    const int xoff = tx0 - 8 - bx0;               // 0, 4, 8 or 12: tile byte column of pixel tx0-8

    // thread 0 initialises the mbarrier and issues the TMA at once; the CTA barrier that publishes the initialisation comes
    // after the per-thread set-up below, so the load is already in flight while the warps gather (the barrier used to sit
    // BEFORE the issue: 5 % of the warp samples waited there)
    if (tid == 0) {
        mbar_init(&bar, 1); fence_mbar_init(); n1 = 0;
        pdl_wait();                                               // the pyramid is the previous kernels' output
        mbar_arrive_expect_tx(&bar, FAST_BOX_H * FAST_BOX_W);
        tma_load_3d(tile, &P.pyr_map[level], bx0, ty0 - 4, img, &bar);
    }

    // score byte (r, col) <-> pixel (tx0-4+col, ty0-1+r) <-> tile byte (xoff + 4 + col, r+3); col = 4j + b, word j in 0..33
    const uint32_t* tw = reinterpret_cast<const uint32_t*>(tile) + (xoff >> 2) + 1;
    const uint8_t* tb = tile + xoff + 4 + 3 * FAST_BOX_W;
    uint32_t* sw = reinterpret_cast<uint32_t*>(score);

    // Work split of phase A (no divisions, immediate offsets): warp w takes the 32 interior words j = 1..32 of rows
    // 4w .. 4w+3 (bits 0..15 of the thread's flag mask); the remainder - rows 32 and 33 (threads 0..63) and the two halo
    // words of every row (threads 64..131) - is a fifth, partial pass (bits 16..19).
    const int warp = tid >> 5;
    const int rlim = g.ry1 - (ty0 - 1);                          // rows r <= rlim lie inside the ROI (+1)
    const bool colok = tx0 + 4 * lane <= g.rx1;
    int r4 = 32 + warp, j4 = 1 + lane;
    if (tid >= 64) { const int h = tid - 64; r4 = h >> 1; j4 = (h & 1) * (FAST_WJ - 1); }
    const bool in4 = tid < 2 * FAST_SH + 64;
    const bool ok4 = in4 && tx0 - 4 + 4 * j4 <= g.rx1 && r4 <= rlim;
    const int pos0 = 4 * warp * FAST_BOX_W + 4 * (1 + lane), pos4 = r4 * FAST_BOX_W + 4 * j4;
    // every score word is cleared by the thread that tests it (phase B writes corner scores after the barrier)
#pragma unroll
    for (int k = 0; k < 4; ++k) sw[(pos0 >> 2) + k * FAST_TP] = 0u;
    if (in4) sw[pos4 >> 2] = 0u;
    pdl_wait();                                                   // every thread, before the tile and the raw lists
    __syncthreads();                                              // mbarrier initialised, n1 = 0
    mbar_wait(&bar, 0);

    // ---- A: SWAR quick reject, 4 pixels per thread: every 9-arc holds one pixel of each opposite pair,
    //         so a corner needs |p-v| > th on (0 or 8) and on (4 or 12) -------------------------------
    const uint32_t kq = (uint32_t)(0x7f - th) * 0x01010101u;
    {
        auto flags = [&](uint32_t v, uint32_t up3, uint32_t dn3, uint32_t left, uint32_t right) -> uint32_t {
            const uint32_t a0 = __vabsdiffu4(dn3, v), a8 = __vabsdiffu4(up3, v);
            const uint32_t a4 = __vabsdiffu4(__byte_perm(v, right, 0x6543), v);
            const uint32_t a12 = __vabsdiffu4(__byte_perm(left, v, 0x4321), v);
            return nibble_of_flags((gt80_const(a0, kq) | gt80_const(a8, kq)) & (gt80_const(a4, kq) | gt80_const(a12, kq)));
        };
        auto quick = [&](const uint32_t* c) -> uint32_t { return flags(c[0], c[-3 * FAST_TP], c[3 * FAST_TP], c[-1], c[1]); };
        // the four rows of a thread are tested unconditionally (the box always holds them) and masked afterwards: no
        // branches, and the centre words of rows 0 and 3 double as the +-3 neighbours of rows 3 and 0
        uint32_t cm;
        {
            const uint32_t* c0 = tw + (4 * warp + 3) * FAST_TP + 1 + lane;
            uint32_t col[10];                                            // tile rows 4w .. 4w+9 of this word column
#pragma unroll
            for (int k = 0; k < 10; ++k) col[k] = c0[(k - 3) * FAST_TP];
            cm = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                cm |= flags(col[k + 3], col[k], col[k + 6], c0[k * FAST_TP - 1], c0[k * FAST_TP + 1]) << (4 * k);
            const int nrows = min(max(rlim - 4 * warp + 1, 0), 4);      // rows 4w .. 4w+nrows-1 lie inside the ROI (+1)
            cm = colok ? (cm & ((1u << (4 * nrows)) - 1u)) : 0u;
        }
        // halo words: NMS only reads the pixel column next to the tile (byte 3 of word 0, byte 0 of word 33)
        if (ok4) cm |= (quick(tw + (r4 + 3) * FAST_TP + j4) & (tid < 64 ? 0xfu : (j4 == 0 ? 0x8u : 0x1u))) << 16;
        // compaction: one shared atomic per thread that has survivors (ptxas aggregates same-address atomics across the warp: one
        // scan + one ATOMS), then the 20 flag bits are written by straight-line predicated stores through a running shared
        // ADDRESS: per bit one predicate, one value add, @p STS, @p address += 2 (inline PTX: the C++ form `list1[o++]`
        // recomputed the list base on the uniform datapath for every bit, ~9 instructions per bit; 0.479 -> 0.461 ms).
        // Measured dead ends: a data-dependent ffs/clear loop (trip count = the warp's worst lane, survivors cluster:
        // 0.498 ms), a row-major warp-scan compaction (fewer bank conflicts in phase B, 46 more instructions per warp),
        // skipping the quick test for warps whose rows lie below the ROI (0.479 -> 0.490 ms: the branch costs more than the
        // 5 % of rows it saves).  Not built, by estimate only: one item per thread + a separate expansion pass (the
        // expansion's trip count is again the warp's worst lane: ~66 M instead of ~72 M warp instructions, plus a barrier).
        if (cm) {
            uint32_t a = smem_u32(list1) + 2u * (uint32_t)atomicAdd(&n1, __popc(cm));
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const uint16_t v = (uint16_t)(pos0 + (b >> 2) * FAST_BOX_W + (b & 3));
                asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u16 [%0], %1; @q add.u32 %0, %0, 2; }"
                             : "+r"(a) : "h"(v), "r"(cm & (1u << b)) : "memory");
            }
            if (cm >> 16) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint16_t v = (uint16_t)(pos4 + b);
                    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.shared.u16 [%0], %1; @q add.u32 %0, %0, 2; }"
                                 : "+r"(a) : "h"(v), "r"(cm & (0x10000u << b)) : "memory");
                }
            }
        }
    }
    __syncthreads();

    // ---- B: exact corner strength of the survivors, both polarities at once in packed s16x2 ---------------------
    // P[k] = (256 + ring_k - v | 256 + v - ring_k << 16); strength = max over the 16 arcs of the min over 9 contiguous P
    // (window 9 = 3 x 3 with three-input VIMNMX3.S16x2) - 256, corner iff max(lo, hi) > th, score = strength - 1.
    const int nc1 = n1;
    int nc2 = 0;                                                  // warp-uniform: corners found by this warp
    for (int base = warp * 32; base < nc1; base += 256) {        // warp-strided: idle warps skip the loop
        const int i = base + lane;
        bool corner = false; int pos = 0;
        if (i < nc1) {
            pos = list1[i];
            const uint8_t* c = tb + pos;
            const uint32_t v = c[0];
            // ONE multiply-add per ring pixel: (x + 256 - v) * 0xFFFF0001 = D - (D << 16), D in [1, 511], + 0x0200 in the
            // high half.  The bias 256 keeps both halves positive.
            const uint32_t cv = (256u - v) * 0xFFFF0001u + 0x02000000u;
            uint32_t P[16];
#define RINGP(k, off) P[k] = (uint32_t)c[off] * 0xFFFF0001u + cv
            RINGP(0, 3 * FAST_BOX_W);       RINGP(1, 3 * FAST_BOX_W + 1);   RINGP(2, 2 * FAST_BOX_W + 2);   RINGP(3, FAST_BOX_W + 3);
            RINGP(4, 3);                    RINGP(5, -FAST_BOX_W + 3);      RINGP(6, -2 * FAST_BOX_W + 2);  RINGP(7, -3 * FAST_BOX_W + 1);
            RINGP(8, -3 * FAST_BOX_W);      RINGP(9, -3 * FAST_BOX_W - 1);  RINGP(10, -2 * FAST_BOX_W - 2); RINGP(11, -FAST_BOX_W - 3);
            RINGP(12, -3);                  RINGP(13, FAST_BOX_W - 3);      RINGP(14, 2 * FAST_BOX_W - 2);  RINGP(15, 3 * FAST_BOX_W - 1);
#undef RINGP
            uint32_t m3[16], best;
#pragma unroll
            for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(P[k], P[(k + 1) & 15], P[(k + 2) & 15]);
            uint32_t m9[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) m9[k] = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
            uint32_t t5[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) t5[k] = __vimax3_s16x2(m9[3 * k], m9[3 * k + 1], m9[3 * k + 2]);
            best = __vimax3_s16x2(__vimax3_s16x2(t5[0], t5[1], t5[2]), t5[3], __vimax3_s16x2(t5[4], m9[15], m9[15]));
            const int strength = (int)max(best & 0xffffu, best >> 16) - 256;
            corner = strength > th;
            if (corner) score[pos] = (uint8_t)(strength - 1);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, corner);
        if (corner) list2[warp][nc2 + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)pos;
        nc2 += __popc(bal);
    }
    if (lane == 0) s_nc2[warp] = nc2;
    __syncthreads();                                              // NMS reads scores written by other warps

    // ---- C: 3x3 strict NMS, only at the corners, + emit ----------------------------------------------------
    int32_t* cnt = raw_count + (size_t)img * P.nlevels + level;
    uint32_t* out = raw + (size_t)img * P.raw_img_stride + g.raw_ofs;
    const int SPB = FAST_BOX_W;
    const int y_end = min(ty0 + FAST_TH, g.ry1);
    // the eight private lists are walked as ONE concatenated list, 32 corners per warp iteration: full iterations instead
    // of eight partly filled ones (a tile holds ~100 corners)
    int pre[9];
    {
        const int4 c0 = *reinterpret_cast<const int4*>(s_nc2), c1 = *reinterpret_cast<const int4*>(s_nc2 + 4);
        pre[0] = 0; pre[1] = c0.x; pre[2] = pre[1] + c0.y; pre[3] = pre[2] + c0.z; pre[4] = pre[3] + c0.w;
        pre[5] = pre[4] + c1.x; pre[6] = pre[5] + c1.y; pre[7] = pre[6] + c1.z; pre[8] = pre[7] + c1.w;
    }
    const int nc_all = pre[8];
    for (int base = warp * 32; base < nc_all; base += 256) {     // warp-strided: idle warps skip the loop
        const int i = base + lane;
        bool keep = false; int px = 0, py = 0, sc = 0;
        if (i < nc_all) {
            int w = 0, start = 0;
#pragma unroll
            for (int k = 1; k < 8; ++k) if (i >= pre[k]) { w = k; start = pre[k]; }
            const int pos = list2[w][i - start];
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
            const int wbase = lane0_atom_add_global(cnt, __popc(bal), lane);
            if (keep) {
                const int idx = wbase + __popc(bal & ((1u << lane) - 1u));
                if (idx < g.raw_cap) out[idx] = ((uint32_t)sc << 24) | ((uint32_t)py << 12) | (uint32_t)px;
            }
        }
    }
}

void launch_fast(sfe_context* c, int nimg, bool after_kernel)
{
    const int ntiles = c->pack.ftile_prefix[c->pack.nlevels];
    if (ntiles == 0) return;
    dim3 grid(ntiles, nimg);
    sfe_launch(c, after_kernel && c->pack.nlevels > 1 && (c->pdl_mask & 2), fast_kernel, grid, dim3(256), 0, c->pack, (const int4*)c->d_ftile_tab, c->d_raw, c->d_raw_count, c->fast_th);
}
