// sfe_internal.cuh -- shared declarations of the stereo front-end CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <utility>
#include "../../include/srba_b200.h"

#define SFE_MAX_LEVELS 8

// ---- tile geometry shared by host and device -------------------------------------------------
// Pyramid: 128x64 output tile, source box fetched by one TMA load (128x32 measured: 0.249 vs 0.237 ms per 256 images - the
// per-thread prologue is paid per tile).
#define PYR_TW 128
#ifndef PYR_TH
#define PYR_TH 64
#endif
#define PYR_BOX_W 176      // >= ceil(127*scale)+3 + 15 (16-byte aligned box start), multiple of 16
#ifndef PYR_BOX_H
#define PYR_BOX_H 88       // >= ceil((PYR_TH-1)*scale)+3, multiple of 8
#endif
#define PYR_MAX_TILES 64   // tiles per level and axis (levels >= 1: up to 8192 x 2048 pixels)
// FAST: 128x32 keypoint tile; scores on 136x34; pixels 142x40 (+ alignment slack) -> TMA box 160x40.
// TMA needs the box start 16-byte aligned in x: kernels round the start down and keep the offset.
#define FAST_TW 128
#define FAST_TH 32
#define FAST_BOX_W 160
#define FAST_BOX_H 40
#define FAST_SH (FAST_TH + 2)
// Blur: 128x32 output tile, halo 3 -> 134x38 used of a 160x40 box.
#define BLUR_TW 128
#define BLUR_TH 32
#define BLUR_BOX_W 160
#define BLUR_BOX_H 40       // shares the 160x40 maps with FAST (rows 38,39 unused)
#define BLUR_ROWS (BLUR_TH + 6)

#ifndef SEL_THREADS
#define SEL_THREADS 256
#endif
#define SEL_TIE_SLACK 64   // extra slots per level for Harris-response ties

struct LevelGeom {
    int w, h, pitch;            // pixels, bytes per row (multiple of 128)
    int quota;                  // features wanted at this level
    int sel_cap;                // quota + SEL_TIE_SLACK
    int rx0, ry0, rx1, ry1;     // keypoint ROI [rx0,rx1) x [ry0,ry1): border filter and FAST margin
    int ftiles_x, ftiles_y;     // FAST tiles over the ROI
    int btiles_x, btiles_y;     // blur tiles over the level
    int raw_cap;                // capacity of the raw corner list
    size_t img_bytes;           // pitch * h
    size_t lvl_ofs;             // byte offset of this level's [NI][h][pitch] block in the pyramid arena
    size_t raw_ofs;             // entry offset of this level in one image's raw list
    int sel_ofs;                // entry offset of this level in one image's selection list
    float scale;                // 1.2^l (float, as OpenCV getScale)
};

// Everything the all-level kernels need, passed as one __grid_constant__ parameter.
struct alignas(64) LevelPack {
    CUtensorMap pyr_map[SFE_MAX_LEVELS];    // 3-D (x, y, image) u8 maps over the pyramid levels, box 160x40
    LevelGeom g[SFE_MAX_LEVELS];
    int nlevels;
    int ftile_prefix[SFE_MAX_LEVELS + 1];   // blockIdx.x -> level for the FAST kernel
    int btile_prefix[SFE_MAX_LEVELS + 1];   // blockIdx.x -> level for the blur kernel
    size_t raw_img_stride;                  // raw entries per image
    int sel_img_stride;                     // selection entries per image
    int kp_cap;                             // keypoint capacity per image
};

// ---- PTX helpers: mbarrier + TMA -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tiled TMA load: box at (x, y, z) of the tensor map -> smem, completion on mbarrier.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The kernels of a batch form a chain in one stream / one CUDA graph.  Launched with the programmatic-serialization attribute, a
// kernel's CTAs are scheduled as soon as the previous kernel's CTAs have all exited - without waiting for that grid's completion
// and memory flush to be processed first: they run their predecessor-independent prologue (tables, shared-memory set-up) and block
// in griddepcontrol.wait until the previous grid has completed and its memory is visible.  Every kernel of the chain executes
// griddepcontrol.wait - in EVERY thread - before its first access to anything a predecessor wrote or still reads (a no-op in a
// kernel launched without the attribute).  Measured: 0.296 -> 0.291 ms per un-batched stereo pair, the seven pyramid launches of a 128-pair
// batch 0.236 -> 0.221 ms with direct launches; inside the batch's CUDA graph the step stays at 1.173 ms (within noise).  NO kernel triggers its dependents early (griddepcontrol.launch_dependents): with the trigger at the top of the kernels
// the dependents' griddepcontrol.wait returned before the primary had finished writing (select read partial FAST lists; measured,
// tools/pdl_debug.py with SFE_PDL_MASK=4) - the implicit trigger at grid completion is the only one used.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- host-side context -----------------------------------------------------------------------
enum Stage { ST_H2D = 0, ST_PYRAMID, ST_FAST, ST_SELECT, ST_BLUR, ST_DESCRIBE, ST_MATCH_LR, ST_D2H, ST_COUNT };

struct sfe_context {
    sfe_config cfg;
    int fast_th, orb_th;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2[3] = {};            // further branches of the pyramid phase (forked from / joined into `stream`, also inside the batch graph)
    cudaStream_t launch_stream = nullptr;    // where sfe_launch puts the next kernel (nullptr: `stream`)
    cudaEvent_t ev_fork = nullptr, ev_branch[3] = {};
    std::string err;
    LevelPack pack;                          // host copy (passed by value at launch)
    std::vector<std::vector<int2>> xtab, ytab;   // host resize tables: {src offset, a0 | a1<<16} per destination column / row
    int4* d_ytab[SFE_MAX_LEVELS] = {};        // per destination row: {src row 0, src row 1 (byte offsets into the kernel's row buffer), b0 << 16, b1 << 16}
    uint32_t* d_gtab[SFE_MAX_LEVELS] = {};
    int4* d_ftile_tab = nullptr;             // FAST tile -> {level, tx0, ty0, 0}
    int4* d_btile_tab = nullptr;             // blur tile -> {level, x0, y0, 0}   // per 4 destination columns: {src base, PRMT selectors, 4 coefficient pairs}
    CUtensorMap pyrsrc_map[SFE_MAX_LEVELS];  // 176x48-box maps over the same levels, source side of the pyramid kernel
    int NI = 0;                              // images the arenas are sized for (2*max_pairs)
    int pdl_mask = 0x3f;                     // debugging: which kernels of the chain may launch dependently (SFE_PDL_MASK)
    int pyr_branches = 2;                    // image ranges of the forked pyramid phase (SFE_PYR_BRANCHES=2..4)
    bool pyr_fork = true;                    // pyramid phase on two streams (SFE_NO_PYR_FORK=1: one stream; A-B timing)
    bool pdl = true;                         // programmatic dependent launch between the kernels of a batch (SFE_NO_PDL=1 turns it off: A-B timing)
    bool force_unfused = false;              // debugging / A-B timing: keep the separate blur kernel even where the fused describe applies
    int sel_cand_cap = 0;                    // select kernel: shared candidate list capacity
    bool select_attr_set = false;
    uint8_t* d_pyr = nullptr;                // pyramid arena
    uint8_t* d_blur = nullptr;               // blurred arena (same layout)
    size_t arena_bytes = 0;
    uint32_t* d_raw = nullptr;               // [NI][raw_img_stride]
    int32_t* d_raw_count = nullptr;          // [NI][L]
    uint32_t* d_sel_xy = nullptr;            // [NI][sel_img_stride]  (y<<16 | x)
    float* d_sel_resp = nullptr;             // [NI][sel_img_stride]
    int32_t* d_sel_count = nullptr;          // [NI][L]
    sfe_keypoint* d_kp = nullptr;            // [NI][kp_cap]
    uint8_t* d_desc = nullptr;               // [NI][kp_cap][32]
    int32_t* d_kp_count = nullptr;           // [NI]
    int32_t* d_desc_ctr = nullptr;           // work counter of the describe kernel (zeroed by the select kernel of the same batch)
    sfe_dmatch* d_matches = nullptr;         // [max_pairs][kp_cap]
    int32_t* d_match_count = nullptr;        // [max_pairs]
    uint32_t* d_rb_best = nullptr;           // [max_pairs][kp_cap] row-band match: packed best key per left keypoint
    uint32_t* d_rb_owner = nullptr;          // [max_pairs][kp_cap] mutual-best owner per right keypoint
    int32_t* d_status = nullptr;             // [NI] device-side error flags (capacity overflow)
    // pinned staging
    uint8_t* h_images = nullptr; size_t h_images_bytes = 0;
    uint8_t* d_packed = nullptr;             // tightly packed upload staging [NI][H][W]
    sfe_keypoint* h_kp = nullptr; uint8_t* h_desc = nullptr; sfe_dmatch* h_matches = nullptr;
    int32_t* h_kp_count = nullptr; int32_t* h_match_count = nullptr; int32_t* h_status = nullptr;
    sfe_pair_counts* h_counts = nullptr;
    int last_npairs = 0; bool have_results = false; bool downloaded = false;
    int ship_kp = 0, ship_m = 0, shipped_kp = 0, shipped_m = 0;   // result slots per image / pair the asynchronous download copies
    size_t d2h_bytes_last = 0; int64_t d2h_refetches = 0;
    // graph cache, one per batch size
    std::vector<cudaGraphExec_t> graphs;
    // timing
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_join = nullptr;
    bool profiling = false;
    cudaEvent_t ev_stage[ST_COUNT + 1] = {};
    int graph_kernels = 0;                   // kernels per graph replay
    int64_t launches = 0;
    // scratch for sfe_match_bf host entry
    uint8_t* d_bf_q = nullptr; uint8_t* d_bf_t = nullptr; uint32_t* d_bf_best = nullptr; int32_t* d_bf_ofs = nullptr;
    int bf_cap_q = 0, bf_cap_t = 0;
};

// Binds the context's device for the duration of a public entry point (and restores the caller's): two contexts on
// different GPUs in one process, or a worker thread whose current device is not the context's, must not launch on /
// allocate from the wrong device.
struct SfeDeviceGuard {
    int prev = -1; bool switched = false;
    explicit SfeDeviceGuard(int dev)
    {
        if (dev < 0 || cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
        if (prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~SfeDeviceGuard() { if (switched) cudaSetDevice(prev); }
    SfeDeviceGuard(const SfeDeviceGuard&) = delete;
    SfeDeviceGuard& operator=(const SfeDeviceGuard&) = delete;
};
#define SFE_BIND_DEVICE(c) SfeDeviceGuard sfe_dev_guard__((c) ? (c)->cfg.device : -1)

#define SFE_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return SFE_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

// Launch on the context's stream; `dependent` = the kernel may start under the previous KERNEL's tail (it calls pdl_wait()).
template <typename... KArgs, typename... Args>
inline void sfe_launch(sfe_context* c, bool dependent, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = c->launch_stream ? c->launch_stream : c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (dependent && c->pdl) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
    c->launches++;
}

// Raises (never lowers) a kernel's dynamic shared memory limit on `device`; defined in sfe_api.cu.
void sfe_raise_dynamic_smem(const void* func, int device, int bytes);

// kernel launchers (defined in the per-stage .cu files)
void launch_pyramid_level(sfe_context* c, int level, int img0, int nimg);      // images [img0, img0 + nimg)
void launch_fast(sfe_context* c, int nimg, bool after_kernel);   // after_kernel: the previous node of the stream is one kernel (may launch dependently)
void launch_select(sfe_context* c, int nimg);
void prepare_select(sfe_context* c);
void prepare_rowband(sfe_context* c);
void launch_blur(sfe_context* c, int nimg);
void launch_describe(sfe_context* c, int nimg);
bool describe_is_fused(const sfe_context* c);   // the 7x7 blur is evaluated inside the describe kernel (no blur launch, no blurred arena)
void prepare_describe(sfe_context* c);
void launch_match_rowband(sfe_context* c, int npairs);
void launch_unpack(sfe_context* c, int nimg);
void launch_match_bf(cudaStream_t s, const uint8_t* d_query, const int32_t* d_q_ofs, const uint8_t* d_train,
                     const int32_t* d_t_ofs, int npairs, int max_nq, uint32_t* d_best);
