// sfe_api.cu -- C-ABI entry points (include/srba_b200.h), context set-up, batch orchestration.
//
// Host-side geometry follows OpenCV 2.4 cv::ORB::operator() exactly (level sizes, per-level quotas,
// resize coefficient tables) so that the kernels stay pure integer; see SURVEY.md A.1 and the per-kernel
// headers for the reference call sites (EST.cpp:112,267,1135-1142).
#include "sfe_internal.cuh"
#include <cmath>
#include <cstdlib>
#include <cfloat>
#include <cstring>
#include <algorithm>

void sfe_upload_blur_taps(const float* taps);
void launch_bf_init(cudaStream_t s, uint32_t* best, int n);
void launch_bf_init_ofs(cudaStream_t s, uint32_t* best, const int32_t* d_q_ofs, int npairs);
void launch_bf_pairs(cudaStream_t s, const uint8_t* dq, const uint8_t* const* train_bases, const sfe_bf_pair* pairs,
                     int npairs, int max_nq, uint32_t* best);

static std::string g_create_err;

#include <map>
#include <mutex>
void sfe_raise_dynamic_smem(const void* func, int device, int bytes)
{
    static std::mutex mu; static std::map<std::pair<const void*, int>, int> cur;
    std::lock_guard<std::mutex> lk(mu);
    int& have = cur[std::make_pair(func, device)];
    if (bytes > have) { cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); have = bytes; }
}

static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }

extern "C" const char* sfe_version(void) { return "srba-stereo-slam_b200 0.1 (sm_100a)"; }

extern "C" void sfe_default_config(sfe_config* c)
{
    memset(c, 0, sizeof(*c));
    c->width = 1241; c->height = 376; c->nlevels = 8; c->nfeatures = 1000;
    c->scale_factor = 1.2f; c->edge_threshold = 31; c->fast_threshold = 20;
    c->blur_mode = SFE_BLUR_CV24_FIXED; c->max_pairs = 1;
    c->max_y_diff = 2.0f; c->min_disparity = 0.0f; c->max_disparity = 1e9f;
    c->orb_max_distance = 60; c->robust_1to1 = 0; c->device = 0; c->use_graph = 1; c->pyramid_mode = SFE_PYRAMID_CV24_LINEAR;
}

extern "C" const char* sfe_last_error(const sfe_context* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" void* sfe_stream(sfe_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- resize tables (OpenCV resize(), INTER_LINEAR, 8-bit): {offset, a0 | a1<<16} -----------------
static void build_resize_table(int ssize, int dsize, bool is_x, std::vector<int2>& tab)
{
    tab.resize(dsize);
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_d(f);
        f -= s;
        if (is_x) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        const float c0 = 1.f - f, c1 = f;
        const int a0 = cv_round_d(c0 * 2048), a1 = cv_round_d(c1 * 2048);
        tab[d].x = s;
        tab[d].y = (a0 & 0xffff) | (a1 << 16);
    }
}

// INTER_LINEAR_EXACT (OpenCV >= 3.4): 8-bit fractions, edge samples taken with full weight.
static void build_resize_table_exact(int ssize, int dsize, std::vector<int2>& tab)
{
    tab.resize(dsize);
    const double inv = (double)dsize / ssize, scale = 1.0 / inv;
    for (int v = 0; v < dsize; ++v) {
        const double f = scale * ((double)v + 0.5) - 0.5;
        const int i = cv_floor_d(f);
        int ofs = 0, c1 = 0;
        if (i >= 0 && ssize > 1) {
            if (i < ssize - 1) { ofs = i; c1 = cv_round_d((f - i) * 256); }
            else ofs = ssize - 1;
        }
        tab[v].x = ofs;
        tab[v].y = ((256 - c1) & 0xffff) | (c1 << 16);
    }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_image_map(sfe_context* c, PFN_encodeTiled enc, CUtensorMap* map, void* base, int w, int h, int pitch,
                          int nimg, int box_w, int box_h)
{
    cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)nimg };
    cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)h };
    cuuint32_t box[3] = { (cuuint32_t)box_w, (cuuint32_t)box_h, 1u };
    cuuint32_t estr[3] = { 1u, 1u, 1u };
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { c->err = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return SFE_ERR_CUDA; }
    return SFE_OK;
}

template <typename T> static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }
template <typename T> static cudaError_t halloc(T** p, size_t n) { return cudaMallocHost((void**)p, n * sizeof(T)); }

static int create_impl(sfe_context* c)
{
    const sfe_config& f = c->cfg;
    if (f.width < 64 || f.height < 64 || f.width > 4095 || f.height > 4095) { c->err = "width/height must be in [64,4095]"; return SFE_ERR_INVALID; }
    if (f.nlevels < 1 || f.nlevels > SFE_MAX_LEVELS) { c->err = "nlevels must be in [1,8]"; return SFE_ERR_INVALID; }
    if (f.nfeatures < 1 || f.nfeatures > 30000) { c->err = "nfeatures must be in [1,30000]"; return SFE_ERR_INVALID; }
    if (!(f.scale_factor >= 1.05f && f.scale_factor <= 1.25f)) { c->err = "scale_factor must be in [1.05,1.25] (tile footprint)"; return SFE_ERR_INVALID; }
    if (f.edge_threshold < 19 || f.edge_threshold > 200) { c->err = "edge_threshold must be in [19,200] (rBRIEF reach)"; return SFE_ERR_INVALID; }
    if (f.max_pairs < 1 || f.max_pairs > 16384) { c->err = "max_pairs must be in [1,16384]"; return SFE_ERR_INVALID; }
    if (f.fast_threshold < 1 || f.fast_threshold > 127) { c->err = "fast_threshold must be in [1,127]"; return SFE_ERR_INVALID; }
    if (f.blur_mode != SFE_BLUR_CV24_FIXED && f.blur_mode != SFE_BLUR_FLOAT) { c->err = "bad blur_mode"; return SFE_ERR_INVALID; }
    if (f.pyramid_mode != SFE_PYRAMID_CV24_LINEAR && f.pyramid_mode != SFE_PYRAMID_CV4_EXACT) { c->err = "bad pyramid_mode"; return SFE_ERR_INVALID; }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= f.device) { c->err = "no usable CUDA device (there is no CPU fallback)"; return SFE_ERR_CUDA; }
    SFE_CUDA(c, cudaSetDevice(f.device));
    cudaDeviceProp prop;
    SFE_CUDA(c, cudaGetDeviceProperties(&prop, f.device));
    if (prop.major != 10) { c->err = std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library is built for sm_100a only"; return SFE_ERR_CUDA; }
    SFE_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SFE_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int b = 0; b < 3; ++b) {
        SFE_CUDA(c, cudaStreamCreateWithFlags(&c->stream2[b], cudaStreamNonBlocking));
        SFE_CUDA(c, cudaEventCreateWithFlags(&c->ev_branch[b], cudaEventDisableTiming));
    }
    SFE_CUDA(c, cudaEventCreate(&c->ev_start));
    SFE_CUDA(c, cudaEventCreate(&c->ev_stop));
    for (int i = 0; i <= ST_COUNT; ++i) SFE_CUDA(c, cudaEventCreate(&c->ev_stage[i]));

    c->fast_th = f.fast_threshold; c->orb_th = f.orb_max_distance;
    c->NI = 2 * f.max_pairs;
    LevelPack& P = c->pack;
    memset(&P, 0, sizeof(P));
    P.nlevels = f.nlevels;

    // ---- level sizes and quotas (cv::ORB::operator(), OpenCV 2.4) ----------------------------
    const int L = f.nlevels;
    {
        const float factor = (float)(1.0 / (double)f.scale_factor);
        float per = f.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
        int sum = 0;
        for (int l = 0; l < L - 1; ++l) { P.g[l].quota = cv_round_d(per); sum += P.g[l].quota; per *= factor; }
        P.g[L - 1].quota = std::max(f.nfeatures - sum, 0);
    }
    size_t arena = 0, raw_ofs = 0; int sel_ofs = 0, ft = 0, bt = 0, cand_cap = 1024;
    for (int l = 0; l < L; ++l) {
        LevelGeom& g = P.g[l];
        g.scale = (float)pow((double)f.scale_factor, (double)l);
        if (f.pyramid_mode == SFE_PYRAMID_CV4_EXACT) {          // cv >= 3: Size(cvRound(cols/scale), cvRound(rows/scale))
            g.w = cv_round_d((double)((float)f.width / g.scale));
            g.h = cv_round_d((double)((float)f.height / g.scale));
        } else {                                                // cv 2.4: scale = 1/getScale(level); cvRound(cols*scale)
            const float inv = 1 / g.scale;
            g.w = cv_round_d((double)((float)f.width * inv));
            g.h = cv_round_d((double)((float)f.height * inv));
        }
        g.pitch = (g.w + 127) / 128 * 128;
        g.img_bytes = (size_t)g.pitch * g.h;
        g.lvl_ofs = arena;
        arena += g.img_bytes * c->NI;
        arena = (arena + 255) / 256 * 256;
        const int m = std::max(f.edge_threshold, 3);
        g.rx0 = m; g.ry0 = m; g.rx1 = g.w - m; g.ry1 = g.h - m;
        const int rw = std::max(g.rx1 - g.rx0, 0), rh = std::max(g.ry1 - g.ry0, 0);
        if (rw == 0 || rh == 0 || g.w < 16 || g.h < 16) { g.rx1 = g.rx0; g.ry1 = g.ry0; g.ftiles_x = g.ftiles_y = 0; }
        else { g.ftiles_x = (g.rx1 - (g.rx0 & ~3) + FAST_TW - 1) / FAST_TW; g.ftiles_y = (rh + FAST_TH - 1) / FAST_TH; }   // tile origin is 4-aligned
        g.btiles_x = (g.w + BLUR_TW - 1) / BLUR_TW; g.btiles_y = (g.h + BLUR_TH - 1) / BLUR_TH;
        g.raw_cap = std::max(((rw + 1) / 2) * ((rh + 1) / 2), 1);      // NMS: no two 8-adjacent maxima
        g.raw_ofs = raw_ofs; raw_ofs += g.raw_cap;
        g.sel_cap = g.quota + SEL_TIE_SLACK;
        g.sel_ofs = sel_ofs; sel_ofs += g.sel_cap;
        P.ftile_prefix[l] = ft; ft += g.ftiles_x * g.ftiles_y;
        P.btile_prefix[l] = bt; bt += g.btiles_x * g.btiles_y;
        cand_cap = std::max(cand_cap, std::min(g.raw_cap, 2 * g.quota + std::max(1024, g.quota)));
        if (g.w < 8 || g.h < 8) { c->err = "pyramid level smaller than 8 px"; return SFE_ERR_INVALID; }
    }
    P.ftile_prefix[L] = ft; P.btile_prefix[L] = bt;
    P.raw_img_stride = raw_ofs; P.sel_img_stride = sel_ofs; P.kp_cap = sel_ofs;
    c->sel_cand_cap = std::min(cand_cap, 24 * 1024);
    c->arena_bytes = arena;
    if (P.kp_cap > 65535) { c->err = "keypoint capacity exceeds 65535"; return SFE_ERR_INVALID; }

    // ---- device memory ------------------------------------------------------------------------
    const int NI = c->NI;
    SFE_CUDA(c, dalloc(&c->d_pyr, arena));
    SFE_CUDA(c, cudaMemset(c->d_pyr, 0, arena));
    c->force_unfused = getenv("SFE_FORCE_UNFUSED_BLUR") != nullptr;
    c->pdl = getenv("SFE_NO_PDL") == nullptr;
    c->pyr_fork = getenv("SFE_NO_PYR_FORK") == nullptr;
    if (getenv("SFE_PYR_BRANCHES")) c->pyr_branches = std::min(4, std::max(2, atoi(getenv("SFE_PYR_BRANCHES"))));
    if (getenv("SFE_PDL_MASK")) c->pdl_mask = atoi(getenv("SFE_PDL_MASK"));
    if (!describe_is_fused(c)) {           // the blurred arena only exists where a separate blur kernel runs (float mode, edge_threshold < 22)
        SFE_CUDA(c, dalloc(&c->d_blur, arena));
        SFE_CUDA(c, cudaMemset(c->d_blur, 0, arena));
    }
    SFE_CUDA(c, dalloc(&c->d_raw, (size_t)NI * P.raw_img_stride));
    SFE_CUDA(c, dalloc(&c->d_raw_count, (size_t)NI * L));
    SFE_CUDA(c, dalloc(&c->d_desc_ctr, (size_t)4));
    SFE_CUDA(c, cudaMemset(c->d_desc_ctr, 0, 16));
    SFE_CUDA(c, dalloc(&c->d_sel_xy, (size_t)NI * P.sel_img_stride));
    SFE_CUDA(c, dalloc(&c->d_sel_resp, (size_t)NI * P.sel_img_stride));
    SFE_CUDA(c, dalloc(&c->d_sel_count, (size_t)NI * L));
    SFE_CUDA(c, dalloc(&c->d_kp, (size_t)NI * P.kp_cap));
    SFE_CUDA(c, dalloc(&c->d_desc, (size_t)NI * P.kp_cap * 32));
    SFE_CUDA(c, dalloc(&c->d_kp_count, (size_t)NI));
    SFE_CUDA(c, dalloc(&c->d_matches, (size_t)f.max_pairs * P.kp_cap));
    SFE_CUDA(c, dalloc(&c->d_match_count, (size_t)f.max_pairs));
    SFE_CUDA(c, dalloc(&c->d_rb_best, (size_t)f.max_pairs * P.kp_cap));
    SFE_CUDA(c, dalloc(&c->d_rb_owner, (size_t)f.max_pairs * P.kp_cap));
    SFE_CUDA(c, dalloc(&c->d_status, (size_t)NI));
    SFE_CUDA(c, cudaMemset(c->d_status, 0, sizeof(int32_t) * NI));
    SFE_CUDA(c, cudaMemset(c->d_kp_count, 0, sizeof(int32_t) * NI));
    SFE_CUDA(c, cudaMemset(c->d_match_count, 0, sizeof(int32_t) * f.max_pairs));
    // pinned staging
    c->h_images_bytes = (size_t)NI * f.width * f.height;
    SFE_CUDA(c, halloc(&c->h_images, c->h_images_bytes));
    SFE_CUDA(c, halloc(&c->h_kp, (size_t)NI * P.kp_cap));
    SFE_CUDA(c, halloc(&c->h_desc, (size_t)NI * P.kp_cap * 32));
    SFE_CUDA(c, halloc(&c->h_matches, (size_t)f.max_pairs * P.kp_cap));
    SFE_CUDA(c, halloc(&c->h_kp_count, (size_t)NI));
    SFE_CUDA(c, halloc(&c->h_match_count, (size_t)f.max_pairs));
    SFE_CUDA(c, halloc(&c->h_status, (size_t)NI));
    SFE_CUDA(c, halloc(&c->h_counts, (size_t)f.max_pairs));

    // ---- resize tables ------------------------------------------------------------------------
    c->xtab.resize(L); c->ytab.resize(L);
    for (int l = 1; l < L; ++l) {
        if (f.pyramid_mode == SFE_PYRAMID_CV4_EXACT) {
            build_resize_table_exact(P.g[l - 1].w, P.g[l].w, c->xtab[l]);
            build_resize_table_exact(P.g[l - 1].h, P.g[l].h, c->ytab[l]);
        } else {
            build_resize_table(P.g[l - 1].w, P.g[l].w, true, c->xtab[l]);
            build_resize_table(P.g[l - 1].h, P.g[l].h, false, c->ytab[l]);
        }
        if ((P.g[l].w + PYR_TW - 1) / PYR_TW > PYR_MAX_TILES || (P.g[l].h + PYR_TH - 1) / PYR_TH > PYR_MAX_TILES) { c->err = "image too large for the pyramid kernel's tile tables"; return SFE_ERR_INVALID; }
        // footprint check against the fixed TMA box
        for (int x0 = 0; x0 < P.g[l].w; x0 += PYR_TW) {
            const int x1 = std::min(x0 + PYR_TW, P.g[l].w) - 1;
            if (c->xtab[l][x1].x + 2 - (c->xtab[l][x0].x & ~15) > PYR_BOX_W) { c->err = "pyramid tile footprint exceeds TMA box (x)"; return SFE_ERR_INVALID; }
        }
        for (int y0 = 0; y0 < P.g[l].h; y0 += PYR_TH) {
            const int y1 = std::min(y0 + PYR_TH, P.g[l].h) - 1;
            if (c->ytab[l][y1].x + 2 - std::max(c->ytab[l][y0].x, 0) > PYR_BOX_H) { c->err = "pyramid tile footprint exceeds TMA box (y)"; return SFE_ERR_INVALID; }
        }
        {   // gather plan of the horizontal pass, one entry per 4 destination columns
            const int ng = (P.g[l].w + 3) / 4;
            std::vector<uint32_t> gt((size_t)ng * 6);
            for (int q = 0; q < ng; ++q) {
                int sx[4]; uint32_t cf[4];
                for (int e = 0; e < 4; ++e) {
                    const int d = std::min(4 * q + e, P.g[l].w - 1);
                    sx[e] = c->xtab[l][d].x; cf[e] = (uint32_t)c->xtab[l][d].y;
                }
                uint32_t sel = 0;
                for (int e = 0; e < 4; ++e) {
                    const int o = sx[e] - sx[0];
                    if (o < 0 || o > 6) { c->err = "resize gather plan: offset out of PRMT range"; return SFE_ERR_INVALID; }
                    sel |= (uint32_t)(o | ((o + 1) << 4)) << (8 * e);
                }
                gt[(size_t)q * 6 + 0] = (uint32_t)sx[0]; gt[(size_t)q * 6 + 1] = sel;
                for (int e = 0; e < 4; ++e) gt[(size_t)q * 6 + 2 + e] = cf[e];
            }
            SFE_CUDA(c, dalloc(&c->d_gtab[l], gt.size()));
            SFE_CUDA(c, cudaMemcpy(c->d_gtab[l], gt.data(), gt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        {   // vertical table as the kernel consumes it: both source rows clamped to the image, coefficients pre-shifted
            const int sh = P.g[l - 1].h;
            std::vector<int4> yt(c->ytab[l].size());
            for (size_t y = 0; y < yt.size(); ++y) {
                const int s = c->ytab[l][y].x, cf = c->ytab[l][y].y;
                yt[y].x = std::min(std::max(s, 0), sh - 1) * (PYR_TW / 2) * 4; yt[y].y = std::min(std::max(s + 1, 0), sh - 1) * (PYR_TW / 2) * 4;
                yt[y].z = (int)((uint32_t)(cf & 0xffff) << 16); yt[y].w = (int)((uint32_t)cf >> 16 << 16);
            }
            SFE_CUDA(c, dalloc(&c->d_ytab[l], yt.size()));
            SFE_CUDA(c, cudaMemcpy(c->d_ytab[l], yt.data(), yt.size() * sizeof(int4), cudaMemcpyHostToDevice));
        }
    }

    // ---- TMA descriptors: 3-D (x, y, image) maps per level, one per box shape; every kernel loads its full box ----
    PFN_encodeTiled enc = nullptr;
    {
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        SFE_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { c->err = "cuTensorMapEncodeTiled not available"; return SFE_ERR_CUDA; }
        enc = (PFN_encodeTiled)fn;
    }
    // The box shape is baked into a map: pyr_map[] (160x40) serves FAST and blur, pyrsrc_map[] (176x88)
    // the source side of the pyramid kernel.
    for (int l = 0; l < L; ++l) {
        const LevelGeom& g = P.g[l];
        int r = make_image_map(c, enc, &P.pyr_map[l], c->d_pyr + g.lvl_ofs, g.w, g.h, g.pitch, NI, FAST_BOX_W, FAST_BOX_H);
        if (r) return r;
        r = make_image_map(c, enc, &c->pyrsrc_map[l], c->d_pyr + g.lvl_ofs, g.w, g.h, g.pitch, NI, PYR_BOX_W, PYR_BOX_H);
        if (r) return r;
    }
    {   // tile tables: blockIdx.x -> (level, tile origin), so the kernels need no search and no division
        std::vector<int4> ft_tab, bt_tab;
        for (int l = 0; l < L; ++l) {
            const LevelGeom& g = P.g[l];
            for (int ty = 0; ty < g.ftiles_y; ++ty) for (int tx = 0; tx < g.ftiles_x; ++tx)
                ft_tab.push_back(make_int4(l, (g.rx0 & ~3) + tx * FAST_TW, g.ry0 + ty * FAST_TH, 0));
            for (int ty = 0; ty < g.btiles_y; ++ty) for (int tx = 0; tx < g.btiles_x; ++tx)
                bt_tab.push_back(make_int4(l, tx * BLUR_TW, ty * BLUR_TH, 0));
        }
        if (!ft_tab.empty()) {
            SFE_CUDA(c, dalloc(&c->d_ftile_tab, ft_tab.size()));
            SFE_CUDA(c, cudaMemcpy(c->d_ftile_tab, ft_tab.data(), ft_tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
        }
        SFE_CUDA(c, dalloc(&c->d_btile_tab, bt_tab.size()));
        SFE_CUDA(c, cudaMemcpy(c->d_btile_tab, bt_tab.data(), bt_tab.size() * sizeof(int4), cudaMemcpyHostToDevice));
    }
    // blur taps for the float mode: getGaussianKernel(7, 2, CV_32F)
    {
        float k[7]; double sum = 0;
        for (int i = 0; i < 7; ++i) { const double x = i - 3.0; k[i] = (float)exp(-0.5 / 4.0 * x * x); sum += k[i]; }
        sum = 1. / sum;
        for (int i = 0; i < 7; ++i) k[i] = (float)(k[i] * sum);
        sfe_upload_blur_taps(k);
    }
    c->graphs.assign(f.max_pairs + 1, nullptr);
    c->ship_kp = std::min(P.kp_cap, f.nfeatures + 16); c->ship_m = std::min(P.kp_cap, f.nfeatures + 16);
    if ((size_t)c->sel_cand_cap * 8 > 200 * 1024 || sizeof(int) * ((size_t)f.height + 2 + (size_t)P.kp_cap) > 200 * 1024) {
        c->err = "configuration needs more than 200 KB of shared memory per CTA"; return SFE_ERR_INVALID;
    }
    prepare_select(c);
    prepare_rowband(c);
    prepare_describe(c);
    SFE_CUDA(c, cudaGetLastError());
    SFE_CUDA(c, cudaDeviceSynchronize());
    return SFE_OK;
}

extern "C" int sfe_create(const sfe_config* cfg, sfe_context** out)
{
    if (!cfg || !out) { g_create_err = "null argument"; return SFE_ERR_INVALID; }
    SfeDeviceGuard guard(cfg->device);     // create_impl binds the device; the caller's current device is restored on return
    sfe_context* c = new sfe_context();
    c->cfg = *cfg;
    const int r = create_impl(c);
    if (r != SFE_OK) { g_create_err = c->err; sfe_destroy(c); *out = nullptr; return r; }
    *out = c;
    return SFE_OK;
}

extern "C" void sfe_destroy(sfe_context* c)
{
    SFE_BIND_DEVICE(c);
    if (!c) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto g : c->graphs) if (g) cudaGraphExecDestroy(g);
    cudaFree(c->d_pyr); cudaFree(c->d_blur); cudaFree(c->d_raw); cudaFree(c->d_raw_count); cudaFree(c->d_desc_ctr);
    cudaFree(c->d_sel_xy); cudaFree(c->d_sel_resp); cudaFree(c->d_sel_count); cudaFree(c->d_kp); cudaFree(c->d_desc);
    cudaFree(c->d_rb_best); cudaFree(c->d_rb_owner); cudaFree(c->d_kp_count); cudaFree(c->d_matches); cudaFree(c->d_match_count); cudaFree(c->d_status);
    cudaFree(c->d_packed); cudaFree(c->d_ftile_tab); cudaFree(c->d_btile_tab); cudaFree(c->d_bf_q); cudaFree(c->d_bf_t); cudaFree(c->d_bf_best); cudaFree(c->d_bf_ofs);
    for (int l = 0; l < SFE_MAX_LEVELS; ++l) { cudaFree(c->d_ytab[l]); cudaFree(c->d_gtab[l]); }
    cudaFreeHost(c->h_images); cudaFreeHost(c->h_kp); cudaFreeHost(c->h_desc); cudaFreeHost(c->h_matches);
    cudaFreeHost(c->h_kp_count); cudaFreeHost(c->h_match_count); cudaFreeHost(c->h_status); cudaFreeHost(c->h_counts);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    for (int i = 0; i <= ST_COUNT; ++i) if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int b = 0; b < 3; ++b) { if (c->ev_branch[b]) cudaEventDestroy(c->ev_branch[b]); if (c->stream2[b]) cudaStreamDestroy(c->stream2[b]); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// ---- thresholds -----------------------------------------------------------------------------------
extern "C" int sfe_set_fast_threshold(sfe_context* c, int th)
{
    if (!c || th < 1 || th > 127) return SFE_ERR_INVALID;
    if (th != c->fast_th) { c->fast_th = th; for (auto& g : c->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; } }
    return SFE_OK;
}
extern "C" int sfe_get_fast_threshold(const sfe_context* c) { return c ? c->fast_th : -1; }
extern "C" int sfe_set_orb_threshold(sfe_context* c, int th)
{
    if (!c || th < 0 || th > 256) return SFE_ERR_INVALID;
    if (th != c->orb_th) { c->orb_th = th; for (auto& g : c->graphs) if (g) { cudaGraphExecDestroy(g); g = nullptr; } }
    return SFE_OK;
}
extern "C" int sfe_get_orb_threshold(const sfe_context* c) { return c ? c->orb_th : -1; }

extern "C" int sfe_num_levels(const sfe_context* c) { return c ? c->pack.nlevels : 0; }
extern "C" int sfe_level_size(const sfe_context* c, int l, int32_t* w, int32_t* h, int32_t* pitch, int32_t* quota)
{
    if (!c || l < 0 || l >= c->pack.nlevels) return SFE_ERR_INVALID;
    if (w) *w = c->pack.g[l].w; if (h) *h = c->pack.g[l].h; if (pitch) *pitch = c->pack.g[l].pitch; if (quota) *quota = c->pack.g[l].quota;
    return SFE_OK;
}
extern "C" int sfe_keypoint_capacity(const sfe_context* c) { return c ? c->pack.kp_cap : 0; }

// ---- batch orchestration ------------------------------------------------------------------------------
static void stage_mark(sfe_context* c, int idx) { if (c->profiling) cudaEventRecord(c->ev_stage[idx], c->stream); }

static void enqueue_kernels(sfe_context* c, int npairs)
{
    const int nimg = 2 * npairs, L = c->pack.nlevels;
    cudaMemsetAsync(c->d_raw_count, 0, sizeof(int32_t) * (size_t)nimg * L, c->stream);
    stage_mark(c, ST_PYRAMID);
    // The pyramid is a chain of L-1 dependent launches whose grids shrink level by level; every launch ends on a partial wave
    // (2.08 waves at level 7: a third round for 8 % of the CTAs).  With the batch split in two image halves on two streams
    // (forked and joined with events, also inside the captured graph) one half's tail runs under the other half's next level.
    const bool fork = L > 2 && nimg >= 64 && !c->profiling && c->pyr_fork;
    if (fork) {
        const int NBR = c->pyr_branches;                          // 2 .. 4 image ranges, the first on the main stream
        cudaEventRecord(c->ev_fork, c->stream);
        for (int b = 1; b < NBR; ++b) cudaStreamWaitEvent(c->stream2[b - 1], c->ev_fork, 0);
        for (int l = 1; l < L; ++l)
            for (int b = 0; b < NBR; ++b) {
                const int i0 = (int)((long long)nimg * b / NBR), i1 = (int)((long long)nimg * (b + 1) / NBR);
                c->launch_stream = b ? c->stream2[b - 1] : nullptr;
                launch_pyramid_level(c, l, i0, i1 - i0);
            }
        c->launch_stream = nullptr;
        for (int b = 1; b < NBR; ++b) { cudaEventRecord(c->ev_branch[b - 1], c->stream2[b - 1]); cudaStreamWaitEvent(c->stream, c->ev_branch[b - 1], 0); }
    } else {
        for (int l = 1; l < L; ++l) launch_pyramid_level(c, l, 0, nimg);
    }
    stage_mark(c, ST_FAST);
    launch_fast(c, nimg, !fork);
    stage_mark(c, ST_SELECT);
    launch_select(c, nimg);
    stage_mark(c, ST_BLUR);
    if (!describe_is_fused(c)) launch_blur(c, nimg);
    stage_mark(c, ST_DESCRIBE);
    launch_describe(c, nimg);
    stage_mark(c, ST_MATCH_LR);
    launch_match_rowband(c, npairs);
    stage_mark(c, ST_D2H);
}

static int run_batch(sfe_context* c, int npairs)
{
    if (npairs < 1 || npairs > c->cfg.max_pairs) { c->err = "npairs out of range"; return SFE_ERR_INVALID; }
    if (c->cfg.use_graph && !c->profiling) {
        cudaGraphExec_t& ge = c->graphs[npairs];
        if (!ge) {
            const int64_t before = c->launches;
            cudaGraph_t g = nullptr;
            SFE_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            enqueue_kernels(c, npairs);
            SFE_CUDA(c, cudaStreamEndCapture(c->stream, &g));
            SFE_CUDA(c, cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            c->graph_kernels = (int)(c->launches - before);
            c->launches = before;
        }
        SFE_CUDA(c, cudaGraphLaunch(ge, c->stream));
        c->launches += c->graph_kernels;
    } else {
        enqueue_kernels(c, npairs);
        SFE_CUDA(c, cudaGetLastError());
    }
    c->last_npairs = npairs; c->have_results = true; c->downloaded = false;
    return SFE_OK;
}

static int upload(sfe_context* c, const uint8_t* images, int npairs, size_t stride, size_t image_stride)
{
    if (!images || npairs < 1 || npairs > c->cfg.max_pairs) { c->err = "bad upload arguments"; return SFE_ERR_INVALID; }
    const LevelGeom& g = c->pack.g[0];
    if (stride < (size_t)g.w || image_stride < stride * (size_t)(g.h - 1) + g.w) { c->err = "bad strides"; return SFE_ERR_INVALID; }
    if (stride == (size_t)g.w && image_stride == (size_t)g.w * g.h) {
        // tightly packed: one linear DMA, then a kernel spreads the rows onto the pitched level-0 arena
        if (!c->d_packed) SFE_CUDA(c, dalloc(&c->d_packed, (size_t)c->NI * g.w * g.h));
        SFE_CUDA(c, cudaMemcpyAsync(c->d_packed, images, (size_t)2 * npairs * g.w * g.h, cudaMemcpyHostToDevice, c->stream));
        launch_unpack(c, 2 * npairs);
        return SFE_OK;
    }
    // one strided 3-D copy: rows of w bytes, h rows, 2*npairs images
    cudaMemcpy3DParms p = {};
    if (image_stride % stride == 0) {
        p.srcPtr = make_cudaPitchedPtr((void*)images, stride, g.w, image_stride / stride);
        p.dstPtr = make_cudaPitchedPtr(c->d_pyr + g.lvl_ofs, g.pitch, g.w, g.h);
        p.extent = make_cudaExtent(g.w, g.h, 2 * npairs);
        p.kind = cudaMemcpyHostToDevice;
        SFE_CUDA(c, cudaMemcpy3DAsync(&p, c->stream));
    } else {
        for (int i = 0; i < 2 * npairs; ++i)
            SFE_CUDA(c, cudaMemcpy2DAsync(c->d_pyr + g.lvl_ofs + (size_t)i * g.img_bytes, g.pitch, images + (size_t)i * image_stride,
                                          stride, g.w, g.h, cudaMemcpyHostToDevice, c->stream));
    }
    return SFE_OK;
}

extern "C" uint8_t* sfe_host_staging(sfe_context* c, size_t* bytes)
{
    if (!c) return nullptr;
    if (bytes) *bytes = c->h_images_bytes;
    return c->h_images;
}

extern "C" int sfe_upload_batch(sfe_context* c, const uint8_t* images, int npairs, size_t stride, size_t image_stride)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    stage_mark(c, ST_H2D);
    return upload(c, images, npairs, stride, image_stride);
}

extern "C" int sfe_process_batch_device(sfe_context* c, int npairs)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    return run_batch(c, npairs);
}

static int fold_status(sfe_context* c, int npairs)
{
    SFE_CUDA(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int32_t) * 2 * npairs, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    int st = SFE_OK;
    for (int i = 0; i < 2 * npairs; ++i) if (c->h_status[i]) st = SFE_ERR_CAPACITY;
    if (st) { c->err = "a device-side list overflowed (raise capacities / lower nfeatures)"; cudaMemsetAsync(c->d_status, 0, sizeof(int32_t) * c->NI, c->stream); }
    return st;
}

extern "C" int sfe_synchronize(sfe_context* c)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    if (c->have_results) return fold_status(c, c->last_npairs);
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFE_OK;
}

extern "C" int sfe_process_batch_host(sfe_context* c, const uint8_t* images, int npairs, size_t stride, size_t image_stride)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    stage_mark(c, ST_H2D);
    int r = upload(c, images, npairs, stride, image_stride);
    if (r) return r;
    r = run_batch(c, npairs);
    if (r) return r;
    return fold_status(c, npairs);
}

extern "C" int sfe_process_stereo(sfe_context* c, const uint8_t* left, const uint8_t* right, size_t stride, int repeat)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    if (!repeat) {
        if (!left || !right) { c->err = "null image"; return SFE_ERR_INVALID; }
        const LevelGeom& g = c->pack.g[0];
        if (stride < (size_t)g.w) { c->err = "bad stride"; return SFE_ERR_INVALID; }
        stage_mark(c, ST_H2D);
        if (stride == (size_t)g.w) {
            // tightly packed rows: two linear DMAs into the packed staging + the unpack kernel (a pitched copy costs one
            // DMA descriptor per row)
            const size_t bytes = (size_t)g.w * g.h;
            if (!c->d_packed) SFE_CUDA(c, dalloc(&c->d_packed, (size_t)c->NI * bytes));
            SFE_CUDA(c, cudaMemcpyAsync(c->d_packed, left, bytes, cudaMemcpyHostToDevice, c->stream));
            SFE_CUDA(c, cudaMemcpyAsync(c->d_packed + bytes, right, bytes, cudaMemcpyHostToDevice, c->stream));
            launch_unpack(c, 2);
        } else {
            SFE_CUDA(c, cudaMemcpy2DAsync(c->d_pyr + g.lvl_ofs, g.pitch, left, stride, g.w, g.h, cudaMemcpyHostToDevice, c->stream));
            SFE_CUDA(c, cudaMemcpy2DAsync(c->d_pyr + g.lvl_ofs + g.img_bytes, g.pitch, right, stride, g.w, g.h, cudaMemcpyHostToDevice, c->stream));
        }
    } else if (!c->have_results) { c->err = "repeat without a previous frame"; return SFE_ERR_STATE; }
    int r = run_batch(c, 1);
    if (r) return r;
    return fold_status(c, 1);
}

// ---- results --------------------------------------------------------------------------------------------
static int fetch_counts(sfe_context* c)
{
    if (!c->have_results) { c->err = "no results: call a process function first"; return SFE_ERR_STATE; }
    const int np = c->last_npairs;
    SFE_CUDA(c, cudaMemcpyAsync(c->h_kp_count, c->d_kp_count, sizeof(int32_t) * 2 * np, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaMemcpyAsync(c->h_match_count, c->d_match_count, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int32_t) * 2 * np, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int p = 0; p < np; ++p) {
        c->h_counts[p].n_left = c->h_kp_count[2 * p]; c->h_counts[p].n_right = c->h_kp_count[2 * p + 1];
        c->h_counts[p].n_matches = c->h_match_count[p];
        c->h_counts[p].status = (c->h_status[2 * p] | c->h_status[2 * p + 1]) ? SFE_ERR_CAPACITY : SFE_OK;
    }
    return SFE_OK;
}

extern "C" int sfe_get_counts(sfe_context* c, int pair, sfe_pair_counts* out)
{
    SFE_BIND_DEVICE(c);
    if (!c || !out) return SFE_ERR_INVALID;
    int r = fetch_counts(c); if (r) return r;
    if (pair < 0 || pair >= c->last_npairs) { c->err = "pair out of range"; return SFE_ERR_INVALID; }
    *out = c->h_counts[pair];
    return SFE_OK;
}

extern "C" int sfe_get_values(sfe_context* c, int pair, sfe_keypoint* kl, sfe_keypoint* kr, uint8_t* dl, uint8_t* dr,
                              sfe_dmatch* m, sfe_pair_counts* counts)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    int r = fetch_counts(c); if (r) return r;
    if (pair < 0 || pair >= c->last_npairs) { c->err = "pair out of range"; return SFE_ERR_INVALID; }
    const sfe_pair_counts pc = c->h_counts[pair];
    const size_t cap = c->pack.kp_cap;
    const size_t il = 2 * (size_t)pair, ir = il + 1;
    if (kl) SFE_CUDA(c, cudaMemcpyAsync(kl, c->d_kp + il * cap, sizeof(sfe_keypoint) * pc.n_left, cudaMemcpyDeviceToHost, c->stream));
    if (kr) SFE_CUDA(c, cudaMemcpyAsync(kr, c->d_kp + ir * cap, sizeof(sfe_keypoint) * pc.n_right, cudaMemcpyDeviceToHost, c->stream));
    if (dl) SFE_CUDA(c, cudaMemcpyAsync(dl, c->d_desc + il * cap * 32, 32 * (size_t)pc.n_left, cudaMemcpyDeviceToHost, c->stream));
    if (dr) SFE_CUDA(c, cudaMemcpyAsync(dr, c->d_desc + ir * cap * 32, 32 * (size_t)pc.n_right, cudaMemcpyDeviceToHost, c->stream));
    if (m) SFE_CUDA(c, cudaMemcpyAsync(m, c->d_matches + (size_t)pair * cap, sizeof(sfe_dmatch) * pc.n_matches, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    if (counts) *counts = pc;
    return SFE_OK;
}

// D2H ships CONTENT, not capacity: an image holds ~nfeatures keypoints in `cap` = nfeatures + 64*nlevels slots and a pair
// far fewer matches.  The asynchronous download copies the first `ship_kp` slots of every image (`ship_m` of every pair)
// with one strided copy per array -- sizes known on the host before the kernels ran, so nothing waits; sfe_wait_batch
// fetches the (rare) remainder when a count exceeds what was shipped and raises the sizes for the next batch.
static int copy_columns(sfe_context* c, void* h, const void* d, size_t elem, int first, int last, int rows)
{
    if (last <= first || rows <= 0) return SFE_OK;
    const size_t pitch = (size_t)c->pack.kp_cap * elem;
    SFE_CUDA(c, cudaMemcpy2DAsync((uint8_t*)h + (size_t)first * elem, pitch, (const uint8_t*)d + (size_t)first * elem, pitch,
                                  (size_t)(last - first) * elem, rows, cudaMemcpyDeviceToHost, c->stream));
    return SFE_OK;
}

static int enqueue_download(sfe_context* c)
{
    const int np = c->last_npairs;
    int r;
    if ((r = copy_columns(c, c->h_kp, c->d_kp, sizeof(sfe_keypoint), 0, c->ship_kp, 2 * np))) return r;
    if ((r = copy_columns(c, c->h_desc, c->d_desc, 32, 0, c->ship_kp, 2 * np))) return r;
    if ((r = copy_columns(c, c->h_matches, c->d_matches, sizeof(sfe_dmatch), 0, c->ship_m, np))) return r;
    SFE_CUDA(c, cudaMemcpyAsync(c->h_kp_count, c->d_kp_count, sizeof(int32_t) * 2 * np, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaMemcpyAsync(c->h_match_count, c->d_match_count, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int32_t) * 2 * np, cudaMemcpyDeviceToHost, c->stream));
    c->shipped_kp = c->ship_kp; c->shipped_m = c->ship_m;
    c->d2h_bytes_last = (size_t)2 * np * c->ship_kp * (sizeof(sfe_keypoint) + 32) + (size_t)np * c->ship_m * sizeof(sfe_dmatch) + (size_t)np * 20;
    return SFE_OK;
}

// After the counts have landed: fetch what the speculative copy missed, adapt the shipping sizes.
static int complete_download(sfe_context* c)
{
    const int np = c->last_npairs, cap = c->pack.kp_cap;
    int kmax = 0, mmax = 0;
    for (int i = 0; i < 2 * np; ++i) kmax = std::max(kmax, std::min(c->h_kp_count[i], cap));
    for (int p = 0; p < np; ++p) mmax = std::max(mmax, std::min(c->h_match_count[p], cap));
    bool more = false; int r;
    if (kmax > c->shipped_kp) {
        if ((r = copy_columns(c, c->h_kp, c->d_kp, sizeof(sfe_keypoint), c->shipped_kp, kmax, 2 * np))) return r;
        if ((r = copy_columns(c, c->h_desc, c->d_desc, 32, c->shipped_kp, kmax, 2 * np))) return r;
        c->d2h_bytes_last += (size_t)2 * np * (kmax - c->shipped_kp) * (sizeof(sfe_keypoint) + 32);
        more = true;
    }
    if (mmax > c->shipped_m) {
        if ((r = copy_columns(c, c->h_matches, c->d_matches, sizeof(sfe_dmatch), c->shipped_m, mmax, np))) return r;
        c->d2h_bytes_last += (size_t)np * (mmax - c->shipped_m) * sizeof(sfe_dmatch);
        more = true;
    }
    if (more) { SFE_CUDA(c, cudaStreamSynchronize(c->stream)); c->d2h_refetches++; }
    c->ship_kp = std::min(cap, std::max(c->ship_kp, kmax + 16));              // keypoint counts sit at the quota: only ever grows
    c->ship_m = std::min(cap, std::max(mmax + 64, c->ship_m - (c->ship_m - mmax) / 4));   // matches vary with the scene: follows the recent maximum
    return SFE_OK;
}

static void fill_counts(sfe_context* c)
{
    for (int p = 0; p < c->last_npairs; ++p) {
        c->h_counts[p].n_left = c->h_kp_count[2 * p]; c->h_counts[p].n_right = c->h_kp_count[2 * p + 1];
        c->h_counts[p].n_matches = c->h_match_count[p];
        c->h_counts[p].status = (c->h_status[2 * p] | c->h_status[2 * p + 1]) ? SFE_ERR_CAPACITY : SFE_OK;
    }
}

extern "C" int sfe_submit_batch_host(sfe_context* c, const uint8_t* images, int npairs, size_t stride, size_t image_stride)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    int r = upload(c, images, npairs, stride, image_stride);
    if (r) return r;
    r = run_batch(c, npairs);
    if (r) return r;
    return enqueue_download(c);
}

extern "C" int sfe_wait_batch(sfe_context* c, const sfe_keypoint** kps, const uint8_t** desc, const sfe_dmatch** matches,
                              const sfe_pair_counts** counts)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    if (!c->have_results) { c->err = "no batch in flight"; return SFE_ERR_STATE; }
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    fill_counts(c);
    { const int r = complete_download(c); if (r) return r; }
    if (kps) *kps = c->h_kp; if (desc) *desc = c->h_desc; if (matches) *matches = c->h_matches; if (counts) *counts = c->h_counts;
    c->downloaded = true;
    for (int p = 0; p < c->last_npairs; ++p)
        if (c->h_counts[p].status) {
            c->err = "a device-side list overflowed (raise capacities / lower nfeatures)";
            cudaMemsetAsync(c->d_status, 0, sizeof(int32_t) * c->NI, c->stream);
            return SFE_ERR_CAPACITY;
        }
    return SFE_OK;
}

extern "C" int sfe_download_batch(sfe_context* c, const sfe_keypoint** kps, const uint8_t** desc, const sfe_dmatch** matches,
                                  const sfe_pair_counts** counts)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    if (!c->have_results) { c->err = "no results"; return SFE_ERR_STATE; }
    stage_mark(c, ST_D2H);
    int r = fetch_counts(c); if (r) return r;                      // synchronous entry: counts first, then exactly the content
    c->shipped_kp = 0; c->shipped_m = 0; c->d2h_bytes_last = (size_t)c->last_npairs * 20;
    const int64_t refetches = c->d2h_refetches;
    r = complete_download(c); if (r) return r;
    c->d2h_refetches = refetches;                                   // not a mis-speculation: nothing had been shipped yet
    stage_mark(c, ST_COUNT);
    if (kps) *kps = c->h_kp; if (desc) *desc = c->h_desc; if (matches) *matches = c->h_matches; if (counts) *counts = c->h_counts;
    c->downloaded = true;
    return SFE_OK;
}

extern "C" int sfe_get_level(sfe_context* c, int img, int level, int blurred, uint8_t* out)
{
    SFE_BIND_DEVICE(c);
    if (!c || !out || level < 0 || level >= c->pack.nlevels || img < 0 || img >= c->NI) return SFE_ERR_INVALID;
    const LevelGeom& g = c->pack.g[level];
    if (blurred && describe_is_fused(c)) {
        // debug / parity tap: the hot path never materialises the blurred levels in this mode -- blur them now, once, on demand
        if (!c->have_results) { c->err = "no results: call a process function first"; return SFE_ERR_STATE; }
        if (!c->d_blur) { SFE_CUDA(c, dalloc(&c->d_blur, c->arena_bytes)); SFE_CUDA(c, cudaMemsetAsync(c->d_blur, 0, c->arena_bytes, c->stream)); }
        launch_blur(c, 2 * c->last_npairs);
        SFE_CUDA(c, cudaGetLastError());
    }
    const uint8_t* base = (blurred ? c->d_blur : c->d_pyr) + g.lvl_ofs + (size_t)img * g.img_bytes;
    SFE_CUDA(c, cudaMemcpy2DAsync(out, g.w, base, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFE_OK;
}

extern "C" int sfe_get_raw_corners(sfe_context* c, int img, int level, uint32_t* out, int cap, int32_t* count)
{
    SFE_BIND_DEVICE(c);
    if (!c || level < 0 || level >= c->pack.nlevels || img < 0 || img >= c->NI || !count) return SFE_ERR_INVALID;
    const LevelGeom& g = c->pack.g[level];
    int32_t n = 0;
    SFE_CUDA(c, cudaMemcpyAsync(&n, c->d_raw_count + (size_t)img * c->pack.nlevels + level, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    *count = n;
    const int m = std::min(std::min(n, cap), g.raw_cap);
    if (out && m > 0) {
        SFE_CUDA(c, cudaMemcpyAsync(out, c->d_raw + (size_t)img * c->pack.raw_img_stride + g.raw_ofs, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, c->stream));
        SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return SFE_OK;
}

extern "C" int sfe_device_results(sfe_context* c, const void** kps, const void** desc, const void** counts, const void** matches,
                                  const void** match_counts)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    if (kps) *kps = c->d_kp; if (desc) *desc = c->d_desc; if (counts) *counts = c->d_kp_count;
    if (matches) *matches = c->d_matches; if (match_counts) *match_counts = c->d_match_count;
    return SFE_OK;
}

// ---- brute-force match ---------------------------------------------------------------------------------
extern "C" int sfe_match_bf_batch_device(sfe_context* c, const uint8_t* dq, const int32_t* dqo, const uint8_t* dt, const int32_t* dto,
                                         int npairs, int max_nq, uint32_t* d_best)
{
    SFE_BIND_DEVICE(c);
    if (!c || !dq || !dqo || !dt || !dto || !d_best || npairs < 1 || npairs > 65535 || max_nq < 1) return SFE_ERR_INVALID;
    // the kernel merges its train splits with atomicMin and returns early for empty train sets: the keys must start
    // at 0xFFFFFFFF.  One init kernel walks the query offsets (they live on the device) up to q_ofs[npairs].
    launch_bf_init_ofs(c->stream, d_best, dqo, npairs);
    launch_match_bf(c->stream, dq, dqo, dt, dto, npairs, max_nq, d_best);
    c->launches += 2;
    SFE_CUDA(c, cudaGetLastError());
    return SFE_OK;
}

extern "C" int sfe_match_bf(sfe_context* c, const uint8_t* q, int nq, const uint8_t* t, int nt, sfe_dmatch* out, int32_t* n_out)
{
    SFE_BIND_DEVICE(c);
    if (!c || nq < 0 || nt < 0 || !n_out) return SFE_ERR_INVALID;
    *n_out = 0;
    if (nt > 65535) { c->err = "sfe_match_bf: train set larger than 65535 descriptors (the packed key keeps 16 index bits)"; return SFE_ERR_INVALID; }
    if (nq == 0 || nt == 0) return SFE_OK;        // BFMatcher: empty train set -> no matches
    if (!q || !t || !out) return SFE_ERR_INVALID;
    if (nq > c->bf_cap_q) {
        cudaFree(c->d_bf_q); cudaFree(c->d_bf_best);
        c->bf_cap_q = std::max(nq, 4096);
        SFE_CUDA(c, dalloc(&c->d_bf_q, (size_t)c->bf_cap_q * 32));
        SFE_CUDA(c, dalloc(&c->d_bf_best, (size_t)c->bf_cap_q));
    }
    if (nt > c->bf_cap_t) {
        cudaFree(c->d_bf_t);
        c->bf_cap_t = std::max(nt, 4096);
        SFE_CUDA(c, dalloc(&c->d_bf_t, (size_t)c->bf_cap_t * 32));
    }
    if (!c->d_bf_ofs) SFE_CUDA(c, dalloc(&c->d_bf_ofs, 4));
    const int32_t ofs[4] = { 0, nq, 0, nt };
    SFE_CUDA(c, cudaMemcpyAsync(c->d_bf_ofs, ofs, sizeof(ofs), cudaMemcpyHostToDevice, c->stream));
    SFE_CUDA(c, cudaMemcpyAsync(c->d_bf_q, q, (size_t)nq * 32, cudaMemcpyHostToDevice, c->stream));
    SFE_CUDA(c, cudaMemcpyAsync(c->d_bf_t, t, (size_t)nt * 32, cudaMemcpyHostToDevice, c->stream));
    launch_bf_init(c->stream, c->d_bf_best, nq);
    launch_match_bf(c->stream, c->d_bf_q, c->d_bf_ofs, c->d_bf_t, c->d_bf_ofs + 2, 1, nq, c->d_bf_best);
    c->launches += 2;
    std::vector<uint32_t> best(nq);
    SFE_CUDA(c, cudaMemcpyAsync(best.data(), c->d_bf_best, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < nq; ++i) {
        out[i].queryIdx = i; out[i].trainIdx = (int)(best[i] & 0xffffu); out[i].imgIdx = 0; out[i].distance = (float)(best[i] >> 16);
    }
    *n_out = nq;
    return SFE_OK;
}

// ---- pair-list brute force over local / peer keyframe stores ---------------------------------------------------
extern "C" int sfe_match_bf_pairs_device(sfe_context* c, const uint8_t* d_query, const uint8_t* const* d_train_bases,
                                         const sfe_bf_pair* d_pairs, int npairs, int max_nq, uint32_t* d_best, size_t best_entries)
{
    SFE_BIND_DEVICE(c);
    if (!c || !d_query || !d_train_bases || !d_pairs || !d_best || npairs < 1 || npairs > 65535 * 16) return SFE_ERR_INVALID;
    SFE_CUDA(c, cudaMemsetAsync(d_best, 0xff, best_entries * sizeof(uint32_t), c->stream));
    for (int p0 = 0; p0 < npairs; p0 += 65535) {                 // gridDim.z limit
        const int n = std::min(65535, npairs - p0);
        launch_bf_pairs(c->stream, d_query, d_train_bases, d_pairs + p0, n, max_nq, d_best);
        c->launches++;
    }
    SFE_CUDA(c, cudaGetLastError());
    return SFE_OK;
}

// ---- raw device memory + CUDA IPC (peer-mapped keyframe stores across the one-process-per-GPU ranks) ------------
extern "C" int sfe_device_alloc(sfe_context* c, size_t bytes, void** out)
{
    SFE_BIND_DEVICE(c);
    if (!c || !out) return SFE_ERR_INVALID;
    SFE_CUDA(c, cudaMalloc(out, bytes));
    return SFE_OK;
}
extern "C" int sfe_device_free(sfe_context* c, void* p) { if (!c) return SFE_ERR_INVALID; SFE_BIND_DEVICE(c); SFE_CUDA(c, cudaFree(p)); return SFE_OK; }
extern "C" int sfe_copy_h2d(sfe_context* c, void* dst, const void* src, size_t bytes)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    SFE_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFE_OK;
}
extern "C" int sfe_copy_d2h(sfe_context* c, void* dst, const void* src, size_t bytes)
{
    SFE_BIND_DEVICE(c);
    if (!c) return SFE_ERR_INVALID;
    SFE_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return SFE_OK;
}
extern "C" int sfe_ipc_export(sfe_context* c, const void* dptr, uint8_t handle[64])
{
    SFE_BIND_DEVICE(c);
    if (!c || !dptr || !handle) return SFE_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    SFE_CUDA(c, cudaIpcGetMemHandle(&h, const_cast<void*>(dptr)));
    memcpy(handle, &h, 64);
    return SFE_OK;
}
extern "C" int sfe_ipc_import(sfe_context* c, const uint8_t handle[64], void** dptr)
{
    SFE_BIND_DEVICE(c);
    if (!c || !handle || !dptr) return SFE_ERR_INVALID;
    cudaIpcMemHandle_t h; memcpy(&h, handle, 64);
    SFE_CUDA(c, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SFE_OK;
}
extern "C" int sfe_ipc_close(sfe_context* c, void* dptr) { if (!c) return SFE_ERR_INVALID; SFE_BIND_DEVICE(c); SFE_CUDA(c, cudaIpcCloseMemHandle(dptr)); return SFE_OK; }

// ---- timing ---------------------------------------------------------------------------------------------
extern "C" int sfe_timer_start(sfe_context* c) { if (!c) return SFE_ERR_INVALID; SFE_BIND_DEVICE(c); SFE_CUDA(c, cudaEventRecord(c->ev_start, c->stream)); return SFE_OK; }
extern "C" int sfe_stream_wait(sfe_context* c, sfe_context* other)
{
    SFE_BIND_DEVICE(c);
    if (!c || !other) return SFE_ERR_INVALID;
    if (c == other) return SFE_OK;
    if (c->cfg.device != other->cfg.device) { c->err = "sfe_stream_wait: contexts live on different devices"; return SFE_ERR_INVALID; }
    if (!other->ev_join) SFE_CUDA(c, cudaEventCreateWithFlags(&other->ev_join, cudaEventDisableTiming));
    SFE_CUDA(c, cudaEventRecord(other->ev_join, other->stream));
    SFE_CUDA(c, cudaStreamWaitEvent(c->stream, other->ev_join, 0));
    return SFE_OK;
}
extern "C" int sfe_timer_stop(sfe_context* c, float* ms)
{
    SFE_BIND_DEVICE(c);
    if (!c || !ms) return SFE_ERR_INVALID;
    SFE_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    SFE_CUDA(c, cudaEventSynchronize(c->ev_stop));
    SFE_CUDA(c, cudaEventElapsedTime(ms, c->ev_start, c->ev_stop));
    return SFE_OK;
}
extern "C" int sfe_set_profiling(sfe_context* c, int on) { if (!c) return SFE_ERR_INVALID; c->profiling = on != 0; return SFE_OK; }
extern "C" int sfe_num_stages(void) { return ST_COUNT; }
extern "C" const char* sfe_stage_name(int s)
{
    static const char* n[ST_COUNT] = { "h2d", "pyramid", "fast_nms", "select_harris", "blur", "orient_brief", "match_rowband", "d2h" };
    return (s >= 0 && s < ST_COUNT) ? n[s] : "";
}
extern "C" int sfe_stage_times(sfe_context* c, float* ms, int32_t* launches)
{
    SFE_BIND_DEVICE(c);
    if (!c || !ms) return SFE_ERR_INVALID;
    SFE_CUDA(c, cudaStreamSynchronize(c->stream));
    const int L = c->pack.nlevels;
    const int nl[ST_COUNT] = { 0, L - 1, 1, 1, describe_is_fused(c) ? 0 : 1, 1, 2, 0 };
    for (int s = 0; s < ST_COUNT; ++s) {
        ms[s] = 0.f;
        if (s == ST_H2D || s == ST_D2H) { if (cudaEventElapsedTime(&ms[s], c->ev_stage[s], c->ev_stage[s + 1]) != cudaSuccess) { ms[s] = 0.f; cudaGetLastError(); } }
        else SFE_CUDA(c, cudaEventElapsedTime(&ms[s], c->ev_stage[s], c->ev_stage[s + 1]));
        if (launches) launches[s] = nl[s];
    }
    return SFE_OK;
}
extern "C" int64_t sfe_kernel_launches(const sfe_context* c) { return c ? c->launches : 0; }
extern "C" int64_t sfe_last_d2h_bytes(const sfe_context* c) { return c ? (int64_t)c->d2h_bytes_last : 0; }
extern "C" int64_t sfe_d2h_refetches(const sfe_context* c) { return c ? c->d2h_refetches : 0; }
