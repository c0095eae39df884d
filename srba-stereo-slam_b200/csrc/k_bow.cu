// k_bow.cu -- DBoW2 vocabulary transform, BoW vector, database add and L1 query on the GPU.
//
// Replaces CBoWManager (reference src/CBoWManager.h:44-111): DBoW2 1.0.2 BriefVocabulary::transform,
// BriefDatabase::add (CBoWManager.h:71-76) and BriefDatabase::query(features, ret, 4)
// (CBoWManager.h:83-88, EST.cpp:426).  Descriptor bytes keep OpenCV order, which is exactly the bit
// order m_change_structure_binary produces (CBoWManager.h:95-109), so the 256-iteration host repack
// disappears.  Semantics: descend from the root taking the child with the smallest
// Hamming distance (first child in file order on ties) until a leaf; v[word] += weight for weight > 0
// (TF-IDF, accumulated in feature order); L1-normalise in ascending word order; query score per entry
// = sum over common words, ascending, of |q-d|-|q|-|d|, reported as -sum/2, best first, ties -> lower id.
// All floating point is FP64 evaluated in DBoW2's order, so scores are bit-identical to the CPU.
//
// Kernels: 8 lanes per descriptor walk the tree (one lane per child, xor+popc, 8-lane shuffle argmin);
// the vocabulary (1.18 MB for the shipped voc.yml.gz) stays resident in L2.  Integer/popc work.
#include "sfe_internal.cuh"
#include <algorithm>
#include <cstring>

#include "bow_internal.cuh"

#define BOW_CUDA(b, call) SFE_CUDA((b)->ctx, call)

__device__ __forceinline__ int ham256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- transform: 8 lanes per descriptor ---------------------------------------------------------------
__global__ void __launch_bounds__(256)
bow_transform_kernel(const int32_t* __restrict__ first, const int32_t* __restrict__ nchild, const uint4* __restrict__ ndesc,
                     const int32_t* __restrict__ nword, const double* __restrict__ nweight,
                     const uint8_t* __restrict__ desc, int n, int depth, int32_t* __restrict__ word, double* __restrict__ weight)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gtid >> 3, c = gtid & 7;
    const int ii = min(i, n - 1);
    const uint4* q = reinterpret_cast<const uint4*>(desc + (size_t)ii * 32);
    const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
    int node = 0;
    for (int d = 0; d < depth; ++d) {
        const int nc = __ldg(nchild + node), fc = __ldg(first + node);
        uint32_t key = 0xffffffffu;
        if (c < nc) {
            const uint4* p = ndesc + 2 * (size_t)(fc + c);
            key = ((uint32_t)ham256(q0, q1, __ldg(p), __ldg(p + 1)) << 4) | (uint32_t)c;
        }
        key = min(key, __shfl_xor_sync(0xffffffffu, key, 1));
        key = min(key, __shfl_xor_sync(0xffffffffu, key, 2));
        key = min(key, __shfl_xor_sync(0xffffffffu, key, 4));
        if (nc > 0) node = fc + (int)(key & 15u);
    }
    if (c == 0 && i < n) { word[i] = __ldg(nword + node); weight[i] = __ldg(nweight + node); }
}

// ---- BoW vector of one keyframe: sort (word, index), segment-sum in feature order, L1-normalise --------
#define BV_THREADS 1024
__global__ void __launch_bounds__(BV_THREADS)
bow_vector_kernel(const int32_t* __restrict__ word, const double* __restrict__ weight, int n, int npow2,
                  int32_t* __restrict__ out_w, double* __restrict__ out_v, int32_t* __restrict__ out_n)
{
    extern __shared__ uint32_t keys[];          // npow2 keys: word<<16 | index  (n <= 65535, word < 32768)
    __shared__ int s_m;
    __shared__ double s_norm;
    const int tid = threadIdx.x;
    for (int i = tid; i < npow2; i += BV_THREADS)
        keys[i] = (i < n && weight[i] > 0.0) ? (((uint32_t)word[i] << 16) | (uint32_t)i) : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += BV_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (tid == 0) s_m = 0;
    __syncthreads();
    // heads of runs, in ascending word order: rank = number of heads before me (serial per head, n small)
    for (int i = tid; i < npow2; i += BV_THREADS) {
        const uint32_t k = keys[i];
        if (k == 0xffffffffu) continue;
        const uint32_t w = k >> 16;
        if (i > 0 && (keys[i - 1] >> 16) == w) continue;
        double s = 0.0;
        for (int j = i; j < npow2 && keys[j] != 0xffffffffu && (keys[j] >> 16) == w; ++j) s += weight[keys[j] & 0xffffu];
        int rank = 0;
        for (int j = 1; j <= i; ++j) rank += (keys[j] >> 16) != (keys[j - 1] >> 16);
        out_w[rank] = (int32_t)w; out_v[rank] = s;
        atomicAdd(&s_m, 1);
    }
    __syncthreads();
    const int m = s_m;
    if (tid == 0) {
        double norm = 0.0;
        for (int i = 0; i < m; ++i) norm += fabs(out_v[i]);
        s_norm = norm; *out_n = m;
    }
    __syncthreads();
    const double norm = s_norm;
    if (norm > 0.0) for (int i = tid; i < m; i += BV_THREADS) out_v[i] = out_v[i] / norm;
}

// ---- query: scatter q into a dense vector, one thread per database entry, gather + accumulate ----------
__global__ void bow_scatter_kernel(const int32_t* __restrict__ w, const double* __restrict__ v, const int32_t* __restrict__ n,
                                   double* __restrict__ dense, int clear)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *n) dense[w[i]] = clear ? 0.0 : v[i];
}

__global__ void bow_score_kernel(const int32_t* __restrict__ entry_ofs, int nentries, const int32_t* __restrict__ dbw,
                                 const double* __restrict__ dbv, const double* __restrict__ dense,
                                 double* __restrict__ score, int32_t* __restrict__ common)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nentries) return;
    double s = 0.0; int any = 0;
    for (int p = entry_ofs[e]; p < entry_ofs[e + 1]; ++p) {
        const double q = dense[dbw[p]];
        if (q != 0.0) { const double d = dbv[p]; s += fabs(q - d) - fabs(q) - fabs(d); any = 1; }
    }
    score[e] = s; common[e] = any;
}

__global__ void __launch_bounds__(1024)
bow_topk_kernel(const double* __restrict__ score, int32_t* __restrict__ common, int nentries, int k,
                int32_t* __restrict__ ids, double* __restrict__ out, int32_t* __restrict__ nout)
{
    __shared__ double sv[32]; __shared__ int si[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int found = 0;
    for (int r = 0; r < k; ++r) {
        double bv = 1e300; int bi = 0x7fffffff;
        for (int e = tid; e < nentries; e += 1024)
            if (common[e] && (score[e] < bv || (score[e] == bv && e < bi))) { bv = score[e]; bi = e; }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = sv[lane]; bi = si[lane];
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { sv[0] = bv; si[0] = bi; }
        }
        __syncthreads();
        bi = si[0]; bv = sv[0];
        if (bi == 0x7fffffff) break;
        if (tid == 0) { ids[found] = bi; out[found] = -bv / 2.0; common[bi] = 0; }
        ++found;
        __syncthreads();
    }
    if (tid == 0) *nout = found;
}

// ---- host side -------------------------------------------------------------------------------------------
extern "C" int sfe_bow_create(sfe_context* ctx, int nnodes, int nwords, const int32_t* child_ofs, const int32_t* child_idx,
                              const uint8_t* node_desc, const double* node_weight, const int32_t* node_word, sfe_bow** out)
{
    SFE_BIND_DEVICE(ctx);
    if (!ctx || !out || nnodes < 2 || nwords < 1 || nwords > 32768 || !child_ofs || !child_idx || !node_desc || !node_weight || !node_word)
        return SFE_ERR_INVALID;
    sfe_bow* b = new sfe_bow(); b->ctx = ctx; b->nnodes = nnodes; b->nwords = nwords;
    // BFS renumbering (children of one node become contiguous, file order preserved)
    std::vector<int32_t> order; order.reserve(nnodes); order.push_back(0);
    std::vector<int32_t> first(nnodes, 0), nch(nnodes, 0), depth_of(nnodes, 0);
    int maxdepth = 0;
    for (size_t h = 0; h < order.size(); ++h) {
        const int id = order[h];
        const int n = child_ofs[id + 1] - child_ofs[id];
        if (n > 8) { ctx->err = "vocabulary branching factor > 8 is not supported"; delete b; return SFE_ERR_INVALID; }
        first[h] = (int32_t)order.size(); nch[h] = n;
        for (int k = 0; k < n; ++k) { order.push_back(child_idx[child_ofs[id] + k]); }
    }
    if ((int)order.size() != nnodes) { ctx->err = "vocabulary tree is not connected"; delete b; return SFE_ERR_INVALID; }
    {   // depth by BFS position
        std::vector<int32_t> dep(nnodes, 0);
        for (int h = 0; h < nnodes; ++h) for (int k = 0; k < nch[h]; ++k) { dep[first[h] + k] = dep[h] + 1; maxdepth = std::max(maxdepth, dep[h] + 1); }
    }
    b->depth = maxdepth;
    std::vector<uint8_t> desc((size_t)nnodes * 32); std::vector<int32_t> word(nnodes); std::vector<double> weight(nnodes);
    for (int h = 0; h < nnodes; ++h) {
        memcpy(&desc[(size_t)h * 32], node_desc + (size_t)order[h] * 32, 32);
        word[h] = node_word[order[h]]; weight[h] = node_weight[order[h]];
    }
    cudaError_t e = cudaSuccess;
    auto up = [&](void** d, const void* h, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(d, bytes); if (e == cudaSuccess) e = cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice); };
    up((void**)&b->d_first, first.data(), sizeof(int32_t) * nnodes);
    up((void**)&b->d_nchild, nch.data(), sizeof(int32_t) * nnodes);
    up((void**)&b->d_desc, desc.data(), (size_t)nnodes * 32);
    up((void**)&b->d_word, word.data(), sizeof(int32_t) * nnodes);
    up((void**)&b->d_weight, weight.data(), sizeof(double) * nnodes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_qdense, sizeof(double) * nwords);
    if (e == cudaSuccess) e = cudaMemset(b->d_qdense, 0, sizeof(double) * nwords);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_top_id, sizeof(int32_t) * 64);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_top_s, sizeof(double) * 64);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_top_n, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_vn, sizeof(int32_t));
    if (e != cudaSuccess) { ctx->err = std::string("sfe_bow_create: ") + cudaGetErrorString(e); sfe_bow_destroy(b); return SFE_ERR_CUDA; }
    cudaFuncSetAttribute(bow_vector_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4);
    cudaGetLastError();
    *out = b;
    return SFE_OK;
}

extern "C" void sfe_bow_destroy(sfe_bow* b)
{
    SFE_BIND_DEVICE((b ? b->ctx : nullptr));
    if (!b) return;
    cudaFree(b->d_first); cudaFree(b->d_nchild); cudaFree(b->d_desc); cudaFree(b->d_word); cudaFree(b->d_weight);
    cudaFree(b->d_in); cudaFree(b->d_w); cudaFree(b->d_wt); cudaFree(b->d_vw); cudaFree(b->d_vv); cudaFree(b->d_vn);
    cudaFree(b->d_entry_ofs); cudaFree(b->d_db_w); cudaFree(b->d_db_v); cudaFree(b->d_qdense); cudaFree(b->d_score);
    cudaFree(b->d_common); cudaFree(b->d_top_id); cudaFree(b->d_top_s); cudaFree(b->d_top_n);
    delete b;
}

static int bow_reserve(sfe_bow* b, int n)
{
    if (n <= b->cap_n) return SFE_OK;
    const int cap = std::max(n, 4096);
    cudaFree(b->d_in); cudaFree(b->d_w); cudaFree(b->d_wt); cudaFree(b->d_vw); cudaFree(b->d_vv);
    BOW_CUDA(b, cudaMalloc((void**)&b->d_in, (size_t)cap * 32));
    BOW_CUDA(b, cudaMalloc((void**)&b->d_w, sizeof(int32_t) * cap));
    BOW_CUDA(b, cudaMalloc((void**)&b->d_wt, sizeof(double) * cap));
    BOW_CUDA(b, cudaMalloc((void**)&b->d_vw, sizeof(int32_t) * cap));
    BOW_CUDA(b, cudaMalloc((void**)&b->d_vv, sizeof(double) * cap));
    b->cap_n = cap;
    return SFE_OK;
}

extern "C" int sfe_bow_transform_device(sfe_bow* b, const uint8_t* d_desc, int n, int32_t* d_word, double* d_weight)
{
    SFE_BIND_DEVICE((b ? b->ctx : nullptr));
    if (!b || n < 0) return SFE_ERR_INVALID;
    if (n == 0) return SFE_OK;
    const int threads = 256, blocks = (n * 8 + threads - 1) / threads;
    bow_transform_kernel<<<blocks, threads, 0, b->ctx->stream>>>(b->d_first, b->d_nchild, b->d_desc, b->d_word, b->d_weight,
                                                                 d_desc, n, b->depth, d_word, d_weight);
    b->ctx->launches++;
    BOW_CUDA(b, cudaGetLastError());
    return SFE_OK;
}

// host descriptors -> device words/weights (+ BoW vector in d_vw/d_vv/d_vn when want_vector)
static int bow_run(sfe_bow* b, const uint8_t* desc, int n, bool want_vector)
{
    if (n > 32768) { b->ctx->err = "more than 32768 descriptors per keyframe"; return SFE_ERR_INVALID; }
    int r = bow_reserve(b, n); if (r) return r;
    cudaStream_t s = b->ctx->stream;
    BOW_CUDA(b, cudaMemcpyAsync(b->d_in, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    r = sfe_bow_transform_device(b, b->d_in, n, b->d_w, b->d_wt); if (r) return r;
    if (want_vector) {
        int np2 = 1; while (np2 < n) np2 <<= 1;
        bow_vector_kernel<<<1, BV_THREADS, (size_t)np2 * 4, s>>>(b->d_w, b->d_wt, n, np2, b->d_vw, b->d_vv, b->d_vn);
        b->ctx->launches++;
        BOW_CUDA(b, cudaGetLastError());
    }
    return SFE_OK;
}

extern "C" int sfe_bow_transform(sfe_bow* b, const uint8_t* desc, int n, int32_t* word, double* weight)
{
    SFE_BIND_DEVICE((b ? b->ctx : nullptr));
    if (!b || n < 0 || (n > 0 && (!desc || !word || !weight))) return SFE_ERR_INVALID;
    if (n == 0) return SFE_OK;
    int r = bow_run(b, desc, n, false); if (r) return r;
    BOW_CUDA(b, cudaMemcpyAsync(word, b->d_w, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, b->ctx->stream));
    BOW_CUDA(b, cudaMemcpyAsync(weight, b->d_wt, sizeof(double) * n, cudaMemcpyDeviceToHost, b->ctx->stream));
    BOW_CUDA(b, cudaStreamSynchronize(b->ctx->stream));
    return SFE_OK;
}

extern "C" int sfe_bow_vector(sfe_bow* b, const uint8_t* desc, int n, int32_t* out_word, double* out_val, int32_t* n_out)
{
    SFE_BIND_DEVICE((b ? b->ctx : nullptr));
    if (!b || n < 0 || !n_out) return SFE_ERR_INVALID;
    *n_out = 0;
    if (n == 0) return SFE_OK;
    int r = bow_run(b, desc, n, true); if (r) return r;
    int32_t m = 0;
    BOW_CUDA(b, cudaMemcpyAsync(&m, b->d_vn, sizeof(int32_t), cudaMemcpyDeviceToHost, b->ctx->stream));
    BOW_CUDA(b, cudaStreamSynchronize(b->ctx->stream));
    if (out_word) BOW_CUDA(b, cudaMemcpyAsync(out_word, b->d_vw, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, b->ctx->stream));
    if (out_val) BOW_CUDA(b, cudaMemcpyAsync(out_val, b->d_vv, sizeof(double) * m, cudaMemcpyDeviceToHost, b->ctx->stream));
    BOW_CUDA(b, cudaStreamSynchronize(b->ctx->stream));
    *n_out = m;
    return SFE_OK;
}

extern "C" int sfe_bow_db_size(const sfe_bow* b) { return b ? (int)b->entry_ofs.size() - 1 : 0; }

extern "C" int sfe_bow_db_add(sfe_bow* b, const uint8_t* desc, int n, int32_t* entry_id)
{
    SFE_BIND_DEVICE((b ? b->ctx : nullptr));
    if (!b || n < 0 || (n > 0 && !desc)) return SFE_ERR_INVALID;
    int32_t m = 0;
    cudaStream_t s = b->ctx->stream;
    if (n > 0) {
        int r = bow_run(b, desc, n, true); if (r) return r;
        BOW_CUDA(b, cudaMemcpyAsync(&m, b->d_vn, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        BOW_CUDA(b, cudaStreamSynchronize(s));
    }
    const size_t used = (size_t)b->entry_ofs.back();
    if (used + m > b->db_cap) {
        const size_t cap = std::max<size_t>((used + m) * 2, 1 << 20);
        int32_t* nw = nullptr; double* nv = nullptr;
        BOW_CUDA(b, cudaMalloc((void**)&nw, sizeof(int32_t) * cap));
        BOW_CUDA(b, cudaMalloc((void**)&nv, sizeof(double) * cap));
        if (used) {
            BOW_CUDA(b, cudaMemcpyAsync(nw, b->d_db_w, sizeof(int32_t) * used, cudaMemcpyDeviceToDevice, s));
            BOW_CUDA(b, cudaMemcpyAsync(nv, b->d_db_v, sizeof(double) * used, cudaMemcpyDeviceToDevice, s));
            BOW_CUDA(b, cudaStreamSynchronize(s));
        }
        cudaFree(b->d_db_w); cudaFree(b->d_db_v);
        b->d_db_w = nw; b->d_db_v = nv; b->db_cap = cap;
    }
    if (m) {
        BOW_CUDA(b, cudaMemcpyAsync(b->d_db_w + used, b->d_vw, sizeof(int32_t) * m, cudaMemcpyDeviceToDevice, s));
        BOW_CUDA(b, cudaMemcpyAsync(b->d_db_v + used, b->d_vv, sizeof(double) * m, cudaMemcpyDeviceToDevice, s));
    }
    b->entry_ofs.push_back((int32_t)(used + m));
    const int ne = (int)b->entry_ofs.size() - 1;
    if (ne + 1 > b->ofs_cap) {
        const int cap = std::max(2 * (ne + 1), 4096);
        cudaFree(b->d_entry_ofs); cudaFree(b->d_score); cudaFree(b->d_common);
        BOW_CUDA(b, cudaMalloc((void**)&b->d_entry_ofs, sizeof(int32_t) * cap));
        BOW_CUDA(b, cudaMalloc((void**)&b->d_score, sizeof(double) * cap));
        BOW_CUDA(b, cudaMalloc((void**)&b->d_common, sizeof(int32_t) * cap));
        b->ofs_cap = cap;
        BOW_CUDA(b, cudaMemcpyAsync(b->d_entry_ofs, b->entry_ofs.data(), sizeof(int32_t) * (ne + 1), cudaMemcpyHostToDevice, s));
    } else {
        BOW_CUDA(b, cudaMemcpyAsync(b->d_entry_ofs + ne, &b->entry_ofs[ne], sizeof(int32_t), cudaMemcpyHostToDevice, s));
    }
    BOW_CUDA(b, cudaStreamSynchronize(s));
    if (entry_id) *entry_id = ne - 1;
    return SFE_OK;
}

extern "C" int sfe_bow_db_query(sfe_bow* b, const uint8_t* desc, int n, int max_results, int32_t* ids, double* scores, int32_t* n_out)
{
    SFE_BIND_DEVICE((b ? b->ctx : nullptr));
    if (!b || n < 0 || !n_out || max_results < 1 || max_results > 64 || !ids || !scores) return SFE_ERR_INVALID;
    *n_out = 0;
    const int ne = (int)b->entry_ofs.size() - 1;
    if (n == 0 || ne == 0) return SFE_OK;
    int r = bow_run(b, desc, n, true); if (r) return r;
    cudaStream_t s = b->ctx->stream;
    const int tb = (n + 255) / 256;
    bow_scatter_kernel<<<tb, 256, 0, s>>>(b->d_vw, b->d_vv, b->d_vn, b->d_qdense, 0);
    bow_score_kernel<<<(ne + 127) / 128, 128, 0, s>>>(b->d_entry_ofs, ne, b->d_db_w, b->d_db_v, b->d_qdense, b->d_score, b->d_common);
    bow_scatter_kernel<<<tb, 256, 0, s>>>(b->d_vw, b->d_vv, b->d_vn, b->d_qdense, 1);
    bow_topk_kernel<<<1, 1024, 0, s>>>(b->d_score, b->d_common, ne, max_results, b->d_top_id, b->d_top_s, b->d_top_n);
    b->ctx->launches += 4;
    BOW_CUDA(b, cudaGetLastError());
    int32_t m = 0;
    BOW_CUDA(b, cudaMemcpyAsync(&m, b->d_top_n, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    BOW_CUDA(b, cudaMemcpyAsync(ids, b->d_top_id, sizeof(int32_t) * max_results, cudaMemcpyDeviceToHost, s));
    BOW_CUDA(b, cudaMemcpyAsync(scores, b->d_top_s, sizeof(double) * max_results, cudaMemcpyDeviceToHost, s));
    BOW_CUDA(b, cudaStreamSynchronize(s));
    *n_out = m;
    return SFE_OK;
}
