#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../../srba-stereo-slam_b200/csrc/sfe_internal.cuh"
__global__ void k(const __grid_constant__ CUtensorMap m, int x, int y, int z, uint8_t* out, int bw, int bh)
{
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_arrive_expect_tx(&bar, bw * bh); tma_load_3d(tile, &m, x, y, z, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    PFN enc = (PFN)fn;
    int w = 1241, h = 376, pitch = 1280, n = 2;
    uint8_t* d; cudaMalloc(&d, (size_t)pitch * h * n);
    std::vector<uint8_t> host((size_t)pitch * h * n);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (uint8_t)(i * 7 + (i / pitch));
    cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice);
    uint8_t* out; cudaMalloc(&out, 65536);
    int boxes[][2] = { {160, 24}, {176, 24}, {144, 24}, {256, 16} };
    for (auto& b : boxes) {
        CUtensorMap m;
        cuuint64_t dims[3] = { (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n };
        cuuint64_t strides[2] = { (cuuint64_t)pitch, (cuuint64_t)pitch * h };
        cuuint32_t box[3] = { (cuuint32_t)b[0], (cuuint32_t)b[1], 1 };
        cuuint32_t es[3] = { 1, 1, 1 };
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int coords[][3] = { {0, 0, 0}, {16, 3, 1}, {144, 326, 0}, {-16, -4, 1}, {1200, 370, 1}, {-32, 375, 0}, {1232, -7, 1} };
        for (auto& c : coords) {
            k<<<1, 128, b[0] * b[1]>>>(m, c[0], c[1], c[2], out, b[0], b[1]);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint8_t> o(b[0] * b[1]);
            int bad = -1;
            if (e == cudaSuccess) {
                cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
                bad = 0;
                for (int yy = 0; yy < b[1]; ++yy) for (int xx = 0; xx < b[0]; ++xx) {
                    int gx = c[0] + xx, gy = c[1] + yy;
                    uint8_t exp = (gx < 0 || gy < 0 || gx >= w || gy >= h) ? 0 : host[(size_t)c[2] * pitch * h + (size_t)gy * pitch + gx];
                    bad += o[yy * b[0] + xx] != exp;
                }
            }
            printf("enc=%d box %dx%d coord (%d,%d,%d): %s bad=%d\n", (int)r, b[0], b[1], c[0], c[1], c[2], cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
