// POPC / LOP3 / xor+popc+add throughput on B200 (sm_100a): the denominator for "popc-pipe % of peak" of the matchers.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = __popc(a[i]) + a[i]; }                  // dependent POPC chains x8 (ILP 8): POPC + IADD
            else if (MODE == 1) { a[i] = (a[i] ^ (a[i] >> 3)) + 0x1234567u; } // ALU only
            else { acc += __popc(a[i] ^ (seed + it)); a[i] += 0x01000193u; } // matcher inner loop shape: xor, popc, add
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> double run(const char* name, int ops_per_iter)
{
    const int blocks = 148 * 8, threads = 256, iters = 20000;
    uint32_t* d; cudaMalloc(&d, blocks * threads * 4);
    k<MODE><<<blocks, threads>>>(d, 100, 1); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, iters, 7); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * ops_per_iter;
    printf("%-28s %8.3f ms  %7.3f T thread-ops/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
    return ops / (ms * 1e-3);
}
int main()
{
    run<0>("popc (+iadd) chains", 8);
    run<1>("alu (xor/shift/add)", 8);
    run<2>("xor+popc+add (matcher loop)", 8);
    return 0;
}
